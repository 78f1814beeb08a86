"""1 s DE/PSD driver -- drop-in for /root/reference/EEG_preprocessing/extract_DE_PSD_features_1per1s.py.

The reference file is a module-level script (it runs on import, :24-58).  Here the body is the callable
``extract_de_psd_1s`` and the script behaviour (walk ./data/Preprocessing/Segmented_Rawf_200Hz_2s/, write
DE_1per1s/ and PSD_1per1s/) lives in ``main()``, run with
``python -m eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per1s``.
"""
import os

import numpy as np

from .. import frontend
from . import _io

fre = 200


def extract_de_psd_1s(raw, fs=200):
    """(B, C, R, ch, 400) -> (DE, PSD), each (B, C, R, 2, ch, 5) float64; window k = samples [200k, 200k+200)
    (reference :34-35, :46-53)."""
    _io.check_fs(fs)
    if raw.ndim != 5 or raw.shape[4] != 2 * fre:
        raise ValueError("raw must be (blocks, concepts, repetitions, channels, 400)")
    like_torch = _io.is_torch(raw)
    de, psd = frontend.de_psd_from_clips(_io.to_device_f32(raw), "1s", check=True)
    return _io.finish((de, psd), like_torch, np.float64)


def main(in_dir="./data/Preprocessing/Segmented_Rawf_200Hz_2s/",
         de_dir="./data/Preprocessing/DE_1per1s", psd_dir="./data/Preprocessing/PSD_1per1s"):
    """Script behaviour of the reference (:17-24, :56-58): every file found under ``in_dir``."""
    names = sorted(f for _, _, files in os.walk(in_dir) for f in files)
    return _io.convert_directory(in_dir, (de_dir, psd_dir), lambda clips: extract_de_psd_1s(clips, fre), names=names)


if __name__ == "__main__":
    main()
