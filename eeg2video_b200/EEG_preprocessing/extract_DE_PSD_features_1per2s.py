"""2 s DE/PSD driver -- drop-in for /root/reference/EEG_preprocessing/extract_DE_PSD_features_1per2s.py.

Input (7, 40, 5, 62, 400); output DE and PSD (7, 40, 5, 62, 5) float32.  One kernel launch per subject instead of
1400 DE_PSD calls.
"""
import os

import numpy as np

from .. import frontend
from . import _io

fre = 200


def extract_de_psd_raw(raw, fs=200):
    """(B, C, R, ch, 400) -> (DE, PSD), each (B, C, R, ch, 5) float32 (reference :16-28)."""
    _io.check_fs(fs)
    if raw.ndim != 5:
        raise ValueError("raw must be (blocks, concepts, repetitions, channels, samples)")
    if raw.shape[4] != 2 * fre:
        raise ValueError(f"cannot reshape array of size {raw.shape[3] * raw.shape[4]} into shape "
                         f"({raw.shape[3]},{2 * fre})")                 # the reference's reshape error (:23)
    like_torch = _io.is_torch(raw)
    de, psd = frontend.de_psd_from_clips(_io.to_device_f32(raw), "2s", check=True)
    return _io.finish((de, psd), like_torch, np.float32)


if __name__ == "__main__":
    for subname in range(1, 21):
        loaded_data = np.load('./data/Preprocessing/Segmented_Rawf_200Hz_2s/sub' + str(subname) + '.npy')
        print("Successfully loaded .npy file.")
        DE_data, PSD_data = extract_de_psd_raw(loaded_data, fre)

        os.makedirs("./data/Preprocessing/DE_1per2s", exist_ok=True)
        os.makedirs("./data/Preprocessing/PSD_1per2s", exist_ok=True)
        np.save("./data/Preprocessing/DE_1per2s/sub" + str(subname) + ".npy", DE_data)
        np.save("./data/Preprocessing/PSD_1per2s/sub" + str(subname) + ".npy", PSD_data)
        print(f"Saved DE data in ./data/Preprocessing/DE_1per2s/sub{str(subname)}.npy")
        print(f"Saved PSD data in ./data/Preprocessing/PSD_1per2s/sub{str(subname)}.npy")
