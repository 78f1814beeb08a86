"""2 s DE/PSD driver -- drop-in for /root/reference/EEG_preprocessing/extract_DE_PSD_features_1per2s.py.

Input (7, 40, 5, 62, 400); output DE and PSD (7, 40, 5, 62, 5) float32.  One kernel launch per subject instead of
1400 DE_PSD calls.
"""
import numpy as np

from .. import frontend
from . import _io

fre = 200


def extract_de_psd_raw(raw, fs=200):
    """(B, C, R, ch, 400) -> (DE, PSD), each (B, C, R, ch, 5) float32 (reference :16-28)."""
    if raw.ndim != 5:
        raise ValueError("raw must be (blocks, concepts, repetitions, channels, samples)")
    if raw.shape[4] != 2 * fre:
        raise ValueError(f"cannot reshape array of size {raw.shape[3] * raw.shape[4]} into shape "
                         f"({raw.shape[3]},{2 * fre})")                 # the reference's reshape error (:23)
    if int(2 * fs) != 2 * fre:
        # the reference reshapes with the module constant (:23) but windows with `fs` (:24): any other rate ends in
        # numpy's broadcast error inside DE_PSD (DE_PSD.py:57)
        raise ValueError(f"operands could not be broadcast together with shapes ({2 * fre},) ({int(2 * fs)},) ")
    like_torch = _io.is_torch(raw)
    de, psd = frontend.de_psd_from_clips(_io.to_device_f32(raw), "2s", check=True)
    return _io.finish((de, psd), like_torch, np.float32)


def main(in_dir="./data/Preprocessing/Segmented_Rawf_200Hz_2s", de_dir="./data/Preprocessing/DE_1per2s",
         psd_dir="./data/Preprocessing/PSD_1per2s", subjects=range(1, 21)):
    """Script behaviour of the reference (:30-42): subjects 1..20, DE_1per2s/ and PSD_1per2s/ next to the input."""
    return _io.convert_directory(in_dir, (de_dir, psd_dir), lambda clips: extract_de_psd_raw(clips, fre),
                                 names=[f"sub{n}.npy" for n in subjects])


if __name__ == "__main__":
    main()
