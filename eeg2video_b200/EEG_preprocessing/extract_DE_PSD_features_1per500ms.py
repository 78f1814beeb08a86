"""500 ms sliding-window DE/PSD driver -- drop-in for
/root/reference/EEG_preprocessing/extract_DE_PSD_features_1per500ms.py (same CLI flags and defaults).
"""
import argparse

import numpy as np

from .. import frontend
from . import _io


def _clips_behind_window_view(raw):
    """If ``raw`` is the strided view seg_sliding_window(clips, 0.5, 0.25) returns, recover ``clips`` (no copy),
    so that the fused kernel reads 1600 B per channel-clip instead of a materialised 2800 B."""
    if raw.ndim != 6 or raw.shape[3] != 7 or raw.shape[5] != 100:
        return None
    if _io.is_torch(raw):
        st, unit = raw.stride(), 1
        if st[5] != unit or st[3] != 50 * unit or st[4] != 400 * unit:
            return None
        return raw.as_strided(raw.shape[:3] + (raw.shape[4], 400), (st[0], st[1], st[2], st[4], st[5]),
                              raw.storage_offset())
    st, unit = raw.strides, raw.itemsize
    if st[5] != unit or st[3] != 50 * unit or st[4] != 400 * unit:
        return None
    return np.lib.stride_tricks.as_strided(raw, shape=raw.shape[:3] + (raw.shape[4], 400),
                                           strides=(st[0], st[1], st[2], st[4], st[5]), writeable=False)


def extract_de_psd_sw(raw, fs, win_sec):
    """(B, C, R, W, ch, L) windows -> (DE, PSD), each (B, C, R, W, ch, 5) float32 (reference :12-29)."""
    if not fs > 0:
        raise ValueError(f"fs must be positive, got {fs!r}")
    if raw.ndim != 6:
        raise ValueError("raw must be (blocks, concepts, repetitions, windows, channels, samples)")
    length = int(win_sec * fs)
    if raw.shape[5] != length or length < 1:
        raise ValueError(f"operands could not be broadcast together with shapes ({raw.shape[5]},) ({length},) ")
    like_torch = _io.is_torch(raw)
    clips = _clips_behind_window_view(raw) if (length == 100 and fs == 200) else None
    if clips is not None:
        de, psd = frontend.de_psd_from_clips(_io.to_device_f32(clips), "500ms", check=True)
    else:
        # any other (fs, win_sec) the reference's DE_PSD accepts: the general kernel behind de_psd_windows
        de, psd = frontend.de_psd_windows(_io.to_device_f32(raw), check=True, fre=fs)
    return _io.finish((de, psd), like_torch, np.float32)


def main(argv=None):
    """CLI of the reference script (:32-58): --raw_dir --de_dir --psd_dir --subs, same defaults."""
    cli = argparse.ArgumentParser(description=__doc__)
    cli.add_argument("--raw_dir", default="./data/Preprocessing/Segmented_500ms_sw")
    cli.add_argument("--de_dir", default="./data/Preprocessing/DE_500ms_sw")
    cli.add_argument("--psd_dir", default="./data/Preprocessing/PSD_500ms_sw")
    cli.add_argument("--subs", nargs="+", type=int, default=list(range(1, 21)))
    opt = cli.parse_args(argv)
    return _io.convert_directory(opt.raw_dir, (opt.de_dir, opt.psd_dir), lambda windows: extract_de_psd_sw(windows, 200, 0.5),
                                 names=[f"sub{n}.npy" for n in opt.subs])


if __name__ == "__main__":
    main()
