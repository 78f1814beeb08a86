"""seg_sliding_window -- drop-in for /root/reference/EEG_preprocessing/segment_sliding_window.py.

The function returns a strided, read-only VIEW exactly like the reference (:11-19); nothing is copied.  The
script entry point materialises the windows with the device gather kernel (bit-exact) before saving.
"""
import os

import numpy as np
import torch
from numpy.lib.stride_tricks import as_strided

from .. import frontend
from . import _io


def seg_sliding_window(data, win_s, step_s, fs=200):
    """(B, C, R, ch, T) -> view (B, C, R, W, ch, win): windows of int(fs*win_s) samples every int(fs*step_s)."""
    win_t = int(fs * win_s)
    step_t = int(fs * step_s)
    if data.ndim != 5:
        raise ValueError("axes don't match array")                    # the reference's transpose error (:19)
    n_time = data.shape[-1]
    if win_t > n_time:
        raise ValueError("window shape cannot be larger than input array shape")
    n_win = (n_time - win_t) // step_t + 1
    if _io.is_torch(data):
        return data.unfold(-1, win_t, step_t).permute(0, 1, 2, 4, 3, 5)
    data = np.asarray(data)
    b, c, r, ch, _ = data.shape
    s = data.strides
    return as_strided(data, shape=(b, c, r, n_win, ch, win_t),
                      strides=(s[0], s[1], s[2], s[4] * step_t, s[3], s[4]), writeable=False)


def materialize_windows(data):
    """(.., ch, 400) -> contiguous (.., 7, ch, 100) by the device gather (what np.save writes at :55)."""
    like_torch = _io.is_torch(data)
    dev = data if like_torch and data.is_cuda else (
        data.to(_io.device()) if like_torch else torch.from_numpy(np.ascontiguousarray(data)).to(_io.device()))
    out = frontend.sliding_windows(dev)
    return out if like_torch else out.cpu().numpy()


if __name__ == "__main__":

    INPUT_DIR = './data/Preprocessing/Segmented_Rawf_200Hz_2s'
    FS = 200
    WIN_S = 0.5
    STEP_S = 0.25
    OUTPUT_DIR = f'./data/Preprocessing/Segmented_{int(1000*WIN_S)}ms_sw'
    os.makedirs(OUTPUT_DIR, exist_ok=True)

    for fname in os.listdir(INPUT_DIR):
        if not fname.endswith('.npy'):
            continue
        data = np.load(os.path.join(INPUT_DIR, fname))
        if data.ndim != 5 or data.shape[-1] != 2 * FS:
            print(f"Skipping {fname}: unexpected shape {data.shape}")
            continue
        windows = materialize_windows(data)
        np.save(os.path.join(OUTPUT_DIR, fname), windows)
        print(f"Saved segmented windows for {fname} -> {windows.shape}")
