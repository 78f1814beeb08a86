"""seg_sliding_window -- drop-in for /root/reference/EEG_preprocessing/segment_sliding_window.py.

The function returns a strided, read-only VIEW exactly like the reference (:11-19); nothing is copied.  The
script entry point materialises the windows with the device gather kernel (bit-exact) before saving.
"""
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


def main(in_dir="./data/Preprocessing/Segmented_Rawf_200Hz_2s", fs=200, win_s=0.5, step_s=0.25):
    """Script behaviour of the reference (:24-57): 500 ms windows every 250 ms -> Segmented_500ms_sw/."""
    if (int(fs * win_s), int(fs * step_s)) != (100, 50):
        raise NotImplementedError("the materialising kernel is built for 100-sample windows every 50 samples")
    out_dir = f"./data/Preprocessing/Segmented_{int(1000 * win_s)}ms_sw"

    def wrong_shape(clips):
        return None if clips.ndim == 5 and clips.shape[-1] == 2 * fs else f"unexpected shape {clips.shape}"
    return _io.convert_directory(in_dir, (out_dir,), materialize_windows, accept=wrong_shape)


if __name__ == "__main__":
    main()
