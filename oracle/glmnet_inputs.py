"""Oracle for the GLMNet input build (TEST INFRASTRUCTURE; see oracle/__init__.py).

PARITY UNPINNED for this row: the reference only *describes* the step (README.md:80-81, :88, :97-99 -- "Raw EEGs are
normalized per channel using the training split statistics"); the scripts that implement it
(EEG2Video/GLMNet/train_glmnet.py, inference_glmnet.py) are not in /root/reference.  This file states the arithmetic:
statistics over the 2 s clip samples of the training blocks (population std, float64), y = (x - mean) / std.
"""
import numpy as np

from .segment import segment_subject


def channel_stats(raw, train_blocks=None):
    """raw (n_blocks, ch, T) -> (mean, std) float64 (ch,) over the clip samples of the selected blocks."""
    raw = np.asarray(raw)
    n_blocks = raw.shape[0]
    pad = np.zeros((max(0, 7 - n_blocks),) + raw.shape[1:], raw.dtype)
    clips = segment_subject(np.concatenate([raw[:7], pad]))[:n_blocks].astype(np.float64)      # (B,40,5,ch,400)
    if n_blocks > 7:
        raise ValueError("oracle handles up to 7 blocks at a time")
    sel = np.arange(n_blocks) if train_blocks is None else np.asarray(train_blocks)
    x = clips[sel]
    return x.mean(axis=(0, 1, 2, 4)), x.std(axis=(0, 1, 2, 4))


def normalised_clips(raw, mean, std):
    """raw (n_blocks <= 7, ch, T) -> float32 (n_blocks, 40, 5, 1, ch, 400): (x - mean[ch]) / std[ch]."""
    raw = np.asarray(raw)
    n_blocks = raw.shape[0]
    pad = np.zeros((max(0, 7 - n_blocks),) + raw.shape[1:], raw.dtype)
    clips = segment_subject(np.concatenate([raw[:7], pad]))[:n_blocks].astype(np.float64)
    y = (clips - np.asarray(mean)[None, None, None, :, None]) / np.asarray(std)[None, None, None, :, None]
    return y[:, :, :, None].astype(np.float32)
