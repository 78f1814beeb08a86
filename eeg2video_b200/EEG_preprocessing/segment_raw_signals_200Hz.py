"""Utilities to segment SEED-DV EEG recordings -- drop-in for
/root/reference/EEG_preprocessing/segment_raw_signals_200Hz.py (same names, defaults and exceptions).

``extract_2s_segment`` is pure index arithmetic and returns a VIEW, like the reference (:56-67); it accepts
numpy arrays (incl. memory maps) and torch tensors.  ``segment_all_files`` materialises the clip tensor with the
device gather kernel (bit-exact) instead of 1400 Python-level copies.
"""
import os

import numpy as np
import torch

from .. import frontend
from . import _io

__all__ = ["extract_2s_segment", "segment_all_files"]

FS = 200
HINT_SECONDS, CLIP_SECONDS = 3, 2           # a concept = 3 s hint + 5 clips of 2 s (SEED-DV protocol)
N_BLOCKS, N_CONCEPTS, N_REPS = 7, 40, 5
_INDEX_LIMITS = (("block", N_BLOCKS), ("concept", N_CONCEPTS), ("repetition", N_REPS))


def clip_start(concept, repetition, fs=FS):
    """First sample of clip (concept, repetition) inside a block row: c * 13 fs + 3 fs + r * 2 fs."""
    per_concept = (HINT_SECONDS + N_REPS * CLIP_SECONDS) * fs
    return concept * per_concept + HINT_SECONDS * fs + repetition * CLIP_SECONDS * fs


def _open_recording(subject, eeg_root):
    if subject is None or subject < 1:
        raise ValueError("`subject` must be >= 1 when no `data` is provided")
    path = os.path.join(eeg_root, f"sub{subject}.npy")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return np.load(path, mmap_mode="r")


def extract_2s_segment(*, block, concept, repetition, subject=None, eeg_root="./data/EEG", fs=FS, data=None):
    """One raw 2-second EEG segment, shape (channels, 2*fs), as a VIEW of ``data[block]`` -- same keyword-only
    contract, defaults and exceptions as the reference function (segment_raw_signals_200Hz.py:15-70).

    ``data``: recording (7, channels, T), numpy (memory maps included) or torch; when omitted,
    ``eeg_root/sub{subject}.npy`` is memory-mapped (subject ids start at 1).
    """
    recording = data if data is not None else _open_recording(subject, eeg_root)
    for (name, count), value in zip(_INDEX_LIMITS, (block, concept, repetition)):
        if not 0 <= value < count:
            raise ValueError(f"`{name}` must be in [0, {count - 1}]")
    lo = clip_start(concept, repetition, fs)
    view = recording[block][:, lo:lo + CLIP_SECONDS * fs]
    if view.shape[1] != CLIP_SECONDS * fs:
        raise RuntimeError("Segment length mismatch")
    return view


def segment_subject(data, fs=FS):
    """(7, ch, T) -> (7, 40, 5, ch, 2*fs), dtype preserved: the array segment_all_files saves per subject.

    numpy in -> numpy out, torch CUDA in -> torch CUDA out.  Bit-exact with 1400 calls of extract_2s_segment.
    """
    like_torch = _io.is_torch(data)
    if data.shape[-1] < clip_start(N_CONCEPTS, 0, fs) - HINT_SECONDS * fs:
        raise RuntimeError("Segment length mismatch")
    if like_torch:
        dev = data if data.is_cuda else data.to(_io.device())
    else:
        arr = np.asarray(data)
        if arr.dtype not in (np.float32, np.float64, np.float16, np.int16):
            raise TypeError(f"unsupported recording dtype {arr.dtype}")
        arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:                       # memory maps: torch wants a writable buffer to wrap
            arr = arr.copy()
        dev = torch.from_numpy(arr).to(_io.device())
    clips = frontend.segment_clips(dev, fs)
    return clips if like_torch else clips.cpu().numpy()


def segment_all_files(eeg_root="./data/EEG", output_dir="./data/Preprocessing/Segmented_Rawf_200Hz_2s", fs=FS):
    """Write one ``(7, 40, 5, channels, 2*fs)`` clip array per ``sub{N}.npy`` recording found in ``eeg_root``
    (reference: segment_raw_signals_200Hz.py:73-110; same defaults, same file naming)."""
    os.makedirs(output_dir, exist_ok=True)
    for name in sorted(os.listdir(eeg_root)):
        stem, ext = os.path.splitext(name)
        if ext != ".npy":
            continue
        int(stem.replace("sub", ""))                 # the reference's file-name contract: sub{N}.npy, else ValueError
        recording = np.load(os.path.join(eeg_root, name))
        np.save(os.path.join(output_dir, name), segment_subject(recording[:N_BLOCKS], fs))


if __name__ == "__main__":
    segment_all_files()
