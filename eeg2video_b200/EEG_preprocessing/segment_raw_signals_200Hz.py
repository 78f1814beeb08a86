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
_BASELINE_SEC = 3
_REPS_PER_CONCEPT = 5
_CONCEPTS_PER_BLOCK = 40


def extract_2s_segment(
    *,
    block,
    concept,
    repetition,
    subject=None,
    eeg_root="./data/EEG",
    fs=FS,
    data=None,
):
    """Return one raw 2-second EEG segment (62 x 2*fs) as a view of ``data[block]``.

    block, concept, repetition : indices of the segment inside a recording.
    subject : 1-indexed subject id, required when ``data`` is None (``eeg_root/sub{subject}.npy`` is memory-mapped).
    data : pre-loaded recording of shape (7, 62, T), numpy or torch.
    """
    if data is None:
        if subject is None or subject < 1:
            raise ValueError("`subject` must be >= 1 when no `data` is provided")
        path = os.path.join(eeg_root, f"sub{subject}.npy")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        data = np.load(path, mmap_mode="r")

    if not 0 <= block <= 6:
        raise ValueError("`block` must be in [0, 6]")
    if not 0 <= concept < _CONCEPTS_PER_BLOCK:
        raise ValueError("`concept` must be in [0, 39]")
    if not 0 <= repetition < _REPS_PER_CONCEPT:
        raise ValueError("`repetition` must be in [0, 4]")

    clip_len = 2 * fs
    first = concept * (_BASELINE_SEC * fs + _REPS_PER_CONCEPT * clip_len) + _BASELINE_SEC * fs + repetition * clip_len
    segment = data[block][:, first:first + clip_len]
    if segment.shape[1] != clip_len:
        raise RuntimeError("Segment length mismatch")
    return segment


def segment_subject(data, fs=FS):
    """(7, ch, T) -> (7, 40, 5, ch, 2*fs), dtype preserved: the array segment_all_files saves per subject.

    numpy in -> numpy out, torch CUDA in -> torch CUDA out.  Bit-exact with 1400 calls of extract_2s_segment.
    """
    like_torch = _io.is_torch(data)
    if data.shape[-1] < _CONCEPTS_PER_BLOCK * (_BASELINE_SEC + 2 * _REPS_PER_CONCEPT) * fs:
        raise RuntimeError("Segment length mismatch")
    if like_torch:
        dev = data if data.is_cuda else data.to(_io.device())
    else:
        arr = np.asarray(data)
        if arr.dtype not in (np.float32, np.float64, np.float16, np.int16):
            raise TypeError(f"unsupported recording dtype {arr.dtype}")
        dev = torch.from_numpy(np.ascontiguousarray(arr)).to(_io.device())
    clips = frontend.segment_clips(dev, fs)
    return clips if like_torch else clips.cpu().numpy()


def segment_all_files(
    eeg_root="./data/EEG",
    output_dir="./data/Preprocessing/Segmented_Rawf_200Hz_2s",
    fs=FS,
):
    """Segment all EEG files into ``(7, 40, 5, 62, 2*fs)`` arrays (one ``sub{N}.npy`` per input file)."""
    os.makedirs(output_dir, exist_ok=True)

    sub_list = [f for f in os.listdir(eeg_root) if f.endswith(".npy")]
    for subname in sub_list:
        int(os.path.splitext(subname)[0].replace("sub", ""))     # same file-name contract as the reference (:83)
        data = np.load(os.path.join(eeg_root, subname))
        np.save(os.path.join(output_dir, subname), segment_subject(data[:7], fs))


if __name__ == "__main__":
    segment_all_files()
