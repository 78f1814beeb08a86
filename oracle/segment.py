"""Oracle restatement of the segmentation index arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/EEG_preprocessing/segment_raw_signals_200Hz.py:15-110 and
segment_sliding_window.py:6-21.  Integer index work only: parity is bit-exact.
"""
import numpy as np

FS = 200
HINT_SEC = 3            # segment_raw_signals_200Hz.py:10  (_BASELINE_SEC)
REPS_PER_CONCEPT = 5    # :11
CONCEPTS_PER_BLOCK = 40  # :12


def clip_start(concept, repetition, fs=FS):
    """Sample offset of clip (concept, repetition) inside a block (:58-64):
    concept * (3 fs + 5 * 2 fs) + 3 fs + repetition * 2 fs."""
    return concept * (HINT_SEC * fs + REPS_PER_CONCEPT * 2 * fs) + HINT_SEC * fs + repetition * 2 * fs


def extract_2s_segment(*, block, concept, repetition, fs=FS, data):
    """(:15-70) with ``data`` given: range checks, then the (channels, 2 fs) slice of ``data[block]``."""
    if not 0 <= block <= 6:
        raise ValueError("`block` must be in [0, 6]")
    if not 0 <= concept < CONCEPTS_PER_BLOCK:
        raise ValueError("`concept` must be in [0, 39]")
    if not 0 <= repetition < REPS_PER_CONCEPT:
        raise ValueError("`repetition` must be in [0, 4]")
    first = clip_start(concept, repetition, fs)
    piece = data[block][:, first:first + 2 * fs]
    if piece.shape[1] != 2 * fs:
        raise RuntimeError("Segment length mismatch")
    return piece


def segment_subject(raw, fs=FS):
    """What segment_all_files (:73-110) builds for one subject: (7, ch, T) -> (7, 40, 5, ch, 2 fs)."""
    raw = np.asarray(raw)
    out = np.empty((7, CONCEPTS_PER_BLOCK, REPS_PER_CONCEPT, raw.shape[1], 2 * fs), dtype=raw.dtype)
    for b in range(7):
        for c in range(CONCEPTS_PER_BLOCK):
            for r in range(REPS_PER_CONCEPT):
                out[b, c, r] = extract_2s_segment(block=b, concept=c, repetition=r, fs=fs, data=raw)
    return out


def seg_sliding_window(data, win_s, step_s, fs=200):
    """segment_sliding_window.py:6-21 by explicit index arithmetic (a copy, not a view):
    (B,C,R,ch,T) -> (B,C,R,W,ch,win) with window w covering samples [w*step, w*step + win)."""
    data = np.asarray(data)
    win = int(fs * win_s)
    step = int(fs * step_s)
    n_all = data.shape[-1] - win + 1            # sliding_window_view length (:11)
    starts = range(0, n_all, step)              # [..., ::step, :] (:15)
    out = np.empty(data.shape[:3] + (len(starts), data.shape[3], win), dtype=data.dtype)
    for w, s in enumerate(starts):
        out[:, :, :, w] = data[..., s:s + win]
    return out


def seq2seq_windows(clips, window_size=100, overlap=50):
    """EEG2Video_New/Seq2Seq/my_autoregressive_transformer.py:309-314, the trainer's inline sliding window:
    EEG = stack([x[..., i:i + window_size] for i in range(0, T - window_size + 1, window_size - overlap)], axis=-1),
    i.e. (..., ch, 400) -> (..., ch, 100, 7) with the window index LAST."""
    clips = np.asarray(clips)
    pieces = [clips[..., i:i + window_size]
              for i in range(0, clips.shape[-1] - window_size + 1, window_size - overlap)]
    return np.stack(pieces, axis=-1)
