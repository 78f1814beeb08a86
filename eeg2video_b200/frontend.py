"""Device-resident front end: the fused calls a pipeline makes when the recordings already live in HBM.

These are the B200-first entry points (no clip or window tensor is ever materialised); the reference-named
functions in ``eeg2video_b200.EEG_preprocessing`` are thin shims over them.
"""
import torch

from . import _lib, ops

MODES = {"500ms": _lib.MODE_500MS, "1s": _lib.MODE_1S, "2s": _lib.MODE_2S}
WINDOWS_PER_CLIP = {"500ms": 7, "1s": 2, "2s": 1}
BLOCKS_PER_SUBJECT = 7
CONCEPTS_PER_BLOCK = 40
REPS_PER_CONCEPT = 5
MIN_BLOCK_LEN = 40 * 2600


def _mode_id(mode):
    if mode in MODES:
        return MODES[mode]
    if mode in MODES.values():
        return mode
    raise ValueError(f"mode must be one of {sorted(MODES)}")


def raise_if_zero_power(status):
    """The reference raises ValueError('math domain error') from math.log (DE_PSD.py:68) when a band has zero
    power; the kernel sets a flag instead.  This reads the flag (one 4-byte D2H copy, synchronises)."""
    if int(status.item()) & _lib.STATUS_ZERO_POWER:
        raise ValueError("math domain error")


def de_psd_from_raw(raw, mode="500ms", check=True):
    """Fused segmentation + DE/PSD from raw recordings.

    raw: float32 CUDA tensor (..., 62, T) -- e.g. (subjects, 7, 62, T) or (7, 62, T); T >= 104000.
    Returns (de, psd), float32, shape (..., 40, 5, [W,] 62, 5) with W = 7 / 2 / (absent for "2s").
    Layouts follow extract_DE_PSD_features_1per500ms.py:16-17, _1per1s.py:34-35, _1per2s.py:17-18.
    """
    if raw.dim() < 3:
        raise ValueError("raw must have shape (..., channels, samples)")
    if raw.shape[-1] < MIN_BLOCK_LEN:
        raise RuntimeError("Segment length mismatch")      # segment_raw_signals_200Hz.py:68-69
    lead = raw.shape[:-2]
    n_ch = raw.shape[-2]
    flat = raw.reshape((-1,) + tuple(raw.shape[-2:]))        # view when the leading axes are contiguous
    de, psd, status = ops.de_psd_from_raw(flat, _mode_id(mode))
    if check:
        raise_if_zero_power(status)
    n_win = de.shape[1]
    shape = tuple(lead) + (CONCEPTS_PER_BLOCK, REPS_PER_CONCEPT) + ((n_win,) if n_win > 1 else ()) + (n_ch, 5)
    return de.reshape(shape), psd.reshape(shape)


def de_psd_from_clips(clips, mode="500ms", check=True):
    """DE/PSD of segmented 2 s clips: float32 CUDA (..., ch, 400) -> (..., [W,] ch, 5)."""
    if clips.dim() < 2 or clips.shape[-1] != 400:
        raise ValueError("clips must have shape (..., channels, 400)")
    lead = clips.shape[:-2]
    n_ch = clips.shape[-2]
    flat = clips.reshape((-1, n_ch, 400)).contiguous()
    de, psd, status = ops.de_psd_from_clips(flat, _mode_id(mode))
    if check:
        raise_if_zero_power(status)
    n_win = de.shape[1]
    shape = tuple(lead) + ((n_win,) if n_win > 1 else ()) + (n_ch, 5)
    return de.reshape(shape), psd.reshape(shape)


def de_psd_windows(x, check=True, fre=200):
    """DE/PSD of pre-cut windows: float32 CUDA (..., L) -> (..., 5).  L in {100, 200, 400} at 200 Hz run on the fused
    kernels; any other window length / sampling rate on the general kernel (ops.de_psd_generic)."""
    length = x.shape[-1]
    lead = x.shape[:-1]
    flat = x.reshape(-1, length)
    if flat.stride(-1) != 1:
        flat = flat.contiguous()
    if fre == 200 and length in (100, 200, 400):
        de, psd, status = ops.de_psd_windows(flat)
    else:
        de, psd, status = ops.de_psd_generic(flat, float(fre), int(length))
    if check:
        raise_if_zero_power(status)
    return de.reshape(tuple(lead) + (5,)), psd.reshape(tuple(lead) + (5,))


def segment_clips(raw, fs=200):
    """(..., ch, T) -> (..., 40, 5, ch, 2 fs): what segment_all_files builds, as a bit-exact device gather."""
    lead = raw.shape[:-2]
    flat = raw.reshape((-1,) + tuple(raw.shape[-2:]))
    clips = ops.segment_clips(flat, int(fs))
    return clips.reshape(tuple(lead) + (CONCEPTS_PER_BLOCK, REPS_PER_CONCEPT, raw.shape[-2], 2 * int(fs)))


def sliding_windows(clips, layout="window_major"):
    """500 ms windows every 250 ms of (..., ch, 400) clips, materialised on the device (bit-exact gather):
      layout="window_major": (..., 7, ch, 100) -- seg_sliding_window(data, 0.5, 0.25) (segment_sliding_window.py:19)
      layout="last":         (..., ch, 100, 7) -- the Seq2Seq trainer's inline loop, torch.stack(windows, dim=-1)
                             (EEG2Video_New/Seq2Seq/my_autoregressive_transformer.py:309-314)."""
    lead = clips.shape[:-2]
    n_ch = clips.shape[-2]
    flat = clips.reshape(-1, n_ch, 400).contiguous()
    if layout == "window_major":
        return ops.sliding_windows(flat).reshape(tuple(lead) + (7, n_ch, 100))
    if layout == "last":
        return ops.sliding_windows_last(flat).reshape(tuple(lead) + (n_ch, 100, 7))
    raise ValueError("layout must be 'window_major' or 'last'")
