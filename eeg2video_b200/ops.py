"""torch.library custom ops over the extern "C" launchers of libeegfe.so.

PyTorch is plumbing here: it owns device memory and streams; all arithmetic happens in the CUDA library.
Every op is CUDA-only -- a CPU tensor is an error, not a fallback.

    eeg2video::de_psd_from_raw(raw, mode)     -> (de, psd, status)
    eeg2video::de_psd_from_clips(clips, mode) -> (de, psd, status)
    eeg2video::de_psd_windows(x)              -> (de, psd, status)
    eeg2video::segment_clips(raw, fs)         -> clips
    eeg2video::sliding_windows(clips)         -> windows

`status` is a 1-element int32 device tensor (see EEGFE_STATUS_* in include/eegfe.h); reading it is the caller's
choice, so that throughput paths never synchronise.
"""
from typing import Tuple

import torch

from . import _lib

WINDOWS_PER_CLIP = {_lib.MODE_500MS: 7, _lib.MODE_1S: 2, _lib.MODE_2S: 1}

_COPY_DTYPES = {
    torch.float32: _lib.DTYPE_F32, torch.float64: _lib.DTYPE_F64,
    torch.float16: _lib.DTYPE_F16, torch.int16: _lib.DTYPE_I16,
}


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"eeg2video ops are CUDA-only (no CPU fallback): `{name}` is on {t.device}")


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


@torch.library.custom_op("eeg2video::de_psd_from_raw", mutates_args=(), device_types="cuda")
def de_psd_from_raw(raw: torch.Tensor, mode: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """raw float32 (n_blocks, n_ch, T), last axis contiguous -> de, psd float32 (n_blocks*200, W, n_ch, 5)."""
    _require_cuda(raw, "raw")
    if raw.dim() != 3 or raw.dtype != torch.float32 or (raw.numel() > 0 and raw.stride(2) != 1):
        raise ValueError("raw must be float32 (n_blocks, n_ch, T) with a contiguous time axis")
    if mode not in WINDOWS_PER_CLIP:
        raise ValueError(f"unknown mode {mode}")
    n_blocks, n_ch, t_len = raw.shape
    n_win = WINDOWS_PER_CLIP[mode]
    lib = _lib.load()
    with torch.cuda.device(raw.device):
        de = torch.empty((n_blocks * 200, n_win, n_ch, 5), dtype=torch.float32, device=raw.device)
        psd = torch.empty_like(de)
        status = torch.zeros(1, dtype=torch.int32, device=raw.device)
        # Rows that are not 16-byte aligned (an odd block length or stride) cannot be the source of a TMA bulk copy:
        # libeegfe then copies the 16-byte aligned span around each row and reads it shifted (separate instantiations of
        # the kernels) -- one launch, no extra pass, no workspace.
        _lib.check(lib.eegfe_de_psd_from_raw(
            raw.data_ptr(), n_blocks, n_ch, t_len, raw.stride(0), raw.stride(1), mode,
            de.data_ptr(), psd.data_ptr(), status.data_ptr(), _stream(raw)))
    return de, psd, status


@de_psd_from_raw.register_fake
def _(raw, mode):
    n_win = WINDOWS_PER_CLIP[mode]
    de = raw.new_empty((raw.shape[0] * 200, n_win, raw.shape[1], 5), dtype=torch.float32)
    return de, torch.empty_like(de), raw.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("eeg2video::de_psd_from_concepts", mutates_args=(), device_types="cuda")
def de_psd_from_concepts(x: torch.Tensor, mode: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x float32 contiguous (n_blocks, n_ch, 40, 2000): the 5 x 400 live samples of every concept, hint periods
    dropped -> de, psd float32 (n_blocks*200, W, n_ch, 5).  Same numbers as de_psd_from_raw on the full recording."""
    _require_cuda(x, "x")
    if x.dim() != 4 or x.shape[2] != 40 or x.shape[3] != 2000 or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("x must be contiguous float32 (n_blocks, n_ch, 40, 2000)")
    if mode not in WINDOWS_PER_CLIP:
        raise ValueError(f"unknown mode {mode}")
    n_blocks, n_ch = x.shape[0], x.shape[1]
    n_win = WINDOWS_PER_CLIP[mode]
    with torch.cuda.device(x.device):
        de = torch.empty((n_blocks * 200, n_win, n_ch, 5), dtype=torch.float32, device=x.device)
        psd = torch.empty_like(de)
        status = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.check(_lib.load().eegfe_de_psd_from_concepts(
            x.data_ptr(), n_blocks, n_ch, n_ch * 80000, 80000, 2000, 0, mode,
            de.data_ptr(), psd.data_ptr(), status.data_ptr(), _stream(x)))
    return de, psd, status


@de_psd_from_concepts.register_fake
def _(x, mode):
    de = x.new_empty((x.shape[0] * 200, WINDOWS_PER_CLIP[mode], x.shape[1], 5), dtype=torch.float32)
    return de, torch.empty_like(de), x.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("eeg2video::de_psd_from_clips", mutates_args=(), device_types="cuda")
def de_psd_from_clips(clips: torch.Tensor, mode: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """clips float32 contiguous (n_clips, n_ch, 400) -> de, psd float32 (n_clips, W, n_ch, 5)."""
    _require_cuda(clips, "clips")
    if clips.dim() != 3 or clips.shape[2] != 400 or clips.dtype != torch.float32 or not clips.is_contiguous():
        raise ValueError("clips must be contiguous float32 (n_clips, n_ch, 400)")
    if mode not in WINDOWS_PER_CLIP:
        raise ValueError(f"unknown mode {mode}")
    n_clips, n_ch, _ = clips.shape
    n_win = WINDOWS_PER_CLIP[mode]
    with torch.cuda.device(clips.device):
        de = torch.empty((n_clips, n_win, n_ch, 5), dtype=torch.float32, device=clips.device)
        psd = torch.empty_like(de)
        status = torch.zeros(1, dtype=torch.int32, device=clips.device)
        _lib.check(_lib.load().eegfe_de_psd_from_clips(
            clips.data_ptr(), n_clips, n_ch, mode, de.data_ptr(), psd.data_ptr(), status.data_ptr(), _stream(clips)))
    return de, psd, status


@de_psd_from_clips.register_fake
def _(clips, mode):
    de = clips.new_empty((clips.shape[0], WINDOWS_PER_CLIP[mode], clips.shape[1], 5), dtype=torch.float32)
    return de, torch.empty_like(de), clips.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("eeg2video::de_psd_windows", mutates_args=(), device_types="cuda")
def de_psd_windows(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x float32 (n_rows, L), L in {100, 200, 400}, unit stride along L -> de, psd float32 (n_rows, 5)."""
    _require_cuda(x, "x")
    if x.dim() != 2 or x.dtype != torch.float32 or x.shape[1] not in (100, 200, 400) or \
            (x.numel() > 0 and x.stride(1) != 1):
        raise ValueError("x must be float32 (n_rows, L) with L in {100, 200, 400} and unit stride along L")
    n_rows, length = x.shape
    row_stride = x.stride(0) if n_rows > 1 else length
    with torch.cuda.device(x.device):
        de = torch.empty((n_rows, 5), dtype=torch.float32, device=x.device)
        psd = torch.empty_like(de)
        status = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.check(_lib.load().eegfe_de_psd_windows(
            x.data_ptr(), n_rows, length, row_stride, de.data_ptr(), psd.data_ptr(), status.data_ptr(), _stream(x)))
    return de, psd, status


@de_psd_windows.register_fake
def _(x):
    de = x.new_empty((x.shape[0], 5), dtype=torch.float32)
    return de, torch.empty_like(de), x.new_empty((1,), dtype=torch.int32)


def band_bins(fre):
    """Inclusive bin ranges of the five bands with the reference's own host expressions (DE_PSD.py:27-29, :35-39, :63):
    fNum = int(f / fre * 200); bins range(fStartNum - 1, fEndNum).  A start of -1 is Python's "last element" (bin 99);
    a bin >= 100 is where the reference raises IndexError (magFFTdata has 100 entries, :59)."""
    lo, hi = [], []
    for f0, f1 in zip((1, 4, 8, 14, 31), (4, 8, 14, 31, 99)):
        lo.append(int(f0 / fre * 200) - 1)
        hi.append(int(f1 / fre * 200) - 1)
    return lo, hi


def hann_weights(length, n_live):
    """First n_live weights of the reference's Hann window of `length` points (DE_PSD.py:51), float64 -> float32."""
    import numpy as np
    n = np.arange(1, n_live + 1, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2 * np.pi * n / (length + 1))).astype(np.float32)


@torch.library.custom_op("eeg2video::de_psd_generic", mutates_args=(), device_types="cuda")
def de_psd_generic(x: torch.Tensor, fre: float, window_points: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x float32 (n_rows, L) with L == window_points, unit stride along L; ANY L >= 1 and sampling rate ->
    de, psd float32 (n_rows, 5).  The general path behind DE_PSD(data, fre, time_window) (DE_PSD.py:8-71)."""
    import ctypes
    _require_cuda(x, "x")
    if x.dim() != 2 or x.dtype != torch.float32 or x.shape[1] != window_points or window_points < 1 or \
            (x.numel() > 0 and x.stride(1) != 1):
        raise ValueError("x must be float32 (n_rows, window_points) with unit stride along the window")
    lo, hi = band_bins(fre)
    for a, b in zip(lo, hi):
        if a < -1:
            raise NotImplementedError(f"fre={fre!r}: band start below bin -1")
        if b > 99:
            raise IndexError(f"index {max(a, 100)} is out of bounds for axis 0 with size 100")     # DE_PSD.py:64
    n_rows, length = x.shape
    n_live = min(length, 200)
    row_stride = x.stride(0) if n_rows > 1 else length
    arr = ctypes.c_int * 5
    with torch.cuda.device(x.device):
        hann = torch.from_numpy(hann_weights(length, n_live)).to(x.device)
        de = torch.empty((n_rows, 5), dtype=torch.float32, device=x.device)
        psd = torch.empty_like(de)
        status = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.check(_lib.load().eegfe_de_psd_generic(
            x.data_ptr(), n_rows, n_live, row_stride, hann.data_ptr(), arr(*lo), arr(*hi), de.data_ptr(),
            psd.data_ptr(), status.data_ptr(), _stream(x)))
    return de, psd, status


@de_psd_generic.register_fake
def _(x, fre, window_points):
    de = x.new_empty((x.shape[0], 5), dtype=torch.float32)
    return de, torch.empty_like(de), x.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("eeg2video::segment_clips", mutates_args=(), device_types="cuda")
def segment_clips(raw: torch.Tensor, fs: int) -> torch.Tensor:
    """raw (n_blocks, n_ch, T) of a 2/4/8-byte dtype -> clips (n_blocks*200, n_ch, 2*fs), bit-exact gather."""
    _require_cuda(raw, "raw")
    if raw.dim() != 3 or raw.dtype not in _COPY_DTYPES or (raw.numel() > 0 and raw.stride(2) != 1):
        raise ValueError("raw must be (n_blocks, n_ch, T) float32/float64/float16/int16 with a contiguous time axis")
    n_blocks, n_ch, t_len = raw.shape
    with torch.cuda.device(raw.device):
        clips = torch.empty((n_blocks * 200, n_ch, 2 * fs), dtype=raw.dtype, device=raw.device)
        _lib.check(_lib.load().eegfe_segment_clips(
            raw.data_ptr(), _COPY_DTYPES[raw.dtype], n_blocks, n_ch, t_len, raw.stride(0), raw.stride(1), fs,
            clips.data_ptr(), _stream(raw)))
    return clips


@segment_clips.register_fake
def _(raw, fs):
    return raw.new_empty((raw.shape[0] * 200, raw.shape[1], 2 * fs))


@torch.library.custom_op("eeg2video::sliding_windows", mutates_args=(), device_types="cuda")
def sliding_windows(clips: torch.Tensor) -> torch.Tensor:
    """clips contiguous (n_clips, n_ch, 400) -> windows (n_clips, 7, n_ch, 100), bit-exact gather."""
    _require_cuda(clips, "clips")
    if clips.dim() != 3 or clips.shape[2] != 400 or clips.dtype not in _COPY_DTYPES or not clips.is_contiguous():
        raise ValueError("clips must be contiguous (n_clips, n_ch, 400) float32/float64/float16/int16")
    n_clips, n_ch, _ = clips.shape
    with torch.cuda.device(clips.device):
        out = torch.empty((n_clips, 7, n_ch, 100), dtype=clips.dtype, device=clips.device)
        _lib.check(_lib.load().eegfe_sliding_windows(
            clips.data_ptr(), _COPY_DTYPES[clips.dtype], n_clips, n_ch, out.data_ptr(), _stream(clips)))
    return out


@sliding_windows.register_fake
def _(clips):
    return clips.new_empty((clips.shape[0], 7, clips.shape[1], 100))


@torch.library.custom_op("eeg2video::sliding_windows_last", mutates_args=(), device_types="cuda")
def sliding_windows_last(clips: torch.Tensor) -> torch.Tensor:
    """clips contiguous (n_clips, n_ch, 400) -> (n_clips, n_ch, 100, 7): the Seq2Seq trainer's window layout
    (my_autoregressive_transformer.py:309-314), bit-exact gather."""
    _require_cuda(clips, "clips")
    if clips.dim() != 3 or clips.shape[2] != 400 or clips.dtype not in _COPY_DTYPES or not clips.is_contiguous():
        raise ValueError("clips must be contiguous (n_clips, n_ch, 400) float32/float64/float16/int16")
    n_clips, n_ch, _ = clips.shape
    with torch.cuda.device(clips.device):
        out = torch.empty((n_clips, n_ch, 100, 7), dtype=clips.dtype, device=clips.device)
        _lib.check(_lib.load().eegfe_sliding_windows_layout(
            clips.data_ptr(), _COPY_DTYPES[clips.dtype], n_clips, n_ch, _lib.WINDOWS_LAST, out.data_ptr(),
            _stream(clips)))
    return out


@sliding_windows_last.register_fake
def _(clips):
    return clips.new_empty((clips.shape[0], clips.shape[1], 100, 7))


@torch.library.custom_op("eeg2video::glmnet_inputs_from_raw", mutates_args=(), device_types="cuda")
def glmnet_inputs_from_raw(raw: torch.Tensor, ch_scale: torch.Tensor, ch_mean: torch.Tensor
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """raw float32 (n_blocks, n_ch, T) + per-channel scale = 1 / std and mean (float32, n_ch) ->
    clips_norm = (x - mean) * scale (n_blocks*200, n_ch, 400), de, psd (n_blocks*200, 7, n_ch, 5), status."""
    _require_cuda(raw, "raw")
    if raw.dim() != 3 or raw.dtype != torch.float32 or (raw.numel() > 0 and raw.stride(2) != 1):
        raise ValueError("raw must be float32 (n_blocks, n_ch, T) with a contiguous time axis")
    n_blocks, n_ch, t_len = raw.shape
    for name, v in (("ch_scale", ch_scale), ("ch_mean", ch_mean)):
        if v.dtype != torch.float32 or v.shape != (n_ch,) or not v.is_contiguous() or v.device != raw.device:
            raise ValueError(f"{name} must be a contiguous float32 ({n_ch},) tensor on {raw.device}")
    with torch.cuda.device(raw.device):
        clips = torch.empty((n_blocks * 200, n_ch, 400), dtype=torch.float32, device=raw.device)
        de = torch.empty((n_blocks * 200, 7, n_ch, 5), dtype=torch.float32, device=raw.device)
        psd = torch.empty_like(de)
        status = torch.zeros(1, dtype=torch.int32, device=raw.device)
        _lib.check(_lib.load().eegfe_glmnet_inputs_from_raw(
            raw.data_ptr(), n_blocks, n_ch, t_len, raw.stride(0), raw.stride(1), ch_scale.data_ptr(),
            ch_mean.data_ptr(), clips.data_ptr(), de.data_ptr(), psd.data_ptr(), status.data_ptr(), _stream(raw)))
    return clips, de, psd, status


@glmnet_inputs_from_raw.register_fake
def _(raw, ch_scale, ch_mean):
    clips = raw.new_empty((raw.shape[0] * 200, raw.shape[1], 400), dtype=torch.float32)
    de = raw.new_empty((raw.shape[0] * 200, 7, raw.shape[1], 5), dtype=torch.float32)
    return clips, de, torch.empty_like(de), raw.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("eeg2video::channel_stats", mutates_args=(), device_types="cuda")
def channel_stats(raw: torch.Tensor, block_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """raw float32 (n_blocks, n_ch, T), block_mask uint8 (n_blocks,) -> per-channel (mean, std) float64 over the
    clip samples of the selected blocks (population std)."""
    _require_cuda(raw, "raw")
    if raw.dim() != 3 or raw.dtype != torch.float32 or (raw.numel() > 0 and raw.stride(2) != 1):
        raise ValueError("raw must be float32 (n_blocks, n_ch, T) with a contiguous time axis")
    n_blocks, n_ch, t_len = raw.shape
    if block_mask.dtype != torch.uint8 or block_mask.shape != (n_blocks,) or block_mask.device != raw.device:
        raise ValueError("block_mask must be uint8 (n_blocks,) on the same device")
    with torch.cuda.device(raw.device):
        work = torch.empty((max(n_blocks * n_ch, 1), 2), dtype=torch.float64, device=raw.device)
        mean = torch.empty(n_ch, dtype=torch.float64, device=raw.device)
        std = torch.empty_like(mean)
        _lib.check(_lib.load().eegfe_channel_stats(
            raw.data_ptr(), n_blocks, n_ch, t_len, raw.stride(0), raw.stride(1),
            block_mask.contiguous().data_ptr(), work.data_ptr(), mean.data_ptr(), std.data_ptr(), _stream(raw)))
    return mean, std


@channel_stats.register_fake
def _(raw, block_mask):
    mean = raw.new_empty((raw.shape[1],), dtype=torch.float64)
    return mean, torch.empty_like(mean)


@torch.library.custom_op("eeg2video::select_units", mutates_args=(), device_types="cuda")
def select_units(feat: torch.Tensor, src_index: torch.Tensor, reduce_windows: bool) -> torch.Tensor:
    """feat float32 (n_units, W, cols), src_index int32 (n_out,) -> (n_out, W, cols), or (n_out, cols) = mean over
    the W windows when `reduce_windows`."""
    _require_cuda(feat, "feat")
    if feat.dim() != 3 or feat.dtype != torch.float32 or not feat.is_contiguous():
        raise ValueError("feat must be contiguous float32 (n_units, n_windows, n_cols)")
    if src_index.dtype != torch.int32 or src_index.dim() != 1 or src_index.device != feat.device:
        raise ValueError("src_index must be an int32 vector on the same device")
    n_units, n_win, cols = feat.shape
    n_out = src_index.shape[0]
    shape = (n_out, cols) if reduce_windows else (n_out, n_win, cols)
    with torch.cuda.device(feat.device):
        out = torch.empty(shape, dtype=torch.float32, device=feat.device)
        _lib.check(_lib.load().eegfe_select_units(feat.data_ptr(), n_units, n_win, cols,
                                                  src_index.contiguous().data_ptr(), n_out, int(bool(reduce_windows)),
                                                  out.data_ptr(), _stream(feat)))
    return out


@select_units.register_fake
def _(feat, src_index, reduce_windows):
    if reduce_windows:
        return feat.new_empty((src_index.shape[0], feat.shape[2]))
    return feat.new_empty((src_index.shape[0], feat.shape[1], feat.shape[2]))


def _as_groups(x, name):
    """(n_rows, n_cols) or (n_groups, n_rows, n_cols) float32 with unit column stride -> (x3d, squeeze)."""
    _require_cuda(x, name)
    if x.dim() not in (2, 3) or x.dtype != torch.float32 or (x.numel() > 0 and x.stride(-1) != 1):
        raise ValueError(f"{name} must be float32 ([n_groups,] n_rows, n_cols) with contiguous columns")
    return (x.unsqueeze(0), True) if x.dim() == 2 else (x, False)


def _strides(x3):
    g, n, c = x3.shape
    return (x3.stride(1) if n > 1 else c), (x3.stride(0) if g > 1 else 0)


@torch.library.custom_op("eeg2video::column_stats", mutates_args=(), device_types="cuda")
def column_stats(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x float32 ([G,] n_rows, n_cols) -> (mean, var, scale) float64 ([G,] n_cols), StandardScaler rules; with a
    leading group axis every group gets its own statistics in the same launches."""
    x3, squeeze = _as_groups(x, "x")
    g, n_rows, n_cols = x3.shape
    row_stride, group_stride = _strides(x3)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        work = torch.empty(int(lib.eegfe_column_stats_workspace(g, n_rows, n_cols)), dtype=torch.float64,
                           device=x.device)
        mean = torch.empty((g, n_cols), dtype=torch.float64, device=x.device)
        var = torch.empty_like(mean)
        scale = torch.empty_like(mean)
        _lib.check(lib.eegfe_column_stats(x3.data_ptr(), g, n_rows, n_cols, row_stride, group_stride,
                                          work.data_ptr(), mean.data_ptr(), var.data_ptr(), scale.data_ptr(),
                                          _stream(x)))
    if squeeze:
        return mean[0], var[0], scale[0]
    return mean, var, scale


@column_stats.register_fake
def _(x):
    m = x.new_empty(tuple(x.shape[:-2]) + (x.shape[-1],), dtype=torch.float64)
    return m, torch.empty_like(m), torch.empty_like(m)


@torch.library.custom_op("eeg2video::standardize", mutates_args=(), device_types="cuda")
def standardize(x: torch.Tensor, mean: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """float32((double(x) - mean) / scale), float32 ([G,] n_rows, n_cols) -> float32 contiguous."""
    x3, squeeze = _as_groups(x, "x")
    g, n_rows, n_cols = x3.shape
    want = (n_cols,) if squeeze else (g, n_cols)
    for name, v in (("mean", mean), ("scale", scale)):
        if v.dtype != torch.float64 or tuple(v.shape) != want or not v.is_contiguous() or v.device != x.device:
            raise ValueError(f"{name} must be a contiguous float64 {want} tensor on {x.device}")
    row_stride, group_stride = _strides(x3)
    with torch.cuda.device(x.device):
        out = torch.empty((g, n_rows, n_cols), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().eegfe_standardize(x3.data_ptr(), g, n_rows, n_cols, row_stride, group_stride,
                                                 mean.data_ptr(), scale.data_ptr(), out.data_ptr(), _stream(x)))
    return out[0] if squeeze else out


@standardize.register_fake
def _(x, mean, scale):
    return x.new_empty(x.shape)


@torch.library.custom_op("eeg2video::de_from_psd", mutates_args=("de",), device_types="cuda")
def de_from_psd_(psd: torch.Tensor, de: torch.Tensor, status: torch.Tensor) -> None:
    """de[...] = log2(100 psd) in place of `de` (same shape, float32, contiguous): bit-identical to the DE the feature
    kernels write next to `psd` (DE_PSD.py:68).  `status` (int32[1]) collects EEGFE_STATUS_ZERO_POWER."""
    _require_cuda(psd, "psd")
    if psd.dtype != torch.float32 or de.dtype != torch.float32 or psd.shape != de.shape or \
            not psd.is_contiguous() or not de.is_contiguous() or de.device != psd.device:
        raise ValueError("psd and de must be contiguous float32 tensors of one shape on one device")
    with torch.cuda.device(psd.device):
        _lib.check(_lib.load().eegfe_de_from_psd(psd.data_ptr(), psd.numel(), de.data_ptr(), status.data_ptr(),
                                                 _stream(psd)))
