"""GLMNet input build (BASELINE.json configs[4]) -- a "next" row beyond the core path.

The reference describes the step (README.md:80-81, :88, :97-99): GLMNet consumes (i) 2 s raw clips, normalised per
channel with the training-split mean / std, shaped (N, 1, 62, 400) for the ShallowNet / glfnet branch
(EEG-VP/models.py:119, :364), and (ii) 500 ms DE/PSD features (N, 7, 62, 5).  Its trainer / inference scripts are not
in the reference tree, so there is no code to pin against; oracle/glmnet_inputs.py states the arithmetic used here.

On the GPU both products come out of ONE pass over the raw recording: the clip rows staged in shared memory for the
FFT leave again as normalised clips, (x - mean) * (1 / std) in float32 (kernel template flag NORM of
eegfe::de_psd_stream_kernel).
"""
import torch

from . import frontend, ops


def channel_stats(raw, train_blocks=None):
    """Per-channel mean and population std of the clip samples over the selected blocks.

    raw: float32 CUDA (..., 62, T) (leading axes are flattened to blocks); train_blocks: boolean mask / index list
    over the flattened blocks (None = all).  Returns (mean, std) float64 CUDA tensors of shape (62,).
    """
    flat = raw.reshape((-1,) + tuple(raw.shape[-2:]))
    mask = torch.ones(flat.shape[0], dtype=torch.uint8, device=flat.device)
    if train_blocks is not None:
        sel = torch.as_tensor(train_blocks, device=flat.device)
        if sel.dtype == torch.bool:
            mask = sel.to(torch.uint8)
        else:
            mask.zero_()
            mask[sel.long()] = 1
    return ops.channel_stats(flat, mask)


def build_inputs(raw, mean, std, check=True):
    """raw (..., 62, T) float32 CUDA, mean / std (62,) -> (clips, de, psd):
    clips (..., 40, 5, 1, 62, 400) normalised raw clips, de / psd (..., 40, 5, 7, 62, 5) 500 ms features."""
    lead = raw.shape[:-2]
    n_ch = raw.shape[-2]
    flat = raw.reshape((-1,) + tuple(raw.shape[-2:]))
    mean64 = torch.as_tensor(mean, dtype=torch.float64, device=flat.device)
    std64 = torch.as_tensor(std, dtype=torch.float64, device=flat.device)
    # a constant channel (std == 0) keeps scale 1, as scikit-learn's scalers do (sklearn _handle_zeros_in_scale), instead
    # of turning the whole channel into inf / NaN
    safe_std = torch.where(std64 == 0, torch.ones_like(std64), std64)
    scale = (1.0 / safe_std).to(torch.float32).contiguous()
    clips, de, psd, status = ops.glmnet_inputs_from_raw(flat, scale, mean64.to(torch.float32).contiguous())
    if check:
        frontend.raise_if_zero_power(status)
    return (clips.reshape(tuple(lead) + (40, 5, 1, n_ch, 400)),
            de.reshape(tuple(lead) + (40, 5, 7, n_ch, 5)), psd.reshape(tuple(lead) + (40, 5, 7, n_ch, 5)))
