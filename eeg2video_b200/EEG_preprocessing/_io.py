"""Host <-> device plumbing shared by the reference-named shims."""
import numpy as np
import torch

FS_SUPPORTED = 200


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("eeg2video_b200 needs a CUDA device: the EEG feature front end has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_torch(x):
    return isinstance(x, torch.Tensor)


def to_device_f32(x):
    """numpy array or torch tensor (any real dtype, any strides) -> float32 CUDA tensor.

    The reference promotes every input to float64 by multiplying with its float64 Hann window (DE_PSD.py:57);
    the kernel computes in float32, which is what the float32 recordings carry anyway.
    """
    if is_torch(x):
        if not x.is_cuda:
            x = x.to(device())
        return x if x.dtype == torch.float32 else x.to(torch.float32)
    arr = np.asarray(x)
    if arr.dtype == object or not (np.issubdtype(arr.dtype, np.floating) or np.issubdtype(arr.dtype, np.integer)
                                   or arr.dtype == np.bool_):
        raise TypeError(f"unsupported dtype {arr.dtype}")
    if arr.dtype != np.float32:
        arr = arr.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device())


def check_fs(fs, what="fs"):
    if int(fs) != FS_SUPPORTED or fs != int(fs):
        raise NotImplementedError(
            f"{what}={fs!r}: the CUDA front end is built for 200 Hz recordings (the only rate the reference's "
            "drivers use; its FFT length is hard-coded to 200, DE_PSD.py:27)")


def finish(tensors, like_torch, dtype):
    """Cast results to the reference's dtype and return them in the caller's array family."""
    if like_torch:
        tdtype = torch.float64 if dtype == np.float64 else torch.float32
        return tuple(t.to(tdtype) for t in tensors)
    return tuple(t.cpu().numpy().astype(dtype, copy=False) for t in tensors)
