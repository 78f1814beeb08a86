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
    if arr.dtype in (np.float64, np.float16, np.int16, np.int32, np.uint8, np.int8, np.int64):
        # upload in the native type and round to float32 on the device (same round-to-nearest result as numpy's
        # astype, without a host pass over the recording)
        if not arr.flags.writeable:
            arr = arr.copy()
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device()).to(torch.float32)
    if arr.dtype != np.float32:
        arr = arr.astype(np.float32)
    if not arr.flags.writeable:
        arr = arr.copy()
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


def convert_directory(in_dir, out_dirs, convert, names=None, accept=None, log=print):
    """The file loop every reference script repeats: for each ``sub{N}.npy`` in ``in_dir`` (or the explicit
    ``names``), load it, run ``convert`` and save result k under ``out_dirs[k]`` with the same file name.
    ``accept(array)`` may veto a file (returns a reason string).  Returns the names written."""
    import os
    for d in out_dirs:
        os.makedirs(d, exist_ok=True)
    if names is None:
        names = sorted(n for n in os.listdir(in_dir) if n.endswith(".npy"))
    written = []
    for name in names:
        array = np.load(os.path.join(in_dir, name))
        reason = accept(array) if accept is not None else None
        if reason:
            log(f"Skipping {name}: {reason}")
            continue
        results = convert(array)
        if not isinstance(results, tuple):
            results = (results,)
        for d, result in zip(out_dirs, results):
            np.save(os.path.join(d, name), result)
        log(f"{name}: " + ", ".join(f"{os.path.join(d, name)} {tuple(r.shape)}" for d, r in zip(out_dirs, results)))
        written.append(name)
    return written
