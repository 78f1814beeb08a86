"""ctypes binding of libeegfe.so (C ABI: include/eegfe.h).  There is no fallback: if the library is missing the
import of anything that computes fails loudly."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeegfe.so")

MODE_500MS, MODE_1S, MODE_2S = 0, 1, 2
STATUS_ZERO_POWER = 1
DTYPE_F32, DTYPE_F64, DTYPE_F16, DTYPE_I16 = 0, 1, 2, 3
EINVAL, ERANGE, EDTYPE = -1, -2, -3
WINDOWS_WINDOW_MAJOR, WINDOWS_LAST = 0, 1

_i64, _int, _ptr = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p

# name -> (restype, argtypes); kept in one place so tests can check it against include/eegfe.h
SIGNATURES = {
    "eegfe_abi_version": (_int, []),
    "eegfe_error_string": (ctypes.c_char_p, [_int]),
    "eegfe_windows_per_clip": (_int, [_int]),
    "eegfe_de_psd_from_raw": (_int, [_ptr, _i64, _int, _i64, _i64, _i64, _int, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_de_psd_from_concepts": (_int, [_ptr, _i64, _int, _i64, _i64, _i64, _i64, _int, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_glmnet_inputs_from_raw": (_int, [_ptr, _i64, _int, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                            _ptr]),
    "eegfe_channel_stats": (_int, [_ptr, _i64, _int, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_de_psd_generic": (_int, [_ptr, _i64, _int, _i64, _ptr, ctypes.POINTER(_int), ctypes.POINTER(_int), _ptr, _ptr,
                                    _ptr, _ptr]),
    "eegfe_de_from_psd": (_int, [_ptr, _i64, _ptr, _ptr, _ptr]),
    "eegfe_copy2d_async": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _int, _ptr]),
    "eegfe_de_psd_from_clips": (_int, [_ptr, _i64, _int, _int, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_de_psd_windows": (_int, [_ptr, _i64, _int, _i64, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_segment_clips": (_int, [_ptr, _int, _i64, _int, _i64, _i64, _i64, _int, _ptr, _ptr]),
    "eegfe_sliding_windows": (_int, [_ptr, _int, _i64, _int, _ptr, _ptr]),
    "eegfe_sliding_windows_layout": (_int, [_ptr, _int, _i64, _int, _int, _ptr, _ptr]),
    "eegfe_select_units": (_int, [_ptr, _i64, _int, _int, _ptr, _i64, _int, _ptr, _ptr]),
    "eegfe_column_stats_workspace": (_i64, [_i64, _i64, _int]),
    "eegfe_column_stats": (_int, [_ptr, _i64, _i64, _int, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_standardize": (_int, [_ptr, _i64, _i64, _int, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "eegfe_launch_geometry": (_int, [_int, ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int),
                                     ctypes.POINTER(_int)]),
    "eegfe_launch_count": (_i64, []),
    "eegfe_tma_launch_count": (_i64, []),
    "eegfe_set_tensor_loads": (_int, [_int]),
    "eegfe_set_cta_limit": (_int, [_int]),
}

_lib = None


class EegfeError(RuntimeError):
    """Non-zero return code from libeegfe."""

    def __init__(self, code, text):
        super().__init__(f"libeegfe error {code}: {text}")
        self.code = code
        self.text = text


def load():
    """Load libeegfe.so once; raises ImportError with build instructions when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built.  Run `python -m eeg2video_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback for the EEG feature front end.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.eegfe_abi_version() != 1:
        raise ImportError("libeegfe.so ABI version mismatch; rebuild with `python -m eeg2video_b200.build --force`")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        text = load().eegfe_error_string(code).decode()
        raise EegfeError(code, text)


def launch_count():
    return int(load().eegfe_launch_count())


def tma_launch_count():
    return int(load().eegfe_tma_launch_count())


def set_cta_limit(max_ctas):
    """Cap the persistent kernels' grid (0 = one CTA per SM); returns the previous cap."""
    return int(load().eegfe_set_cta_limit(int(max_ctas)))


def set_tensor_loads(on):
    """Switch the 200-sample-row kernels to one TMA tensor copy per tile (measurement option); returns the old value."""
    return bool(load().eegfe_set_tensor_loads(int(bool(on))))


def launch_geometry(mode):
    g, b, s, r = _int(), _int(), _int(), _int()
    check(load().eegfe_launch_geometry(mode, ctypes.byref(g), ctypes.byref(b), ctypes.byref(s), ctypes.byref(r)))
    return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "rows_per_tile": r.value}
