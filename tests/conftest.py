import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = os.environ.get("EEG2VIDEO_REFERENCE", "/root/reference")
SCALE = np.float32(2.0 ** -5)          # int16 code -> float32 signal (tests/golden/make_golden.py)

# parity bars from BASELINE.json north_star: PSD within 1e-4 relative, DE within 1e-4 absolute (log2 domain)
PSD_RTOL = 1e-4
DE_ATOL = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def decode(codes):
    return codes.astype(np.float32) * SCALE


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def reference():
    """The live reference package (build container only)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "EEG_preprocessing")):
        pytest.skip("reference checkout not present")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    # the reference package shares its name with our mirror's sub-package; import it explicitly by path order
    mods = {}
    for name in ("DE_PSD", "segment_raw_signals_200Hz", "segment_sliding_window",
                 "extract_DE_PSD_features_1per2s", "extract_DE_PSD_features_1per500ms"):
        mods[name] = importlib.import_module("EEG_preprocessing." + name)
    return mods


@pytest.fixture(scope="session")
def hostemu():
    """Host build of the kernel's arithmetic core (tests/hostemu/hostemu.cpp) -- test infrastructure only."""
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostemu", "_build")
    lib = os.path.join(out_dir, "libeegfe_hostemu.so")
    csrc = os.path.join(ROOT, "eeg2video_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("bandpower.cuh", "cplx.cuh", "eegfe_tables.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", lib, src])
    dll = ctypes.CDLL(lib)
    dll.hostemu_band_energy.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    dll.hostemu_band_energy.restype = ctypes.c_int

    def band_features(x):
        """x (n, L) float32 -> (de, psd) float64 computed from the emulated fp32 band energies."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        e = np.zeros((x.shape[0], 5), dtype=np.float32)
        rc = dll.hostemu_band_energy(x.ctypes.data, x.shape[0], x.shape[1], e.ctypes.data)
        assert rc == 0
        # the device epilogue multiplies by the float32 reciprocal of the bin count (inv_count() in eegfe_kernels.cu)
        psd = e * (np.float32(1.0) / np.array([4, 5, 7, 18, 69], dtype=np.float32))
        with np.errstate(divide="ignore"):
            de = np.log2(np.float32(100.0) * psd.astype(np.float32)).astype(np.float64)
        return de, psd.astype(np.float64)

    def band_energy(x, shift=None, vec=None):
        """raw float32 band energies (n, 5); with shift / vec: the shifted-span read path of the kernels."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        e = np.zeros((x.shape[0], 5), dtype=np.float32)
        if shift is None:
            rc = dll.hostemu_band_energy(x.ctypes.data, x.shape[0], x.shape[1], e.ctypes.data)
        else:
            rc = dll.hostemu_band_energy_shifted(x.ctypes.data, x.shape[0], x.shape[1], shift, vec, e.ctypes.data)
        assert rc == 0
        return e
    dll.hostemu_band_energy_shifted.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p]
    dll.hostemu_band_energy_shifted.restype = ctypes.c_int
    band_features.band_energy = band_energy
    return band_features


def assert_features_close(de, psd, de_ref, psd_ref):
    de, psd, de_ref, psd_ref = (np.asarray(a, dtype=np.float64) for a in (de, psd, de_ref, psd_ref))
    assert de.shape == de_ref.shape and psd.shape == psd_ref.shape
    rel = np.abs(psd - psd_ref) / psd_ref
    assert rel.max() <= PSD_RTOL, f"PSD max relative error {rel.max():.3e} > {PSD_RTOL}"
    err = np.abs(de - de_ref)
    assert err.max() <= DE_ATOL, f"DE max absolute error {err.max():.3e} > {DE_ATOL}"
