"""The C-ABI shared library: it loads, exports every function include/eegfe.h declares, and answers the calls
that need no GPU.  CPU only (no kernel is launched)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from eeg2video_b200 import _lib

HEADER = os.path.join(ROOT, "include", "eegfe.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"^\s*(?:const\s+char\*|int64_t|int)\s+(eegfe_\w+)\s*\(", text, flags=re.M)


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run python -m eeg2video_b200.build"
    assert os.path.dirname(_lib.LIB_PATH) == os.path.join(ROOT, "eeg2video_b200")


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 10
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(dll, name), f"{name} declared in eegfe.h but not exported"
    assert sorted(names) == sorted(_lib.SIGNATURES), "ctypes binding table out of sync with eegfe.h"


def test_header_argument_counts_match_binding():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(name + r"\s*\(([^)]*)\)", text)
        assert m, name
        args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == len(argtypes), name


def test_no_gpu_calls():
    lib = _lib.load()
    assert lib.eegfe_abi_version() == 1
    assert [lib.eegfe_windows_per_clip(m) for m in (0, 1, 2)] == [7, 2, 1]
    assert lib.eegfe_windows_per_clip(9) == _lib.EINVAL
    assert lib.eegfe_error_string(0) == b"success"
    assert lib.eegfe_error_string(_lib.ERANGE) == b"Segment length mismatch"
    # argument validation happens before anything touches the device
    assert lib.eegfe_de_psd_from_raw(None, 0, 62, 104000, 0, 104000, 0, None, None, None, None) == 0      # empty
    assert lib.eegfe_de_psd_from_raw(None, 1, 62, 104000, 62 * 104000, 104000, 0, None, None, None, None) == _lib.EINVAL
    assert lib.eegfe_de_psd_from_raw(None, 1, 62, 104000, 62 * 104000, 104000, 7, None, None, None, None) == _lib.EINVAL
    assert lib.eegfe_de_psd_from_clips(None, 0, 62, 1, None, None, None, None) == 0
    assert lib.eegfe_de_psd_windows(None, 0, 100, 100, None, None, None, None) == 0
    assert lib.eegfe_de_psd_windows(None, 4, 150, 150, None, None, None, None) == _lib.EINVAL
    assert lib.eegfe_segment_clips(None, 99, 1, 62, 104000, 0, 104000, 200, None, None) == _lib.EDTYPE
    assert lib.eegfe_sliding_windows(None, 0, 0, 62, None, None) == 0
    with pytest.raises(_lib.EegfeError, match="Segment length mismatch"):
        _lib.check(_lib.ERANGE)
    assert _lib.launch_count() == 0


def test_block_too_short_is_erange():
    lib = _lib.load()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.eegfe_de_psd_from_raw(p, 1, 62, 103999, 62 * 103999, 103999, 0, p, p, None, None) == _lib.ERANGE
    assert lib.eegfe_segment_clips(p, 0, 1, 62, 103999, 62 * 103999, 103999, 200, p, None) == _lib.ERANGE


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under eeg2video_b200/ may reference it."""
    pkg = os.path.join(ROOT, "eeg2video_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "hostemu" not in text or f == "cplx.cuh", f


def test_hot_kernels_stay_near_the_instruction_cache_size():
    """B200's L1.5 instruction cache holds 32 KB; a fully unrolled FFT kernel of 37-48 KB streamed its instructions
    from L2 and lost 15 %.  Tripwire: every feature kernel's SASS body (16 bytes per
    instruction, including its rarely executed inlined tile-duty code) stays below 36 KB.  Out-of-line callees sit
    behind the body -- the body ends where the first CALL target begins.  Exempt: the ring kernel's SHIFT
    instantiations (rows that are not 16-byte aligned, read with one or two loads per sample pair: up to 37 KB)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    listing = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    sizes, callees, name = {}, {}, None
    for line in listing.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            sizes[name] = 0
            callees[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            sizes[name] += 1
            c = re.search(r"CALL\.REL\.NOINC 0x([0-9a-f]+)", line)
            if c:
                callees[name].append(int(c.group(1), 16) // 16)
    sizes = {k: min([v] + callees[k]) for k, v in sizes.items()}
    hot = {k: v * 16 for k, v in sizes.items() if "de_psd_kernel" in k or "de_psd_stream_kernel" in k}
    assert len(hot) >= 5
    shifted = re.compile(r"de_psd_kernelINS_3CfgI.*EELi[12]ELb0EEEvNS_3JobE$")       # de_psd_kernel<Cfg, SHIFT = 1 | 2>
    assert any(shifted.search(k) for k in hot)
    too_big = {k: v for k, v in hot.items() if v > 36 * 1024 and not shifted.search(k)}
    assert all(v <= 40 * 1024 for v in hot.values()), hot
    assert not too_big, too_big


def test_no_warp_diverges_in_front_of_the_block_barrier():
    """Seen twice on B200: an instantiation of the streaming kernel in which ptxas emitted no BSSY / BSYNC pair around the
    set-up `if (tid == 0)` / `if (tid < 112)` hung on every input -- the lane groups of the split warps are not
    reconverged by BAR.SYNC and then run code that keeps per-warp state in uniform registers as independent groups
    (DESIGN.md 4.1).  The set-up is now written with whole-warp conditions only.  Tripwire on the SASS: in front of the
    first BAR.SYNC of every feature kernel, predicated branches outside a BSSY region are at most the known whole-warp
    ones (tile guard, `tid < kMetaThreads` / `tid < kWorkers`, `tid < 32`, the GLMNet table guard)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    listing = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    kernels, name = {}, None
    for line in listing.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            kernels[name].append(line)
    checked = 0
    for k, body in kernels.items():
        if "de_psd_kernel" not in k and "de_psd_stream_kernel" not in k:
            continue
        depth, bare = 0, 0
        for line in body:
            if "BAR.SYNC" in line:
                break
            if "BSSY" in line:
                depth += 1
            elif "BSYNC" in line:
                depth -= 1
            elif re.search(r"@!?P\d\s+BRA\b", line) and depth == 0:
                bare += 1
        else:
            pytest.fail(f"{k}: no block barrier found")
        assert bare <= 4, (k, bare)
        checked += 1
    assert checked >= 10
