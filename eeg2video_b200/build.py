"""Build libeegfe.so (the CUDA kernels + C ABI) in-tree for sm_100a.

    python -m eeg2video_b200.build          # or: __graft_entry__.build()

nvcc cross-compiles without a GPU; the resulting eeg2video_b200/libeegfe.so is git-ignored but travels with the
repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libeegfe.so")
SOURCES = ["eegfe_kernels.cu"]
HEADERS = ["bandpower.cuh", "cplx.cuh", "eegfe_stream.cuh", "eegfe_consumers.cuh", "eegfe_tables.h", os.path.join("..", "..", "include", "eegfe.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: cannot build libeegfe.so")
    return path


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(CSRC, "gen_tables.py")]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile the library if it is missing or stale; returns its path."""
    tables = os.path.join(CSRC, "eegfe_tables.h")
    if not os.path.exists(tables):
        subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_tables.py")])
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) \
        + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libeegfe.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
