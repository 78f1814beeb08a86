#!/usr/bin/env python
"""A/B kernel timing of several builds of libeegfe.so in ONE process on ONE GPU (development tool, not the bench).

    python tools/kbench.py [--subjects 24] [--modes 500ms,1s,2s] [--reps 5] lib_a.so lib_b.so ...

Each library is loaded with ctypes; the fused from-raw kernel is timed with CUDA events (20 launches per
measurement, inputs larger than L2), variants interleaved, best and median of `reps` reported.  Also checks that
all variants produce bit-identical features.
"""
import argparse
import ctypes
import statistics
import sys

import torch

MODES = {"500ms": (0, 7), "1s": (1, 2), "2s": (2, 1)}
ALGO_BYTES = {"500ms": 1600 / 7 + 40, "1s": 840.0, "2s": 840.0}


def load(path):
    lib = ctypes.CDLL(path)
    fn = lib.eegfe_de_psd_from_raw
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--subjects", type=int, default=24)
    ap.add_argument("--modes", default="500ms,1s,2s")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--launches", type=int, default=20)
    ap.add_argument("--peak", type=float, default=6541.8)
    ap.add_argument("--block-len", type=int, default=104000,
                    help="samples per channel row; 104001 / 104002 make the rows 4- / 8-byte aligned only (TMA copies of the aligned span, read shifted)")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity to set (32/64/128), 0 = leave")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.init()
    if args.l2_fetch:
        rt = ctypes.CDLL("libcudart.so.12")
        val = ctypes.c_size_t()
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(args.l2_fetch))
        val2 = ctypes.c_size_t()
        rt.cudaDeviceGetLimit(ctypes.byref(val2), 5)
        print(f"cudaLimitMaxL2FetchGranularity {val.value} -> {val2.value} (rc {rc})")
    n_blocks, n_ch, t_len = args.subjects * 7, 62, args.block_len
    g = torch.Generator(device=dev).manual_seed(1)
    raw = torch.randn((n_blocks, n_ch, t_len), device=dev, generator=g) * 30.0
    fns = [(p, load(p)) for p in args.libs]
    stream = torch.cuda.current_stream().cuda_stream
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    for mode in args.modes.split(","):
        mid, nwin = MODES[mode]
        cw = n_blocks * 200 * nwin * n_ch
        outs, times = {}, {p: [] for p, _ in fns}
        for p, fn in fns:
            de = torch.empty((n_blocks * 200, nwin, n_ch, 5), device=dev)
            psd = torch.empty_like(de)
            outs[p] = (de, psd)

        def launch(p, fn):
            de, psd = outs[p]
            rc = fn(raw.data_ptr(), n_blocks, n_ch, t_len, raw.stride(0), raw.stride(1), mid, de.data_ptr(),
                    psd.data_ptr(), status.data_ptr(), stream)
            assert rc == 0, (p, rc)
        for p, fn in fns:
            for _ in range(3):
                launch(p, fn)
        torch.cuda.synchronize()
        for _ in range(args.reps):
            for p, fn in fns:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.launches):
                    launch(p, fn)
                e1.record()
                torch.cuda.synchronize()
                times[p].append(e0.elapsed_time(e1) / args.launches)
        ref = outs[fns[0][0]]
        for p, _ in fns:
            best, med = min(times[p]), statistics.median(times[p])
            same = torch.equal(outs[p][0], ref[0]) and torch.equal(outs[p][1], ref[1])
            print(f"{mode:6s} {p:40s} best {best * 1e3:8.1f} us  {cw / best / 1e6:7.2f} Gcw/s  "
                  f"{cw * ALGO_BYTES[mode] / best / 1e6 / args.peak * 100:5.1f}% hbm | median {cw / med / 1e6:7.2f} Gcw/s"
                  f" | same_as_first={same}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
