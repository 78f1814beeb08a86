#!/usr/bin/env python
"""Time the kernels of the consumer-side build one by one (development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eeg2video_b200 import _lib, consumers, ops  # noqa: E402


def timed(fn, reps=50):
    for _ in range(5):
        out = fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, out


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    variants = sys.argv[2:]                 # other builds of the library (tools/build_variant.sh) to A/B in this process
    dev = torch.device("cuda:0")
    feat = torch.randn((S * 1400, 2, 310), device=dev) * 3 + 20
    gt = np.stack([np.random.default_rng(0).permutation(40) + 1 for _ in range(7)])
    one = consumers.clip_index(range(6), gt, range(1, 41))
    idx = torch.from_numpy(np.concatenate([one + s * 1400 for s in range(S)]).astype(np.int32)).to(dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    us, x = timed(lambda: ops.select_units(feat, idx, True))
    print(f"select_units          {us:8.1f} us")
    x = x.reshape(S, 1200, 310)
    us, (m, v, sc) = timed(lambda: ops.column_stats(x))
    print(f"column_stats (4 krn)  {us:8.1f} us")
    us, out = timed(lambda: ops.standardize(x, m, sc))
    print(f"standardize           {us:8.1f} us")
    # the statistics kernels one by one through the C ABI
    work = torch.empty(2 * int(lib.eegfe_column_stats_workspace(S, 1200, 310)), dtype=torch.float64, device=dev)
    mean = torch.empty((S, 310), dtype=torch.float64, device=dev)
    var, scale = torch.empty_like(mean), torch.empty_like(mean)
    us, _ = timed(lambda: lib.eegfe_column_stats(x.data_ptr(), S, 1200, 310, 310, 1200 * 310, work.data_ptr(),
                                                  mean.data_ptr(), var.data_ptr(), scale.data_ptr(), stream))
    print(f"column_stats via ABI  {us:8.1f} us (no torch allocations)")
    # each step captured alone as a CUDA graph of 20 repetitions: device time per launch without host overhead
    y = torch.empty((S * 1200, 310), device=dev)
    o2 = torch.empty((S, 1200, 310), device=dev)
    steps = {
        "select_units": lambda st: lib.eegfe_select_units(feat.data_ptr(), S * 1400, 2, 310, idx.data_ptr(), S * 1200, 1,
                                                          y.data_ptr(), st),
        "column_stats": lambda st: lib.eegfe_column_stats(x.data_ptr(), S, 1200, 310, 310, 1200 * 310, work.data_ptr(),
                                                          mean.data_ptr(), var.data_ptr(), scale.data_ptr(), st),
        "standardize": lambda st: lib.eegfe_standardize(x.data_ptr(), S, 1200, 310, 310, 1200 * 310, mean.data_ptr(),
                                                        scale.data_ptr(), o2.data_ptr(), st),
    }
    steps["all three"] = lambda st: [steps[k](st) for k in ("select_units", "column_stats", "standardize")]
    import ctypes
    for path in variants:
        v = ctypes.CDLL(os.path.abspath(path))
        for fn_name in ("eegfe_select_units", "eegfe_column_stats", "eegfe_standardize"):
            getattr(v, fn_name).restype, getattr(v, fn_name).argtypes = _lib.SIGNATURES[fn_name]
        tag = os.path.basename(path)
        sel = lambda st, v=v: v.eegfe_select_units(feat.data_ptr(), S * 1400, 2, 310, idx.data_ptr(), S * 1200, 1,
                                                   y.data_ptr(), st)
        sta = lambda st, v=v: v.eegfe_column_stats(x.data_ptr(), S, 1200, 310, 310, 1200 * 310, work.data_ptr(),
                                                   mean.data_ptr(), var.data_ptr(), scale.data_ptr(), st)
        std = lambda st, v=v: v.eegfe_standardize(x.data_ptr(), S, 1200, 310, 310, 1200 * 310, mean.data_ptr(),
                                                  scale.data_ptr(), o2.data_ptr(), st)
        steps[f"{tag} select"] = sel
        steps[f"{tag} stats"] = sta
        steps[f"{tag} standardize"] = std
        steps[f"{tag} all three"] = lambda st, a=sel, b=sta, c=std: [a(st), b(st), c(st)]
    for name, fn in steps.items():
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn(side.cuda_stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                fn(side.cuda_stream)
        us, _ = timed(g.replay, reps=10)
        print(f"graph of 20 x {name:28s} {us / 20:8.1f} us each")


if __name__ == "__main__":
    main()
