#!/usr/bin/env python
"""Time eegfe_de_psd_windows on dense pre-cut windows (development tool)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eeg2video_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--len", type=int, default=100)
    ap.add_argument("--rows", type=int, default=14582400)
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)
    dev = torch.device("cuda:0")
    x = torch.randn((args.rows, args.len), device=dev) * 30
    de = torch.empty((args.rows, 5), device=dev)
    psd = torch.empty_like(de)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(lib.eegfe_de_psd_windows(x.data_ptr(), args.rows, args.len, args.len, de.data_ptr(), psd.data_ptr(),
                                            status.data_ptr(), stream))
    for _ in range(3):
        step()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    b = args.len * 4 + 40
    print(f"windows {args.len}: {best * 1e3:.1f} us  {args.rows / best / 1e6:.2f} Gcw/s  {args.rows * b / best / 1e6:.0f} GB/s "
          f"({args.rows * b / best / 1e6 / 6541.8 * 100:.1f}% of measured HBM peak)")


if __name__ == "__main__":
    main()
