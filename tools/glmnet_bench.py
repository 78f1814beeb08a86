#!/usr/bin/env python
"""Time eegfe_glmnet_inputs_from_raw alone (development tool; bench.py reports the same under next_rows)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eeg2video_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=24)
    ap.add_argument("--launches", type=int, default=20)
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)
    dev = torch.device("cuda:0")
    nb = args.subjects * 7
    raw = torch.randn((nb, 62, 104000), device=dev) * 30
    scale = torch.full((62,), 1 / 30.0, device=dev)
    shift = torch.zeros(62, device=dev)
    lib = _lib.load()
    clips = torch.empty((nb * 200, 62, 400), device=dev)
    de = torch.empty((nb * 200, 7, 62, 5), device=dev)
    psd = torch.empty_like(de)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(lib.eegfe_glmnet_inputs_from_raw(raw.data_ptr(), nb, 62, 104000, raw.stride(0), raw.stride(1),
                                                    scale.data_ptr(), shift.data_ptr(), clips.data_ptr(), de.data_ptr(),
                                                    psd.data_ptr(), status.data_ptr(), stream))
    for _ in range(3):
        step()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.launches):
            step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / args.launches)
    rows = nb * 200 * 62
    print(f"glmnet inputs: {best * 1e3:.1f} us  {rows * 7 / best / 1e6:.2f} Gcw/s  {rows * 3480 / best / 1e6:.0f} GB/s "
          f"({rows * 3480 / best / 1e6 / 6541.8 * 100:.1f}% of measured HBM peak)")


if __name__ == "__main__":
    main()
