#!/usr/bin/env python
"""Compute + gather time of cohort.process_cohort against the number of kernel chunks per rank (development tool).

    torchrun --nproc-per-node 8 tools/gather_sweep.py [--subjects 24] [--chunks 1,2,3,4,6,8]

Destination tensors preallocated, kernels through the C ABI into preallocated per-rank buffers, CUDA events, max over
ranks -- the same method as bench.py's `gather` object.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eeg2video_b200 import _lib, cohort, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=24)
    ap.add_argument("--chunks", default="1,2,3,4,6,8")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--reserved", default="32:16", help="comma list of dst:src SMs left to NCCL (cohort.RESERVED_SMS_*)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    S = args.subjects
    raw = synth.synth_cohort(range(rank * S, rank * S + S), dev)
    shape = (S * world, 7, 40, 5, 7, 62, 5)
    out = (torch.empty(shape, device=dev), torch.empty(shape, device=dev)) if rank == 0 else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if rank == 0:
        mine = tuple(t[:S].reshape(S, 1400, 7, 62, 5) for t in out)
    else:
        mine = (torch.empty((S, 1400, 7, 62, 5), device=dev), torch.empty((S, 1400, 7, 62, 5), device=dev))
    at = [0]

    def kern(x):
        n = x.shape[0]
        flat = x.reshape(n * 7, 62, x.shape[-1])
        de, psd = mine[0][at[0]:at[0] + n], mine[1][at[0]:at[0] + n]
        _lib.check(lib.eegfe_de_psd_from_raw(flat.data_ptr(), flat.shape[0], 62, flat.shape[2], flat.stride(0),
                                             flat.stride(1), 0, de.data_ptr(), psd.data_ptr(), status.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream))
        at[0] += n
        return de.reshape(n, 7, 40, 5, 7, 62, 5), psd.reshape(n, 7, 40, 5, 7, 62, 5)

    for reserved in args.reserved.split(","):
      cohort.RESERVED_SMS_DST, cohort.RESERVED_SMS_SRC = (int(v) for v in reserved.split(":"))
      for n_chunks in [int(c) for c in args.chunks.split(",")]:
          chunk = max(1, -(-S // n_chunks))
          times = []
          for i in range(1 + args.reps):
              dist.barrier()
              torch.cuda.synchronize()
              a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              at[0] = 0
              a.record()
              cohort.process_cohort(raw, S * world, chunk_subjects=chunk, compute=kern, out=out)
              b.record()
              torch.cuda.synchronize()
              t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
              dist.all_reduce(t, op=dist.ReduceOp.MAX)
              if i:
                  times.append(float(t.item()))
          if rank == 0:
              cw = world * S * 607600
              best, med = min(times), sorted(times)[len(times) // 2]
              print(f"chunks/rank {n_chunks:2d} (chunk {chunk:2d} subjects): best {best:6.3f} ms  median {med:6.3f} ms  "
                    f"-> {cw / med / 1e6:6.2f} G cw/s with gather   SMs left to NCCL dst:src = {reserved}",
                    flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
