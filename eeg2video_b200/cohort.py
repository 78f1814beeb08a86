"""Cohort-level driver: shard subjects over the GPUs of one box, run the fused kernel per shard, gather the
feature tensors to rank 0.

Every channel-window is independent, so there is NO communication on the hot path: each rank processes its own
contiguous range of subjects out of its own HBM.  The only collective is the final gather of the float32 DE / PSD
tensors (24.3 MB per subject in 500 ms mode) to rank 0 over NCCL (NVLink 5 / NVSwitch); it is chunked so that
it can overlap the next chunk's kernel on a side stream.

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests (which exercise the sharding
and gather logic with a stand-in compute function).
"""
import torch
import torch.distributed as dist

from . import frontend


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of `n_items` for `rank` of `world` (first n_items % world ranks get one
    extra item)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world):
    return [shard_bounds(n_items, r, world)[1] - shard_bounds(n_items, r, world)[0] for r in range(world)]


def gather_to_rank0(local, n_total, group=None, dst=0):
    """Gather per-rank tensors (leading axis = that rank's subjects, in shard order) into one tensor of
    `n_total` leading entries on rank `dst`.  Returns the full tensor on `dst`, None elsewhere.
    Uneven shards are handled (sizes follow shard_bounds)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: expected {sizes[rank]} leading entries, got {local.shape[0]}")
    local = local.contiguous()
    if rank == dst:
        full = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        parts = list(full.split(sizes, dim=0))
        if all(s == sizes[0] for s in sizes):
            dist.gather(local, gather_list=parts, dst=dst, group=group)
        else:                                   # ragged: point-to-point from each rank
            parts[dst].copy_(local)
            reqs = [dist.irecv(parts[r], src=r, group=group) for r in range(world) if r != dst and sizes[r] > 0]
            for q in reqs:
                q.wait()
        return full
    if all(s == sizes[0] for s in sizes):
        dist.gather(local, gather_list=None, dst=dst, group=group)
    elif sizes[rank] > 0:
        dist.send(local, dst=dst, group=group)
    return None


def process_shard(raw, mode="500ms", chunk_subjects=None, compute=None):
    """Run the fused kernel over this rank's subjects.

    raw: (n_local_subjects, 7, 62, T) float32 on this rank's GPU.  Returns (de, psd) with the subject axis leading.
    `compute` (test hook) replaces frontend.de_psd_from_raw with another callable of the same contract.
    """
    fn = compute or (lambda x: frontend.de_psd_from_raw(x, mode, check=False))
    n = raw.shape[0]
    if n == 0:
        return None, None
    step = chunk_subjects or n
    des, psds = [], []
    for lo in range(0, n, step):
        de, psd = fn(raw[lo:lo + step])
        des.append(de)
        psds.append(psd)
    return (des[0], psds[0]) if len(des) == 1 else (torch.cat(des), torch.cat(psds))


def process_cohort(raw_local, n_subjects_total, mode="500ms", chunk_subjects=None, group=None, compute=None,
                   overlap=False):
    """Shard-local compute + gather to rank 0.  Returns (de, psd) for the whole cohort on rank 0, (None, None)
    elsewhere.

    overlap=True (needs chunk_subjects and equal shards): the gather of chunk i is issued as soon as its kernel has
    been enqueued and runs on the communication stream while the compute stream works on chunk i + 1, so only the
    last chunk's transfer is exposed (the gather is ~5x the kernel time at 8 GPUs, rank 0's NVLink ingest being the
    limit, so what overlap hides is the compute, not the transfer).  Measured on 2 B200s with 24 subjects per GPU in
    chunks of 6: 7.5 ms against 3.8 ms for kernel + one big gather -- eight small gathers and per-chunk launches cost
    more than the 1.2 ms of compute they hide -- so it stays off by default and pays only when a chunk's kernel time
    is several milliseconds (hundreds of subjects per GPU).
    """
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if overlap and distributed and chunk_subjects:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        sizes = shard_sizes(n_subjects_total, world)
        if all(s == sizes[0] for s in sizes) and sizes[0] > 0:
            return _process_cohort_overlapped(raw_local, sizes[0], world, rank, mode, chunk_subjects, group, compute)
    de, psd = process_shard(raw_local, mode, chunk_subjects, compute)
    de_all = gather_to_rank0(de, n_subjects_total, group)
    psd_all = gather_to_rank0(psd, n_subjects_total, group)
    return de_all, psd_all


def _process_cohort_overlapped(raw_local, n_local, world, rank, mode, chunk, group, compute):
    fn = compute or (lambda x: frontend.de_psd_from_raw(x, mode, check=False))
    full = [None, None]
    works, keep = [], []
    for lo in range(0, n_local, chunk):
        hi = min(lo + chunk, n_local)
        outs = fn(raw_local[lo:hi])                          # enqueued on the current (compute) stream
        for k, part in enumerate(outs):
            part = part.contiguous()
            if rank == 0:
                if full[k] is None:
                    full[k] = torch.empty((world * n_local,) + tuple(part.shape[1:]), dtype=part.dtype,
                                          device=part.device)
                dest = [full[k][r * n_local + lo:r * n_local + hi] for r in range(world)]
                works.append(dist.gather(part, gather_list=dest, dst=0, group=group, async_op=True))
            else:
                works.append(dist.gather(part, gather_list=None, dst=0, group=group, async_op=True))
            keep.append(part)                                # alive until its transfer has finished
    for w in works:
        w.wait()
    return (full[0], full[1]) if rank == 0 else (None, None)
