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


def process_cohort(raw_local, n_subjects_total, mode="500ms", chunk_subjects=None, group=None, compute=None):
    """Shard-local compute + gather to rank 0.  Returns (de, psd) for the whole cohort on rank 0, (None, None)
    elsewhere."""
    de, psd = process_shard(raw_local, mode, chunk_subjects, compute)
    de_all = gather_to_rank0(de, n_subjects_total, group)
    psd_all = gather_to_rank0(psd, n_subjects_total, group)
    return de_all, psd_all
