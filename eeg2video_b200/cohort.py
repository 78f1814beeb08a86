"""Cohort-level driver: shard subjects over the GPUs of one box, run the fused kernel per shard, gather the
feature tensors to rank 0.

Every channel-window is independent, so there is NO communication on the hot path: each rank processes its own
contiguous range of subjects out of its own HBM.  The only exchange is the final gather of the features to rank 0
over NCCL (NVLink 5 / NVSwitch), and rank 0's NVLink ingest (~0.77 TB/s measured peer copy) is what bounds it, so
the gather moves as few bytes as possible and starts as early as possible:

  * only PSD crosses the links.  DE = log2(100 PSD) is an elementwise function of the float32 PSD values, and
    ``eegfe_de_from_psd`` evaluates it with the device expression of the feature kernels: rank 0 rebuilds DE bit for
    bit from what it received (half the bytes of gathering both tensors);
  * point-to-point, per chunk: a rank sends chunk k's PSD as soon as that chunk's kernel has been enqueued, rank 0
    posts the matching receives (one NCCL group per round) before it computes its own chunk k, and rebuilds DE for
    round k - 1 while round k is in flight.

``gather="both"`` keeps the plain collective (NCCL ``gather`` of DE and of PSD) for comparison.

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests (which exercise the sharding and
gather logic with stand-in compute functions; the kernels are covered by the gpu tests, the NCCL path by
tests/test_cohort_nccl.py).
"""
import contextlib
import os

import torch
import torch.distributed as dist

from . import frontend

# SMs a rank leaves to NCCL while a gather is in flight (the feature kernels otherwise fill every SM and the NCCL send /
# receive kernels wait for a CTA to retire): the gathering rank receives from world - 1 peers, a sender feeds one.
# Measured on 8 B200s (tools/gather_sweep.py, 24 subjects per GPU, 8 chunks): compute + gather 4.09 ms with 0 / 0 SMs
# left, 3.88 ms with 20 / 8, 3.51 ms with 32 / 8, 3.45 ms with 32 / 16, 3.54 ms with 48 / 8, 3.65 ms with 64 / 8.
RESERVED_SMS_DST = int(os.environ.get("EEGFE_COHORT_RESERVED_SMS_DST", "32"))
RESERVED_SMS_SRC = int(os.environ.get("EEGFE_COHORT_RESERVED_SMS_SRC", "16"))


@contextlib.contextmanager
def _leave_sms_to_nccl(device, reserved):
    """Cap the persistent kernels' grid at (SMs - reserved) for the duration of a distributed cohort run."""
    if reserved <= 0 or device is None or device.type != "cuda":
        yield
        return
    from . import _lib
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    old = _lib.set_cta_limit(max(1, sms - reserved))
    try:
        yield
    finally:
        _lib.set_cta_limit(old)


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


def chunk_ranges(n_local, chunk):
    """[(lo, hi), ...] covering range(n_local) in steps of `chunk` (None / 0: one chunk)."""
    step = int(chunk) if chunk else max(int(n_local), 1)
    return [(lo, min(lo + step, n_local)) for lo in range(0, n_local, step)]


def _distributed(group):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def gather_to_rank0(local, n_total, group=None, dst=0, out=None):
    """Gather per-rank tensors (leading axis = that rank's subjects, in shard order) into one tensor of
    `n_total` leading entries on rank `dst`.  Returns the full tensor on `dst`, None elsewhere.
    Uneven shards are handled (sizes follow shard_bounds).  `out`: preallocated destination on `dst`."""
    if not _distributed(group):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: expected {sizes[rank]} leading entries, got {local.shape[0]}")
    local = local.contiguous()
    if rank == dst:
        full = out if out is not None else torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype,
                                                       device=local.device)
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


def _default_compute(mode):
    return lambda x: frontend.de_psd_from_raw(x, mode, check=False)


def _default_rebuild(status_box):
    """de <- log2(100 psd) with the kernels' own expression (ops.de_from_psd_); zero-power flags accumulate in a
    device int that `status_box` holds."""
    from . import ops

    def rebuild(psd, de):
        if status_box[0] is None:
            status_box[0] = torch.zeros(1, dtype=torch.int32, device=psd.device)
        ops.de_from_psd_(psd, de, status_box[0])
    return rebuild


def process_shard(raw, mode="500ms", chunk_subjects=None, compute=None):
    """Run the fused kernel over this rank's subjects.

    raw: (n_local_subjects, 7, 62, T) float32 on this rank's GPU.  Returns (de, psd) with the subject axis leading.
    `compute` (test hook) replaces frontend.de_psd_from_raw with another callable of the same contract.
    """
    fn = compute or _default_compute(mode)
    n = raw.shape[0]
    if n == 0:
        return None, None
    des, psds = [], []
    for lo, hi in chunk_ranges(n, chunk_subjects):
        de, psd = fn(raw[lo:hi])
        des.append(de)
        psds.append(psd)
    return (des[0], psds[0]) if len(des) == 1 else (torch.cat(des), torch.cat(psds))


def run_cohort(n_local, loader, n_subjects_total, mode="500ms", chunk_subjects=None, group=None, compute=None,
               rebuild_de=None, dst=0, out=None, device=None):
    """Shard-local compute, chunk by chunk, + PSD-only point-to-point gather + DE rebuilt on rank `dst`.
    See _run_cohort for the arguments; `device` (default: the current CUDA device when the backend is NCCL) is the GPU
    whose feature kernels leave RESERVED_SMS_* SMs to the NCCL kernels while the gather is in flight."""
    if _distributed(group) and dist.get_backend(group) == "nccl":
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        reserved = RESERVED_SMS_DST if dist.get_rank(group) == dst else RESERVED_SMS_SRC
        with _leave_sms_to_nccl(dev, reserved):
            return _run_cohort(n_local, loader, n_subjects_total, mode, chunk_subjects, group, compute, rebuild_de, dst, out)
    return _run_cohort(n_local, loader, n_subjects_total, mode, chunk_subjects, group, compute, rebuild_de, dst, out)


def _run_cohort(n_local, loader, n_subjects_total, mode="500ms", chunk_subjects=None, group=None, compute=None,
                rebuild_de=None, dst=0, out=None):
    """Shard-local compute, chunk by chunk, + PSD-only point-to-point gather + DE rebuilt on rank `dst`.

    loader(lo, hi) -> raw recordings (hi - lo, 7, ch, T) of this rank's LOCAL subjects lo..hi-1 on this rank's device
    (a slice of a resident tensor, or a function that reads / synthesises one chunk at a time -- this is how a cohort
    larger than one GPU's memory is streamed: 1000 subjects are 180 GB of raw recordings).
    Returns (de, psd) for the whole cohort on rank `dst`, (None, None) elsewhere.  Every rank must call it with the
    same n_subjects_total and chunk_subjects; n_local must equal this rank's shard size (shard_bounds).
    rebuild_de(psd, de): fills `de` from `psd` (default: ops.de_from_psd_; CPU tests pass a stand-in).
    out: (de, psd) preallocated cohort tensors on rank `dst` (a pipeline that runs cohort after cohort reuses them).
    """
    fn = compute or _default_compute(mode)
    distributed = _distributed(group)
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    sizes = shard_sizes(n_subjects_total, world)
    if n_local != sizes[rank]:
        raise ValueError(f"rank {rank}: shard holds {n_local} subjects, expected {sizes[rank]}")
    starts = [sum(sizes[:r]) for r in range(world)]
    chunks = [chunk_ranges(sizes[r], chunk_subjects) for r in range(world)]
    n_rounds = max(len(c) for c in chunks) if chunks else 0
    status_box = [None]
    rebuild = rebuild_de or _default_rebuild(status_box)

    if rank != dst:
        keep, reqs = [], []
        for k, (lo, hi) in enumerate(chunks[rank]):
            _, psd = fn(loader(lo, hi))
            psd = psd.contiguous()
            keep.append(psd)                                   # alive until its transfer has finished
            reqs += dist.batch_isend_irecv([dist.P2POp(dist.isend, psd, dst, group)])
        for q in reqs:
            q.wait()
        return None, None

    full_de, full_psd = out if out is not None else (None, None)
    pending = []                                               # [(requests, [(lo, hi) global subject ranges])]

    def finish(entry):
        reqs, spans = entry
        for q in reqs:
            q.wait()
        for lo, hi in spans:
            rebuild(full_psd[lo:hi], full_de[lo:hi])

    for k in range(n_rounds):
        own = chunks[rank][k] if k < len(chunks[rank]) else None
        de = psd = None
        if own is not None and full_psd is None:
            de, psd = fn(loader(*own))                         # the first chunk also tells the feature shape
            shape = (n_subjects_total,) + tuple(psd.shape[1:])
            full_psd = torch.empty(shape, dtype=psd.dtype, device=psd.device)
            full_de = torch.empty(shape, dtype=de.dtype, device=de.device)
        if full_psd is None:
            raise ValueError("the gathering rank needs at least one subject of its own to learn the feature shape")
        ops_k, spans = [], []
        if distributed:
            for r in range(world):
                if r == dst or k >= len(chunks[r]):
                    continue
                lo, hi = chunks[r][k]
                spans.append((starts[r] + lo, starts[r] + hi))
                ops_k.append(dist.P2POp(dist.irecv, full_psd[starts[r] + lo:starts[r] + hi], r, group))
        reqs = dist.batch_isend_irecv(ops_k) if ops_k else []
        if own is not None:
            if psd is None:
                de, psd = fn(loader(*own))
            lo, hi = starts[rank] + own[0], starts[rank] + own[1]
            if psd.data_ptr() != full_psd[lo:hi].data_ptr():       # a compute hook may write straight into `out`
                full_psd[lo:hi].copy_(psd)
            if de.data_ptr() != full_de[lo:hi].data_ptr():
                full_de[lo:hi].copy_(de)
        if pending:
            finish(pending.pop(0))                             # DE of round k - 1 while round k is in flight
        pending.append((reqs, spans))
    while pending:
        finish(pending.pop(0))
    if status_box[0] is not None:
        frontend.raise_if_zero_power(status_box[0])
    return full_de, full_psd


def process_cohort(raw_local, n_subjects_total, mode="500ms", chunk_subjects=None, group=None, compute=None,
                   overlap=False, gather="psd", rebuild_de=None, out=None):
    """Shard-local compute + gather to rank 0.  Returns (de, psd) for the whole cohort on rank 0, (None, None)
    elsewhere.

    gather="psd" (default): run_cohort -- PSD-only point-to-point gather, DE rebuilt on rank 0, chunk k's transfer
    overlapping chunk k + 1's kernel when chunk_subjects is given.
    gather="both": the plain collective, NCCL `gather` of DE and of PSD after all kernels (with overlap=True and equal
    shards: one asynchronous gather pair per chunk).  Measured on 8 B200s, 24 subjects per GPU: 4.08 GB into rank 0
    in 5.8 ms against 1.25 ms of compute -- kept as the baseline the default is compared with in bench.py.
    """
    if gather == "psd":
        return run_cohort(raw_local.shape[0], lambda lo, hi: raw_local[lo:hi], n_subjects_total, mode, chunk_subjects,
                          group, compute, rebuild_de, out=out)
    if gather != "both":
        raise ValueError("gather must be 'psd' or 'both'")
    if overlap and _distributed(group) and chunk_subjects:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        sizes = shard_sizes(n_subjects_total, world)
        if all(s == sizes[0] for s in sizes) and sizes[0] > 0:
            return _process_cohort_overlapped(raw_local, sizes[0], world, rank, mode, chunk_subjects, group, compute)
    de, psd = process_shard(raw_local, mode, chunk_subjects, compute)
    de_all = gather_to_rank0(de, n_subjects_total, group, out=None if out is None else out[0])
    psd_all = gather_to_rank0(psd, n_subjects_total, group, out=None if out is None else out[1])
    return de_all, psd_all


def _process_cohort_overlapped(raw_local, n_local, world, rank, mode, chunk, group, compute):
    fn = compute or _default_compute(mode)
    full = [None, None]
    works, keep = [], []
    for lo, hi in chunk_ranges(n_local, chunk):
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
