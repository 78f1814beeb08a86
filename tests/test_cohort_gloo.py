"""The N > 1 path (subject sharding + gather to rank 0) on CPU with the gloo backend, world size 2.
The compute step is a stand-in with the same contract as the fused kernel; the kernel itself is covered by the
gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eeg2video_b200 import cohort


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _standin(raw):
    """(n, 7, ch, T) -> 'de', 'psd' of shape (n, 7, 3, ch, 5): cheap deterministic function of the input."""
    n, b, ch, _ = raw.shape
    base = raw[..., :15].reshape(n, b, ch, 3, 5).permute(0, 1, 3, 2, 4).contiguous()
    return base * 2.0, base + 1.0


def _rebuild(psd, de):
    """Stand-in for ops.de_from_psd_: the elementwise map that turns _standin's 'psd' into its 'de'."""
    de.copy_((psd - 1.0) * 2.0)


def _worker(rank, world, port, n_subjects, result_path, overlap=False, gather="psd", chunk=2):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_subjects * 7 * 4 * 32, dtype=torch.float32).reshape(n_subjects, 7, 4, 32)
        lo, hi = cohort.shard_bounds(n_subjects, rank, world)
        de, psd = cohort.process_cohort(full[lo:hi], n_subjects, chunk_subjects=chunk, compute=_standin,
                                        overlap=overlap, gather=gather, rebuild_de=_rebuild)
        if rank == 0:
            want_de, want_psd = _standin(full)
            ok = torch.equal(de, want_de) and torch.equal(psd, want_psd)
            torch.save({"ok": ok, "shape": tuple(de.shape)}, result_path)
        else:
            assert de is None and psd is None
    finally:
        dist.destroy_process_group()


def _run(n_subjects, tmp_path, overlap=False, gather="psd", chunk=2):
    path = os.path.join(tmp_path, f"res_{n_subjects}_{int(overlap)}_{gather}_{chunk}.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_subjects, path, overlap, gather, chunk), nprocs=2, join=True)
    res = torch.load(path)
    assert res["ok"] and res["shape"][0] == n_subjects


def test_psd_only_gather_rebuilds_de(tmp_path):
    """The default: PSD crosses point to point, chunk by chunk; rank 0 rebuilds DE -- even shards, ragged shards
    (rank 0 has more chunks than rank 1 and vice versa), one chunk per rank."""
    _run(6, str(tmp_path))
    _run(5, str(tmp_path))
    _run(7, str(tmp_path), chunk=3)
    _run(4, str(tmp_path), chunk=None)


def test_even_shards_gather(tmp_path):
    _run(6, str(tmp_path), gather="both")


def test_ragged_shards_gather(tmp_path):
    _run(5, str(tmp_path), gather="both")


def test_overlapped_chunked_gather(tmp_path):
    """Chunk i's gather is in flight while chunk i + 1 is computed; same result as the sequential path."""
    _run(10, str(tmp_path), overlap=True, gather="both")    # 5 subjects per rank, chunks of 2, 2, 1
    _run(5, str(tmp_path), overlap=True, gather="both")     # ragged shards: falls back to the sequential path


def test_streaming_loader_matches_resident_tensor():
    """run_cohort with a loader that materialises one chunk at a time (how a cohort larger than HBM is processed)."""
    full = torch.arange(5 * 7 * 4 * 32, dtype=torch.float32).reshape(5, 7, 4, 32)
    seen = []

    def loader(lo, hi):
        seen.append((lo, hi))
        return full[lo:hi].clone()
    de, psd = cohort.run_cohort(5, loader, 5, chunk_subjects=2, compute=_standin, rebuild_de=_rebuild)
    want = _standin(full)
    assert seen == [(0, 2), (2, 4), (4, 5)]
    assert torch.equal(de, want[0]) and torch.equal(psd, want[1])


def test_single_process_is_identity():
    full = torch.arange(3 * 7 * 4 * 32, dtype=torch.float32).reshape(3, 7, 4, 32)
    de, psd = cohort.process_cohort(full, 3, compute=_standin, rebuild_de=_rebuild)
    want = _standin(full)
    assert torch.equal(de, want[0]) and torch.equal(psd, want[1])
