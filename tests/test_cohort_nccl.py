"""The sharded cohort on real GPUs over NCCL: the gathered result must equal the single-GPU result bit for bit
("sharded result equals single-GPU result bit-for-bit").  Needs >= 2 CUDA devices:

    gpurun --gpus 2 -- python -m pytest tests/test_cohort_nccl.py -m gpu -q

Every rank synthesises its own subjects (seeded by GLOBAL subject id), runs the fused kernel, and rank 0 -- which
alone sees the gathered tensors -- recomputes EVERY subject of EVERY rank locally from the same seeds and compares.
Also covers one process driving two devices (per-device library state).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _n_devices():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs_two = pytest.mark.skipif(_n_devices() < 2, reason="needs >= 2 CUDA devices")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _blocks(subject, device, n_blocks=2):
    from eeg2video_b200 import synth
    return synth.synth_blocks(n_blocks, 1000 + subject, device=device)


def _worker(rank, world, port, n_subjects, mode, chunk, gather, result_path):
    from eeg2video_b200 import cohort, frontend
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = cohort.shard_bounds(n_subjects, rank, world)
        raw = torch.stack([_blocks(s, dev) for s in range(lo, hi)])               # (n_local, 2, 62, T)
        de, psd = cohort.process_cohort(raw, n_subjects, mode=mode, chunk_subjects=chunk, gather=gather)
        torch.cuda.synchronize()
        if rank == 0:
            ok, bad = True, []
            for s in range(n_subjects):                                            # every rank's every subject
                want_de, want_psd = frontend.de_psd_from_raw(_blocks(s, dev), mode)
                if not (torch.equal(de[s], want_de) and torch.equal(psd[s], want_psd)):
                    ok = False
                    bad.append(s)
            torch.save({"ok": ok, "bad": bad, "shape": tuple(de.shape)}, result_path)
        else:
            assert de is None and psd is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@needs_two
@pytest.mark.parametrize("n_subjects,mode,chunk,gather", (
    (6, "500ms", None, "psd"), (5, "500ms", 2, "psd"), (7, "1s", 1, "psd"), (4, "2s", None, "psd"),
    (6, "500ms", None, "both"), (5, "2s", None, "both")))
def test_gathered_cohort_equals_single_gpu(tmp_path, n_subjects, mode, chunk, gather):
    world = 2
    path = os.path.join(str(tmp_path), "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_subjects, mode, chunk, gather, path), nprocs=world, join=True)
    res = torch.load(path)
    assert res["ok"], f"subjects that differ from the single-GPU result: {res['bad']}"
    assert res["shape"][0] == n_subjects


@needs_two
def test_one_process_drives_two_devices():
    """libeegfe keeps cudaFuncSetAttribute / SM-count state per device: the first launch on a second GPU of the same
    process used to fail with cudaErrorInvalidValue."""
    from eeg2video_b200 import frontend, pipeline
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        raw = _blocks(3, dev)
        per_mode = [frontend.de_psd_from_raw(raw, m) for m in ("500ms", "1s", "2s")]
        wins = frontend.sliding_windows(frontend.segment_clips(raw[:1]))
        per_mode.append(frontend.de_psd_windows(wins))
        host = raw.cpu().pin_memory()
        per_mode.append(tuple(t.to(dev) for t in pipeline.features_from_host(host, "500ms", device=dev)))
        torch.cuda.synchronize(dev)
        outs.append([(a.cpu(), b.cpu()) for a, b in per_mode])
    for (a0, b0), (a1, b1) in zip(*outs):
        assert torch.equal(a0, a1) and torch.equal(b0, b1)
