#!/usr/bin/env python
"""bench.py -- DE+PSD channel-windows/s of the fused front end on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--subjects S] [--mode 500ms|1s|2s]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference [--gpus N --steps K --warmup W]

Workload (BASELINE.json configs[1] at cohort scale, north_star): 500 ms sliding windows (250 ms hop, 7 per 2 s
clip) cut straight out of raw synthetic SEED-DV recordings (62 ch, 200 Hz, 7 blocks x 104000 samples per
subject), S subjects resident per GPU (weak scaling: per-GPU work is fixed as N grows).  One *step* = one pass
of the fused kernel over the whole resident batch = S * 607600 channel-windows.  The batch (S * 180.5 MB) is
many times the 126 MB L2, so no L2 flush is needed between steps.

One JSON line is printed by rank 0; see the README section "Benchmark" / DESIGN.md "Measurement" for the keys.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CW_PER_SUBJECT = {"500ms": 7 * 40 * 5 * 7 * 62, "1s": 7 * 40 * 5 * 2 * 62, "2s": 7 * 40 * 5 * 62}
# algorithmic bytes per channel-window (SURVEY.md 8d / BASELINE.md 4): live input samples + 40 B of features
BYTES_PER_CW = {"500ms": 1600.0 / 7 + 40, "1s": 840.0, "2s": 840.0}
# fp32 lane-operations per channel-window of the pruned FFT + band power (counted from the SASS of the kernel:
# 2 x packed f32x2 instructions + scalar FP instructions per window)
FP32_LANE_OPS_PER_CW = {"500ms": 2420.0, "1s": 2660.0, "2s": 2660.0}
KERNEL_NAME = {"500ms": "eegfe::de_psd_stream_kernel (500 ms sliding, fused segmentation)",
               "1s": "eegfe::de_psd_kernel<CfgOneSec>", "2s": "eegfe::de_psd_kernel<CfgTwoSec>"}
_REAL_STDOUT = None


def capture_stdout():
    """Everything that libraries print to fd 1 (NCCL's version banner, warnings) goes to stderr; the JSON line is
    the only thing written to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


_T0 = time.perf_counter()


def log(msg):
    """Progress notes on stderr (the JSON line is the only thing on stdout)."""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write(f"[bench {time.perf_counter() - _T0:7.1f} s] {msg}\n")
        sys.stderr.flush()


METRIC = "de_psd_channel_windows_per_s"
UNIT = "channel-windows/s"


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's loop-for-loop port of the reference (oracle.de_psd_loop), one process per host core
# ---------------------------------------------------------------------------------------------------------------
_CPU_CLIPS = None


def _cpu_init(path):
    global _CPU_CLIPS
    import numpy as np
    _CPU_CLIPS = np.load(path, mmap_mode="r")


REF_STAGE = os.path.join(ROOT, "baseline", "_ref")          # unmodified copy of the reference's EEG_preprocessing/*.py,
                                                              # staged by __graft_entry__.build() (git-ignored)


_REF_DE_PSD = None


def reference_de_psd():
    """(DE_PSD callable, kind): the reference's own DE_PSD when its files are staged under baseline/_ref (kind
    "reference"), else the oracle's loop-for-loop port (kind "port")."""
    global _REF_DE_PSD
    if _REF_DE_PSD is None:
        _REF_DE_PSD = _load_reference_de_psd()
    return _REF_DE_PSD


def _load_reference_de_psd():
    path = os.path.join(REF_STAGE, "EEG_preprocessing", "DE_PSD.py")
    if os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_eeg2video_reference_DE_PSD", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.DE_PSD, "reference"
    import oracle
    return oracle.de_psd_loop, "port"


def _cpu_work(args):
    """Reference arithmetic for `n` clips starting at `first` (500 ms driver loop, 1per500ms.py:20-27)."""
    first, n, mode = args
    import numpy as np
    de_psd, _ = reference_de_psd()
    clips = _CPU_CLIPS
    done = 0
    for i in range(first, first + n):
        clip = np.asarray(clips[i % clips.shape[0]])
        if mode == "500ms":
            for w in range(7):
                de_psd(clip[:, 50 * w:50 * w + 100], 200, 0.5)
            done += 7 * clip.shape[0]
        elif mode == "1s":
            de_psd(clip[:, :200], 200, 1)
            de_psd(clip[:, 200:], 200, 1)
            done += 2 * clip.shape[0]
        else:
            de_psd(clip, 200, 2)
            done += clip.shape[0]
    return done


class CpuArm:
    """Pool of worker processes running the reference's per-channel Python loops on a bounded sample of clips."""

    def __init__(self, mode, n_sample_clips=200):
        import numpy as np
        import torch
        from eeg2video_b200 import synth
        import oracle
        self.mode = mode
        self.kind = reference_de_psd()[1]
        self.what = ("the reference's own EEG_preprocessing/DE_PSD.py (unmodified copy under baseline/_ref)"
                     if self.kind == "reference" else "oracle.de_psd_loop (loop-for-loop port of DE_PSD.py:8-71)")
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        raw = synth.synth_blocks(1, 1001, device="cpu").numpy()                      # one block, (1, 62, 104000)
        padded = np.concatenate([raw, np.zeros((6,) + raw.shape[1:], np.float32)])
        clips = oracle.segment_subject(padded)[0].reshape(200, 62, 400)[:n_sample_clips]
        self.tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False)
        np.save(self.tmp.name, clips)
        self.n_clips = clips.shape[0]
        torch.set_num_threads(1)
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(self.tmp.name,))
        self.pool.map(_cpu_work, [(0, 1, mode)] * self.cores)                          # start-up, untimed

    def run(self, clips_per_worker):
        """Every worker processes `clips_per_worker` clips; returns (channel-windows, seconds)."""
        jobs = [(w * clips_per_worker, clips_per_worker, self.mode) for w in range(self.cores)]
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_work, jobs, chunksize=1))
        return done, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
        os.unlink(self.tmp.name)


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (its Python loops, restated in
    oracle/de_psd.py -- the reference itself is Python and cannot travel to the GPU box), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = CpuArm(args.mode)
    per_step = args.cpu_clips_per_step
    for _ in range(args.warmup):
        arm.run(max(1, per_step // 4))
    total_cw, total_s = 0, 0.0
    for _ in range(args.steps):
        cw, s = arm.run(per_step)
        total_cw += cw
        total_s += s
    arm.close()
    value = total_cw / total_s
    sample = (f"{args.steps} steps x {arm.cores} workers x {per_step} synthetic 2 s clips (62 ch) each, "
              f"mode {args.mode}, {arm.what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def workload_config(args, geometry):
    cfg = {
        "workload": f"extract_DE_PSD_features_1per{args.mode} fused with segmentation from raw SEED-DV-shaped "
                    f"recordings (BASELINE configs[1] at cohort scale, configs[3] sharding)",
        "mode": args.mode, "subjects_per_gpu": args.subjects, "blocks_per_subject": 7, "channels": 62,
        "fs_hz": 200, "block_len": 104000,
        "channel_windows_per_step_per_gpu": args.subjects * CW_PER_SUBJECT[args.mode],
        "l2_policy": f"inputs larger than L2 ({args.subjects * 180.544:.0f} MB resident batch per GPU vs 126 MB L2)",
        "parallelism": f"subjects sharded over {args.gpus} GPU(s), no data-path collective",
    }
    if geometry:
        cfg["kernel_geometry"] = geometry
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md clocks line)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.out = None

    def start(self):
        try:
            self.out = tempfile.NamedTemporaryFile(mode="w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @staticmethod
    def _epoch(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, windows=None):
        """Summary over all samples; `windows` = {name: (t0, t1)} (time.time() bounds) adds one summary per window,
        from the samples whose own timestamp falls inside it."""
        empty = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return (empty, {k: dict(empty) for k in windows}) if windows else empty
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        rows = []
        for row in self.out.read().strip().splitlines():
            cells = [c.strip() for c in row.split(",")]
            if len(cells) < 9:
                continue
            try:
                rec = (self._epoch(cells[0]), float(cells[1]), float(cells[2]), float(cells[3]))
            except ValueError:
                continue
            flags = [name for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                 "sw_power_cap"), cells[5:9]) if cell.lower().startswith("active")]
            rows.append(rec + (flags,))
        self.out.close()
        os.unlink(self.out.name)

        def summary(sel):
            sm = sorted(r[1] for r in sel)
            reasons = sorted({f for r in sel for f in r[4]})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in sel), default=None),
                    "power_w_max": max((r[3] for r in sel), default=None), "reasons": reasons, "samples": len(sm)}
        total = summary(rows)
        if not windows:
            return total
        return total, {k: summary([r for r in rows if r[0] is not None and t0 <= r[0] <= t1])
                       for k, (t0, t1) in windows.items()}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture, scaled
    to this run's launch size by the profile's own bytes-per-channel-window (None until a capture exists)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        rec = json.load(f).get(mode)
    return rec


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from eeg2video_b200 import _lib, cohort, frontend, ops, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    if args.tensor_loads:
        _lib.set_tensor_loads(True)
    mode = args.mode
    mode_id = frontend.MODES[mode]
    S = args.subjects
    cw_step_gpu = S * CW_PER_SUBJECT[mode]

    # ---- resident synthetic batch: this rank's subjects (global ids rank*S .. rank*S+S-1) ----
    raw = synth.synth_cohort(range(rank * S, rank * S + S), dev).reshape(S * 7, 62, 104000)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log("resident batch ready; parity gate")
    # ---- parity gate on EVERY rank's first subject (oracle = checker only); the line reports the worst rank ----
    parity = None
    if not args.skip_parity:
        import numpy as np
        import oracle
        blocks = raw[:1]
        de, psd, _ = ops.de_psd_from_raw(blocks, mode_id)
        padded = np.concatenate([blocks.cpu().numpy(), np.zeros((6, 62, 104000), np.float32)])
        clips = oracle.segment_subject(padded)[0]
        seg_ok = bool(np.array_equal(ops.segment_clips(blocks, 200).cpu().numpy().reshape(clips.shape), clips))
        win, hop, nwin, tw = {"500ms": (100, 50, 7, 0.5), "1s": (200, 200, 2, 1), "2s": (400, 0, 1, 2)}[mode]
        wins = np.stack([clips[..., w * hop:w * hop + win] for w in range(nwin)], axis=2)   # (40,5,W,62,win)
        de_ref, psd_ref = oracle.de_psd_closed_form(wins, 200, tw)
        got_de = de.cpu().numpy().reshape(de_ref.shape)
        got_psd = psd.cpu().numpy().reshape(psd_ref.shape)
        mine = [0.0 if seg_ok else 1.0, float(np.max(np.abs(got_psd - psd_ref) / psd_ref)),
                float(np.max(np.abs(got_de - de_ref)))]
        if world > 1:
            t = torch.tensor(mine, dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mine = [float(v) for v in t.tolist()]
        parity = {"segmentation_bit_exact": mine[0] == 0.0, "psd_max_rel": mine[1], "de_max_abs": mine[2],
                  "checked_channel_windows": int(de_ref.size // 5) * world, "ranks_checked": world}
        if not (parity["segmentation_bit_exact"] and parity["psd_max_rel"] <= 1e-4 and parity["de_max_abs"] <= 1e-4):
            raise SystemExit(f"parity gate failed: {parity}")

    log("timed region")
    # ---- device-resident throughput: K launches of the fused kernel, CUDA events on the launch stream ----
    with torch.cuda.device(dev):
        de_buf = torch.empty((S * 7 * 200, ops.WINDOWS_PER_CLIP[mode_id], 62, 5), dtype=torch.float32, device=dev)
        psd_buf = torch.empty_like(de_buf)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev)

    def step():
        _lib.check(lib.eegfe_de_psd_from_raw(raw.data_ptr(), raw.shape[0], 62, 104000, raw.stride(0), raw.stride(1),
                                             mode_id, de_buf.data_ptr(), psd_buf.data_ptr(), status.data_ptr(),
                                             stream.cuda_stream))

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches_before = _lib.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("eegfe_timed")        # ncu --nvtx --nvtx-include "eegfe_timed/" lists exactly this region
    t_burst0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.nvtx.range_pop()
    barrier()
    t_burst1 = time.time()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    gpu_launches = _lib.launch_count() - launches_before
    value = world * cw_step_gpu * args.steps / (elapsed_ms * 1e-3)

    log("sustained run")
    # ---- the same launches back to back for >= args.sustain_s seconds, on every rank: the throughput the 1 kW power
    #      cap allows (`value` above is a ~25 ms burst entered from idle), timed with CUDA events like `value`, with
    #      the SM clock sampled over exactly this window ----
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s / (elapsed_ms * 1e-3 / args.steps)) + 1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_sus0 = time.time()
        e0.record(stream)
        for _ in range(n_sus):
            step()
        e1.record(stream)
        barrier()
        t_sus1 = time.time()
        sus_ms = max_over_ranks(e0.elapsed_time(e1))
        sustained = {"value": world * cw_step_gpu * n_sus / (sus_ms * 1e-3), "unit": UNIT, "steps": n_sus,
                     "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus}
    if rank == 0:
        windows = {"burst": (t_burst0, t_burst1)}
        if sustained is not None:
            # skip the first 0.3 s: the clock is still settling from the burst level
            windows["sustained"] = (t_sus0 + min(0.3, 0.25 * (t_sus1 - t_sus0)), t_sus1)
        clocks, per_window = sampler.stop(windows)
        clocks["sampled_over"] = "timed region + the sustained continuation of the same launches (nvidia-smi, 50 ms period)"
        clocks["timed_region_only"] = per_window["burst"]
        if sustained is not None:
            sustained["clocks"] = per_window["sustained"]

    log("single subject, other modes, next rows")
    # ---- BASELINE configs[1] as written: ONE subject, one launch (latency-bound; SURVEY.md 8d asks for us per call) ----
    single = None
    if rank == 0:
        one = raw[:7]
        s_de = torch.empty((7 * 200, ops.WINDOWS_PER_CLIP[mode_id], 62, 5), dtype=torch.float32, device=dev)
        s_psd = torch.empty_like(s_de)

        def s_step():
            _lib.check(lib.eegfe_de_psd_from_raw(one.data_ptr(), 7, 62, 104000, one.stride(0), one.stride(1), mode_id,
                                                 s_de.data_ptr(), s_psd.data_ptr(), status.data_ptr(),
                                                 stream.cuda_stream))
        for _ in range(3):
            s_step()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(50):
            s_step()
        s1.record(stream)
        torch.cuda.synchronize()
        s_us = 1e3 * s0.elapsed_time(s1) / 50
        single = {"us_per_call": s_us, "value": CW_PER_SUBJECT[mode] / (s_us * 1e-6), "unit": UNIT,
                  "note": "one subject = 180.5 MB of raw input, comparable to the 126 MB L2; 50 back-to-back launches"}

    # ---- the other two analysis modes over the same resident batch (kernel only; informative, rank 0) ----
    other_modes = {}
    if rank == 0 and not args.skip_other_modes:
        for om in ("500ms", "1s", "2s"):
            if om == mode:
                continue
            om_id = frontend.MODES[om]
            o_de = torch.empty((S * 7 * 200, ops.WINDOWS_PER_CLIP[om_id], 62, 5), dtype=torch.float32, device=dev)
            o_psd = torch.empty_like(o_de)

            def o_step():
                _lib.check(lib.eegfe_de_psd_from_raw(raw.data_ptr(), raw.shape[0], 62, 104000, raw.stride(0),
                                                     raw.stride(1), om_id, o_de.data_ptr(), o_psd.data_ptr(),
                                                     status.data_ptr(), stream.cuda_stream))
            for _ in range(3):
                o_step()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record(stream)
            for _ in range(args.steps):
                o_step()
            o1.record(stream)
            torch.cuda.synchronize()
            o_ms = o0.elapsed_time(o1) / args.steps
            o_cw = S * CW_PER_SUBJECT[om]
            o_gbs = o_cw * BYTES_PER_CW[om] / (o_ms * 1e-3) / 1e9
            other_modes[om] = {"value": o_cw / (o_ms * 1e-3), "unit": UNIT, "kernel_ms": o_ms,
                               "hbm_gbs": o_gbs, "hbm_frac": o_gbs / measured_peaks()[0], "kernel": KERNEL_NAME[om]}
            del o_de, o_psd
        # BASELINE configs[1], second entry: the reference's own call pattern, DE_PSD on MATERIALISED
        # 500 ms windows (S, 7, 40, 5, 7, 62, 100) -- 440 B per channel-window instead of the fused 268.6 B
        clips_all = ops.segment_clips(raw, 200)
        wins = ops.sliding_windows(clips_all).reshape(-1, 100)
        del clips_all
        w_de = torch.empty((wins.shape[0], 5), dtype=torch.float32, device=dev)
        w_psd = torch.empty_like(w_de)

        def w_step():
            _lib.check(lib.eegfe_de_psd_windows(wins.data_ptr(), wins.shape[0], 100, 100, w_de.data_ptr(),
                                                w_psd.data_ptr(), status.data_ptr(), stream.cuda_stream))
        for _ in range(3):
            w_step()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for _ in range(args.steps):
            w_step()
        o1.record(stream)
        torch.cuda.synchronize()
        o_ms = o0.elapsed_time(o1) / args.steps
        o_gbs = wins.shape[0] * 440.0 / (o_ms * 1e-3) / 1e9
        other_modes["500ms_precut_windows"] = {
            "value": wins.shape[0] / (o_ms * 1e-3), "unit": UNIT, "kernel_ms": o_ms, "hbm_gbs": o_gbs,
            "hbm_frac": o_gbs / measured_peaks()[0], "algorithmic_bytes_per_channel_window": 440.0,
            "kernel": "eegfe::de_psd_stream_kernel<StreamCfgWin100> (eegfe_de_psd_windows)"}
        del wins, w_de, w_psd

    # ---- recordings whose rows are NOT 16-byte aligned (an odd block length: VERDICT round 1, weak 8): the same data in
    #      buffers of T = 104001 (rows 4-byte aligned) and T = 104002 (8-byte aligned) samples per row -- one launch each,
    #      TMA copies of the aligned span around every row, read shifted; results must equal the aligned ones bit for bit
    if rank == 0 and not args.skip_other_modes:
        nb = min(S, 8) * 7
        unaligned = {"subjects": nb // 7, "matches_aligned_rows": True}
        for t_len in (104001, 104002):
            odd = torch.zeros((nb, 62, t_len), dtype=torch.float32, device=dev)
            odd[..., :104000].copy_(raw[:nb])
            for om in ("500ms", "1s", "2s"):
                om_id = frontend.MODES[om]
                u_de = torch.empty((nb * 200, ops.WINDOWS_PER_CLIP[om_id], 62, 5), dtype=torch.float32, device=dev)
                u_psd = torch.empty_like(u_de)
                a_de, a_psd = torch.empty_like(u_de), torch.empty_like(u_de)

                def u_step(src=odd, t=t_len, de=u_de, psd=u_psd):
                    _lib.check(lib.eegfe_de_psd_from_raw(src.data_ptr(), nb, 62, t, src.stride(0), src.stride(1), om_id,
                                                         de.data_ptr(), psd.data_ptr(), status.data_ptr(),
                                                         stream.cuda_stream))
                u_step(raw, 104000, a_de, a_psd)                        # the aligned rows of the same recordings
                for _ in range(3):
                    u_step()
                o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                o0.record(stream)
                for _ in range(args.steps):
                    u_step()
                o1.record(stream)
                torch.cuda.synchronize()
                u_ms = o0.elapsed_time(o1) / args.steps
                unaligned[f"T={t_len} {om}"] = (nb // 7) * CW_PER_SUBJECT[om] / (u_ms * 1e-3)
                if not (torch.equal(u_de, a_de) and torch.equal(u_psd, a_psd)):
                    unaligned["matches_aligned_rows"] = False
                del u_de, u_psd, a_de, a_psd
            del odd
        unaligned["unit"] = UNIT
        other_modes["rows_not_16_byte_aligned"] = unaligned

    # ---- next row: GLMNet input build = normalised 2 s clips + 500 ms features in one pass ----
    next_rows = {}
    if rank == 0 and not args.skip_other_modes:
        from eeg2video_b200 import glmnet_inputs
        mean, std = glmnet_inputs.channel_stats(raw, None)
        scale = (1.0 / std).to(torch.float32).contiguous()
        shift = mean.to(torch.float32).contiguous()                         # (x - mean) * (1 / std)
        g_clips = torch.empty((S * 7 * 200, 62, 400), dtype=torch.float32, device=dev)
        g_de = torch.empty((S * 7 * 200, 7, 62, 5), dtype=torch.float32, device=dev)
        g_psd = torch.empty_like(g_de)

        def g_step():
            _lib.check(lib.eegfe_glmnet_inputs_from_raw(raw.data_ptr(), raw.shape[0], 62, 104000, raw.stride(0),
                                                        raw.stride(1), scale.data_ptr(), shift.data_ptr(),
                                                        g_clips.data_ptr(), g_de.data_ptr(), g_psd.data_ptr(),
                                                        status.data_ptr(), stream.cuda_stream))
        for _ in range(3):
            g_step()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            g_step()
        g1.record(stream)
        torch.cuda.synchronize()
        g_ms = g0.elapsed_time(g1) / args.steps
        rows = S * 7 * 200 * 62                                   # channel-clips
        g_bytes = rows * (1600 + 1600 + 7 * 40)                   # clip in, normalised clip out, 7 x (DE+PSD) out
        next_rows["glmnet_inputs_from_raw"] = {
            "value": rows * 7 / (g_ms * 1e-3), "unit": UNIT, "kernel_ms": g_ms,
            "algorithmic_bytes_per_channel_clip": 3480, "hbm_gbs": g_bytes / (g_ms * 1e-3) / 1e9,
            "hbm_frac": g_bytes / (g_ms * 1e-3) / 1e9 / measured_peaks()[0],
            "kernel": "eegfe::de_psd_stream_kernel<NORM> (500 ms features + per-channel normalised clips, one pass)",
            "features_identical_to_plain_kernel": bool(torch.equal(g_de, de_buf) and torch.equal(g_psd, psd_buf))}
        del g_clips, g_de, g_psd

        # ---- next rows (8f ranks 2, 3): consumer-side input build on the 1 s features of every resident subject ----
        # concept re-ordering + mean over the two windows + 310 columns + StandardScaler, one group per subject
        from eeg2video_b200 import consumers
        import numpy as np
        f_de, _, _ = ops.de_psd_from_raw(raw, frontend.MODES["1s"])                 # (S*7*200, 2, 62, 5)
        f_units = f_de.reshape(S * 1400, 2, 310)
        perm_rng = np.random.default_rng(0)
        gt = np.stack([perm_rng.permutation(40) + 1 for _ in range(7)])             # a synthetic label table
        one = consumers.clip_index(range(6), gt, range(1, 41))                      # 1200 clips of a subject
        idx = torch.from_numpy(np.concatenate([one + s_ * 1400 for s_ in range(S)]).astype(np.int32)).to(dev)

        def c_step():
            x = ops.select_units(f_units, idx, True).reshape(S, 1200, 310)
            m, _, sc = ops.column_stats(x)
            return ops.standardize(x, m, sc)
        for _ in range(3):
            c_ref = c_step()
        c_before = _lib.launch_count()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(args.steps):
            c_out = c_step()
        c1.record(stream)
        torch.cuda.synchronize()
        c_ms = c0.elapsed_time(c1) / args.steps
        c_kernels = int((_lib.launch_count() - c_before) // args.steps)
        # the same six kernels captured once as a CUDA graph: ONE launch per build (consumers.GraphedBuild)
        graphed = consumers.GraphedBuild(c_step)
        for _ in range(3):
            graphed.replay()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            g_out = graphed.replay()
        g1.record(stream)
        torch.cuda.synchronize()
        g_ms = g0.elapsed_time(g1) / args.steps
        # bytes: gather reads 2 windows and writes x; the two statistics passes read x; the transform reads x, writes out
        c_bytes = S * 1200 * 310 * 4 * (2 + 1 + 2 + 2)
        next_rows["consumer_inputs_semantic_1s"] = {
            "subjects": S, "rows": S * 1200, "cols": 310, "kernels_per_build": c_kernels,
            "ms_graph_replay": g_ms, "launches_per_build_graphed": 1,
            "hbm_gbs": c_bytes / (g_ms * 1e-3) / 1e9, "hbm_frac": c_bytes / (g_ms * 1e-3) / 1e9 / measured_peaks()[0],
            "ms_kernel_by_kernel": c_ms, "hbm_frac_kernel_by_kernel": c_bytes / (c_ms * 1e-3) / 1e9 / measured_peaks()[0],
            "graph_matches_kernel_by_kernel": bool(torch.equal(g_out, c_ref)),
            "us_per_subject": 1e3 * g_ms / S,
            "note": "36 MB working set (L2-resident after the gather); byte count = 7 passes over it",
            "column_mean_abs_max": float(c_out.double().mean(dim=1).abs().max())}
        del graphed, g_out, c_ref
        del f_de, f_units, c_out

    log("end to end")
    # ---- end to end: pinned host recordings -> H2D -> fused kernel -> D2H of the features, every step ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    numa_cpus = pipeline.bind_to_gpu_numa_node(dev) if world > 1 else None     # node-local pinned buffers per rank
    # compact="auto": the pipeline times the strided (live samples only) and the contiguous (whole rows) upload of its
    # first chunk during the warm-up run and keeps the faster -- which one wins differs from host to host
    pipe = pipeline.HostPipeline(dev, 62, 104000, chunk_blocks=args.chunk_blocks, mode=mode)
    raw_host = torch.empty((S * 7, 62, 104000), dtype=torch.float32).pin_memory()
    raw_host.copy_(raw)
    de_host = torch.empty(pipe.feature_shape(S * 7), dtype=torch.float32).pin_memory()
    psd_host = torch.empty_like(de_host).pin_memory()
    pipe.run(raw_host, de_host, psd_host)                                   # warm-up
    barrier()
    launches_e2e0 = _lib.launch_count()
    torch.cuda.nvtx.range_push("eegfe_e2e")
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(raw_host, de_host, psd_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    torch.cuda.nvtx.range_pop()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * cw_step_gpu * e2e_steps / e2e_s
    e2e_ok_local = bool(torch.equal(de_host, de_buf.cpu()) and torch.equal(psd_host, psd_buf.cpu()))
    e2e_ok = max_over_ranks(0.0 if e2e_ok_local else 1.0) == 0.0           # every rank's host result == its device result

    # ---- what the host can deliver: every rank copies its pinned recordings to its GPU at the same time, plain
    #      contiguous cudaMemcpyAsync (the ceiling of any upload scheme), then the same pipeline with contiguous
    #      whole-row uploads (23 % more bytes, one DMA descriptor per chunk) instead of the strided live-sample upload ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        raw.copy_(raw_host, non_blocking=True)
    torch.cuda.synchronize()
    h2d_ceiling = 2 * raw_host.numel() * 4 / max_over_ranks(time.perf_counter() - t0) / 1e9      # GB/s per GPU
    # ... and with the features flowing back at the same time, as they do in the pipeline (0.175 B down per B up)
    side = torch.cuda.Stream()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        raw.copy_(raw_host, non_blocking=True)
        with torch.cuda.stream(side):
            de_host.copy_(de_buf, non_blocking=True)
            psd_host.copy_(psd_buf, non_blocking=True)
    torch.cuda.synchronize()
    h2d_ceiling_duplex = 2 * raw_host.numel() * 4 / max_over_ranks(time.perf_counter() - t0) / 1e9
    e2e_layouts = {}
    for layout_name, layout_compact in (("contiguous_upload", False), ("strided_upload", True)):
        pipe_c = pipeline.HostPipeline(dev, 62, 104000, chunk_blocks=args.chunk_blocks, mode=mode, compact=layout_compact)
        de_host.zero_()
        psd_host.zero_()
        pipe_c.run(raw_host, de_host, psd_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe_c.run(raw_host, de_host, psd_host)
        torch.cuda.synchronize()
        e2e_c_s = max_over_ranks(time.perf_counter() - t0)
        e2e_layouts[layout_name] = {
            "value": world * cw_step_gpu * e2e_steps / e2e_c_s, "unit": UNIT,
            "h2d_bytes_per_step": int(pipe_c.h2d_bytes(S * 7)),
            "h2d_gbs_per_gpu": pipe_c.h2d_bytes(S * 7) / (e2e_c_s / e2e_steps) / 1e9,
            "matches_device_result": max_over_ranks(0.0 if torch.equal(de_host, de_buf.cpu()) else 1.0) == 0.0}
        del pipe_c
    e2e_launches = _lib.launch_count() - launches_e2e0

    log("gather")
    # ---- final gather of the feature tensors to rank 0 (the only exchange; reported, not in `value`) ----
    def check_every_rank(full_de, full_psd, per_rank, pick):
        """rank 0: recompute one subject of EVERY rank's shard locally (subjects are seeded by global id) and compare
        it with what arrived over NVLink.  Returns (all ranks match, ranks checked)."""
        ok = True
        for r in range(world):
            sid = r * per_rank + pick(r)
            blocks = synth.synth_subject(sid, device=dev)
            w_de, w_psd, _ = ops.de_psd_from_raw(blocks, mode_id)
            ok = ok and bool(torch.equal(full_de[sid].reshape(w_de.shape), w_de) and
                             torch.equal(full_psd[sid].reshape(w_psd.shape), w_psd))
        return ok, world

    class ShardKernels:
        """compute hook for cohort.run_cohort: the fused kernel through the C ABI into PREALLOCATED per-rank feature
        buffers (chunk after chunk, in call order), one CUDA event pair around each launch -- so that what the events
        measure is kernels, not the caching allocator."""

        def __init__(self, n_local, into=None):
            """into: (de, psd) tensors whose first n_local subjects receive the features (rank 0 passes its own slice
            of the cohort tensors, so that nothing has to be copied there afterwards)."""
            shape = (n_local, 7 * 200, ops.WINDOWS_PER_CLIP[mode_id], 62, 5)
            if into is None:
                self.de = torch.empty(shape, dtype=torch.float32, device=dev)
                self.psd = torch.empty_like(self.de)
            else:
                self.de, self.psd = (t[:n_local].reshape(shape) for t in into)
            self.events, self.at = [], 0

        def reset(self):
            self.events, self.at = [], 0

        def __call__(self, x):
            n = x.shape[0]
            flat = x.reshape(n * 7, 62, x.shape[-1])
            de, psd = self.de[self.at:self.at + n], self.psd[self.at:self.at + n]
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            _lib.check(lib.eegfe_de_psd_from_raw(flat.data_ptr(), flat.shape[0], 62, flat.shape[2], flat.stride(0),
                                                 flat.stride(1), mode_id, de.data_ptr(), psd.data_ptr(),
                                                 status.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
            b_.record()
            self.events.append((a_, b_))
            self.at += n
            lead = (n, 7, 40, 5) + ((de.shape[2],) if de.shape[2] > 1 else ()) + (62, 5)
            return de.reshape(lead), psd.reshape(lead)

        def kernel_seconds(self):
            return sum(a_.elapsed_time(b_) for a_, b_ in self.events) * 1e-3

    gather = None
    if world > 1:
        raw5 = raw.reshape(S, 7, 62, 104000)
        feat_shape = (S * world, 7, 40, 5) + ((ops.WINDOWS_PER_CLIP[mode_id],) if ops.WINDOWS_PER_CLIP[mode_id] > 1 else ()) + (62, 5)
        out_full = None
        if rank == 0:
            out_full = (torch.empty(feat_shape, dtype=torch.float32, device=dev),
                        torch.empty(feat_shape, dtype=torch.float32, device=dev))
        # (a) baseline: the plain collective -- NCCL gather of DE and of PSD after the kernels
        reps = 3
        flat_out = None if out_full is None else (out_full[0].reshape(S * world, -1), out_full[1].reshape(S * world, -1))
        for i in range(1 + reps):                 # first pass warms up the NCCL channels, untimed
            if i == 1:
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
            cohort.gather_to_rank0(de_buf.reshape((S, -1)), S * world, out=None if flat_out is None else flat_out[0])
            cohort.gather_to_rank0(psd_buf.reshape((S, -1)), S * world, out=None if flat_out is None else flat_out[1])
        g1.record()
        barrier()
        both_ok = check_every_rank(out_full[0], out_full[1], S, lambda r: (3 * r + 1) % S)[0] if rank == 0 else True
        both_ms = max_over_ranks(g0.elapsed_time(g1) / reps)
        if rank == 0:
            out_full[0].zero_()
            out_full[1].zero_()
        # (b) the product path, compute INCLUDED: cohort.process_cohort -- chunked kernels, PSD only over point-to-point
        #     NCCL as each chunk finishes, DE rebuilt on rank 0 (eegfe_de_from_psd)
        chunk = max(1, S // args.gather_chunks)
        kern = ShardKernels(S, into=out_full)
        for i in range(1 + reps):
            if i == 1:
                barrier()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
            kern.reset()
            cohort.process_cohort(raw5, S * world, mode=mode, chunk_subjects=chunk, compute=kern, out=out_full)
        p1.record()
        barrier()
        psd_ms = max_over_ranks(p0.elapsed_time(p1) / reps)
        all_ok, n_checked = (True, world)
        if rank == 0:
            all_ok, n_checked = check_every_rank(out_full[0], out_full[1], S, lambda r: (5 * r + 2) % S)
            all_ok = all_ok and bool(torch.equal(out_full[0][:S].reshape(de_buf.shape), de_buf))
        del out_full, flat_out, kern
        psd_bytes = psd_buf.numel() * 4 * (world - 1)
        kernel_ms_step = elapsed_ms / args.steps
        gather = {
            "path": f"cohort.process_cohort: kernels in {args.gather_chunks} chunks per rank, PSD only over point-to-point "
                    "NCCL as each chunk finishes, DE rebuilt on rank 0 with the kernels' own log2 expression; "
                    "destination tensors preallocated",
            "ms_compute_and_gather": psd_ms, "bytes_into_rank0": psd_bytes,
            "value_with_gather": world * cw_step_gpu / (psd_ms * 1e-3),
            "gbs_into_rank0": psd_bytes / (max(psd_ms - kernel_ms_step / args.gather_chunks, 1e-3) * 1e-3) / 1e9,
            "gbs_note": "bytes into rank 0 / (time - first chunk's kernel); NVLink 5 ingest: 900 GB/s nominal, ~770 GB/s "
                        "measured peer copy (B200_PROFILING.md)",
            "all_ranks_match": bool(all_ok), "ranks_checked": n_checked,
            "check": "rank 0 recomputed one subject of every rank's shard from its seed and torch.equal'ed DE and PSD "
                     "with the gathered slices",
            "collective_de_and_psd": {
                "ms": both_ms, "bytes_into_rank0": 2 * psd_bytes, "gbs_into_rank0": 2 * psd_bytes / (both_ms * 1e-3) / 1e9,
                "value_with_gather": world * cw_step_gpu / ((kernel_ms_step + both_ms) * 1e-3),
                "all_ranks_match": bool(both_ok),
                "note": "round-1 path (NCCL gather of both tensors after the kernels), kept as the comparison"}}

    log("cohort")
    # ---- BASELINE configs[3] as written: the 1000-subject cohort, 500 ms sliding windows, sharded by subject ----
    cohort_big = None
    if args.cohort_subjects > 0:
        total = args.cohort_subjects
        lo_g, hi_g = cohort.shard_bounds(total, rank, world)
        n_local = hi_g - lo_g
        del de_buf, psd_buf
        free_b = torch.cuda.mem_get_info(dev)[0]
        resident = n_local * synth.BYTES_PER_SUBJECT + (n_local + (total if rank == 0 else 0)) * 24.304e6 < 0.8 * free_b
        chunk_c = args.cohort_chunk
        big_shape = (total, 7, 40, 5) + ((ops.WINDOWS_PER_CLIP[mode_id],) if ops.WINDOWS_PER_CLIP[mode_id] > 1 else ()) + (62, 5)
        out_big = None
        if rank == 0:
            out_big = (torch.empty(big_shape, dtype=torch.float32, device=dev),
                       torch.empty(big_shape, dtype=torch.float32, device=dev))
        kern = ShardKernels(n_local, into=out_big)
        if resident:
            big = synth.synth_cohort(range(lo_g, hi_g), dev)
            loader = lambda lo, hi: big[lo:hi]                               # noqa: E731
        else:
            stage = torch.empty((chunk_c, 7, 62, 104000), dtype=torch.float32, device=dev)

            def loader(lo, hi):
                return synth.synth_cohort(range(lo_g + lo, lo_g + hi), dev, out=stage[:hi - lo])
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        c0.record()
        big_de, big_psd = cohort.run_cohort(n_local, loader, total, mode=mode, chunk_subjects=chunk_c, compute=kern,
                                            out=out_big)
        c1.record()
        barrier()
        wall_s = time.perf_counter() - t_wall
        kernel_s = max_over_ranks(kern.kernel_seconds())
        total_s = max_over_ranks(c0.elapsed_time(c1) * 1e-3)
        big_ok = True
        if rank == 0:
            for r in range(world):
                r_lo, r_hi = cohort.shard_bounds(total, r, world)
                sid = r_lo + (7 * r + 3) % max(1, r_hi - r_lo)
                w_de, w_psd = frontend.de_psd_from_raw(synth.synth_subject(sid, device=dev), mode, check=False)
                big_ok = big_ok and bool(torch.equal(big_de[sid], w_de) and torch.equal(big_psd[sid], w_psd))
        cw_total = total * CW_PER_SUBJECT[mode]
        cohort_big = {
            "subjects": total, "subjects_per_gpu": n_local, "chunk_subjects": chunk_c, "channel_windows": cw_total,
            "raw_resident": bool(resident),
            "kernel_seconds": kernel_s, "value_kernels_only": cw_total / kernel_s,
            "seconds_compute_and_gather": total_s if resident else None,
            "value_with_gather": cw_total / total_s if resident else None,
            "wall_seconds_including_synthesis": wall_s,
            "features_on_rank0_gb": total * 24.304e6 / 1e9,
            "all_ranks_match": bool(big_ok), "ranks_checked": world,
            "note": ("recordings resident in HBM before the timed region; cohort.run_cohort = kernels per chunk + PSD-only "
                     "point-to-point gather + DE rebuilt on rank 0" if resident else
                     "180 GB of recordings do not fit one GPU: synthesised chunk by chunk inside the loop, so only the "
                     "kernels are timed (one CUDA event pair per launch); there is no gather at 1 GPU")}
        del big_de, big_psd, out_big

    log("cpu baseline, report")
    if rank == 0:
        peak, peak_src = measured_peaks()
        kernel_ms = elapsed_ms / args.steps                       # one launch per step
        achieved = cw_step_gpu * BYTES_PER_CW[mode] / (kernel_ms * 1e-3) / 1e9
        traffic = profiled_traffic(mode)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None if traffic is None else traffic["dram_bytes_per_cw"] * cw_step_gpu,
                    "peak_source": peak_src, "kernel": KERNEL_NAME[mode], "kernel_ms": kernel_ms,
                    "algorithmic_bytes_per_channel_window": BYTES_PER_CW[mode],
                    "note": "500 ms mode is bounded by the FP32 pipe, not HBM (DESIGN.md 'Rooflines'); see 'fp32_pipe'"
                    if mode == "500ms" else ""}
        # the CUDA-core FP32 pipe is what actually bounds the 500 ms kernel: 128 lanes/clk/SM at the clock seen under load
        # numerator and denominator from the SAME window: the sustained run and the SM clock sampled during it
        if sustained is not None and sustained["clocks"].get("sm_mhz"):
            sm_mhz, fp_ms, fp_window = sustained["clocks"]["sm_mhz"], sustained["ms_per_step"], "sustained run"
        else:
            sm_mhz, fp_ms, fp_window = 1965.0, kernel_ms, "timed region at the maximum SM clock (no clock samples)"
        fp_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
        fp_ach = cw_step_gpu * FP32_LANE_OPS_PER_CW[mode] / (fp_ms * 1e-3) / 1e12
        fp32_pipe = {"achieved": fp_ach, "peak": fp_peak, "unit": "T fp32 lane-op/s (an FMA counts once)",
                     "frac": fp_ach / fp_peak, "lane_ops_per_channel_window": FP32_LANE_OPS_PER_CW[mode],
                     "window": fp_window,
                     "peak_source": f"148 SMs x 128 lanes x {sm_mhz:.0f} MHz (median SM clock sampled over the same window)"}
        # the bound that actually binds the 500 ms kernel: the sub-partition's issue port -- a packed f32x2
        # instruction holds it for two cycles, anything else for one.  Instruction counts per window-warp come from the
        # committed ncu capture (profiles/ncu_r02_500ms.txt: smsp__inst_executed.sum / window-warps); cycles from this run.
        issue_model = None
        if mode == "500ms":
            inst, packed = 1990.0, 982.0
            cycles = fp_ms * 1e-3 * sm_mhz * 1e6 * 148 * 4 / (cw_step_gpu / 32.0)
            issue_model = {"instructions_per_window_warp": inst, "packed_f32x2_per_window_warp": packed,
                           "issue_cycles_per_window_warp": inst + packed, "measured_cycles_per_window_warp": cycles,
                           "frac": (inst + packed) / cycles, "window": fp_window,
                           "note": "issue-port occupancy of the 500 ms kernel; the arithmetic core alone needs 2539 cycles"}
        if sustained is not None:
            sus_gbs = sustained["value"] / world * BYTES_PER_CW[mode] / 1e9
            sustained["hbm_gbs_per_gpu"] = sus_gbs
            sustained["hbm_frac"] = sus_gbs / peak
        cpu = None
        if not args.skip_cpu_baseline and world == 1:
            arm = CpuArm(mode)
            cw, s = arm.run(args.cpu_clips_per_step * 24)
            arm.close()
            cpu = {"value": cw / s, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                   "sample": f"{arm.cores} workers x {args.cpu_clips_per_step * 24} synthetic 2 s clips (62 ch) each, "
                             f"mode {mode}, {arm.what}, {s:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, _lib.launch_geometry(mode_id)),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes(S * 7)),
                    "d2h_bytes_per_step": int(2 * de_host.numel() * 4), "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "matches_device_result": e2e_ok,
                    "h2d_gbs_per_gpu": pipe.h2d_bytes(S * 7) / (e2e_s / e2e_steps) / 1e9,
                    "h2d_ceiling_gbs_per_gpu": h2d_ceiling, "h2d_ceiling_gbs_aggregate": h2d_ceiling * world,
                    "h2d_frac_of_ceiling": pipe.h2d_bytes(S * 7) / (e2e_s / e2e_steps) / 1e9 / h2d_ceiling,
                    "h2d_ceiling_with_d2h_gbs_per_gpu": h2d_ceiling_duplex,
                    "h2d_frac_of_ceiling_with_d2h": pipe.h2d_bytes(S * 7) / (e2e_s / e2e_steps) / 1e9 / h2d_ceiling_duplex,
                    "h2d_ceiling_how": "all ranks at once: 2 x contiguous pinned cudaMemcpyAsync of the resident batch, "
                                       "max over ranks",
                    "upload_layout": "strided: live samples only" if pipe.compact else "contiguous: whole rows",
                    "upload_probe": pipe.upload_probe,
                    "contiguous_upload": e2e_layouts["contiguous_upload"],
                    "strided_upload": e2e_layouts["strided_upload"],
                    "bound": "host-to-device copy (PCIe Gen5 x16, ~55 GB/s per GPU in practice for a contiguous copy)",
                    "rank0_numa_cpus": None if numa_cpus is None else len(numa_cpus),
                    "path": "pinned host recordings -> HostPipeline (chunked H2D in the layout its probe picked / fused kernel "
                            "/ D2H of DE+PSD, 3 streams) -> pinned host features"},
            "gpu_launches": int(gpu_launches), "gpu_launches_e2e": int(e2e_launches),
            "roofline": roofline, "fp32_pipe": fp32_pipe, "issue_model": issue_model, "single_subject": single, "other_modes": other_modes, "next_rows": next_rows,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        line["value_sustained"] = None if sustained is None else sustained["value"]
        line["sustained"] = sustained
        if gather:
            line["gather"] = gather
        if cohort_big:
            line["cohort_1000" if args.cohort_subjects == 1000 else f"cohort_{args.cohort_subjects}"] = cohort_big
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--mode", default="500ms", choices=("500ms", "1s", "2s"))
    ap.add_argument("--subjects", type=int, default=24, help="subjects resident per GPU (one step = all of them)")
    ap.add_argument("--chunk-blocks", type=int, default=28, help="blocks per in-flight chunk of the e2e pipeline")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-clips-per-step", type=int, default=40, help="clips per worker per CPU step")
    ap.add_argument("--sustain-s", "--clock-probe-s", dest="sustain_s", type=float, default=2.0,
                    help="seconds of back-to-back launches after the timed steps (sustained throughput + clocks); 0 = off")
    ap.add_argument("--gather-chunks", type=int, default=8, help="kernel chunks per rank in the gather measurement")
    ap.add_argument("--cohort-subjects", type=int, default=1000, help="BASELINE configs[3] cohort size; 0 = skip")
    ap.add_argument("--cohort-chunk", type=int, default=25, help="subjects per kernel launch in the cohort run")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-other-modes", action="store_true")
    ap.add_argument("--tensor-loads", action="store_true",
                    help="measurement option: 2 s tiles by one TMA tensor copy each (eegfe_set_tensor_loads)")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
