"""bench.py's CPU arm obeys the output contract (one JSON line on stdout, the keys the driver reads) -- CPU only.
The GPU arm needs a B200; its line is checked by hand against the same key list in profiles/bench_r02_n1.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-clips-per-step", "1"], capture_output=True, text=True, timeout=300,
                         env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "de_psd_channel_windows_per_s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_lines_carry_the_contract_keys():
    """The GPU arm's lines recorded on B200s in this round (profiles/) have every key the task's contract names."""
    need = BASE_KEYS | {"clocks", "gpu_launches", "roofline", "value_sustained", "parity"}
    for name in ("bench_r02_n1.json", "bench_r02_n8.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            line = json.load(f)
        assert need <= set(line), sorted(need - set(line))
        roof = line["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof) and roof["bound"] == "hbm"
        assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0
        if line["n_gpus"] > 1:
            assert line["gather"]["all_ranks_match"] is True and line["cohort_1000"]["all_ranks_match"] is True
