"""bench.py's reference arm (`--impl reference`) runs on host cores only, so its JSON contract is checked here,
without a GPU: one line, the keys the driver reads, the same metric / unit / config as the product arm, and a
cpu_baseline block that describes the run.  (The product arm needs a B200; its line is checked by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    base = ["--steps", "1", "--warmup", "1", "--full-scan", "0"]
    for flag in ("--steps", "--warmup", "--full-scan"):
        if flag in extra:
            i = base.index(flag)
            del base[i:i + 2]
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *base, "--cpu-rows", "20000", "--cpu-queries", "2", *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference"
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    cfg = d["config"]
    assert "10000000x768" in cfg["workload"] and cfg["k"] == 10 and cfg["batch"] == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "20000 rows" in cb["sample"] and cb["unit"] == d["unit"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["single_thread"]["cores"] == 1 and d["single_thread"]["value"] > 0


def test_reference_arm_honours_steps_and_validates_the_extrapolation():
    d = _run("--rows", "60000", "--full-scan", "1", "--steps", "3", "--warmup", "2")
    assert d["steps"] == 3 and d["warmup"] == 2
    v = d["full_scan_validation"]
    assert v["rows"] == 60000 and 0.2 < v["ratio_measured_over_extrapolated"] < 5.0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--cpu-rows", "20000", "--cpu-queries", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_follows_workload_flags():
    d = _run("--workload", "c3")
    assert d["config"]["k"] == 100 and "top-100" in d["config"]["workload"]
