"""bench.py's reference arm (`--impl reference`: the CPU restatement of the reference loops timed on the host cores) runs without a
GPU; its one JSON line must carry the keys of the measurement contract.  The GPU arm is exercised on the B200 box by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv, env=None):
    e = dict(os.environ, **(env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)
    return r


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--workload", "gx3", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value = cells * ndte / seconds per step
    assert abs(d["value"] - 100 * 116 * 120 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_reference_arm_under_torchrun_only_rank0_works():
    """N > 1: rank 0 alone runs and prints the line, the other ranks exit 0 without work."""
    r = run_bench("--impl", "reference", "--workload", "gx3", "--steps", "1", "--warmup", "1", "--gpus", "2",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
