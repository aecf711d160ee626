"""bench.py's contract, as far as a machine without a GPU can check it: the reference arm (`--impl reference`) times the
reference's own object code (oracle/_ref) or the port on the host cores and prints ONE JSON line with the keys the driver
reads; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(extra_env=None):
    env = dict(os.environ, RMG_BENCH_CPU_SAMPLE="8000")
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_line_with_the_drivers_keys():
    lines = run_reference()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "voxels/s" and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "V=7076160" in d["config"]["workload"] and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "voxels" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    assert run_reference({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
