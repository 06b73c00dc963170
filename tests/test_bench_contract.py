"""bench.py prints exactly ONE JSON line with the keys the driver reads (both arms), on tiny overrides of the workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    j = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "1500", "--draws", "40")
    assert BASE <= set(j) and j["impl"] == "reference"
    assert j["unit"] == "tree-evals/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


@pytest.mark.gpu
def test_our_arm_line():
    j = _run("--steps", "3", "--warmup", "3", "--points", "300000", "--draws", "64", "--cpu-budget", "1")
    assert BASE | {"clocks", "gpu_launches", "roofline", "roofline_k0", "strong"} <= set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 3 and j["warmup"] == 3 and j["dtype"] == "f64" and j["scaling"] == "weak"
    assert j["gpu_launches"] > 0 and j["value"] > 0
    r = j["roofline"]
    # the dominant kernel is bound by the shared-memory pipe, not HBM; the HBM figures ride along
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "hbm"} <= set(r) and r["bound"] == "smem"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] <= 1.0
    assert r["hbm"]["unit"] == "GB/s" and abs(r["hbm"]["frac"] - r["hbm"]["achieved"] / r["hbm"]["peak"]) < 1e-12
    k0 = j["roofline_k0"]
    assert k0["bound"] == "hbm" and abs(k0["frac"] - k0["achieved"] / k0["peak"]) < 1e-12
    assert j["strong"]["cfg2_fixed_1e6_points"]["scaling"] == "strong"
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 300000 * 10 * 8 and e["d2h_bytes_per_step"] > 0
    c = j["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and isinstance(c["sample"], str)
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"])


def test_reference_arm_uses_all_cores_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm (rank 0 only) must still use the host."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--points", "1500", "--draws", "40"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and j["n_gpus"] == 2
    # the other ranks print nothing and exit 0
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--points", "1500", "--draws", "40"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
