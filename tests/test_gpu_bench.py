"""The GPU arm of bench.py keeps the driver's contract: one JSON line with the base keys, `roofline`, `e2e` with the bytes it
copies, `gpu_launches`, `clocks`, and a result that is bit-identical between the resident and the host-buffer call."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_gpu_arm_contract():
    env = dict(os.environ)
    for key in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "VS_FUSED_VARIANT", "VS_NCCL"):
        env.pop(key, None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu"], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["unit"] == "evals/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    k, n = d["config"]["k"], d["config"]["n"]
    assert (k, n) == (20, 1 << 24) and d["config"]["evals_per_step"] == 2 * n * (1 + k)
    assert abs(d["value"] - d["config"]["evals_per_step"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0.3 < r["frac"] < 1.0 and 0.3 < r["frac_executed"] < r["frac"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["algorithmic_flops_per_launch"] == 6008 * n
    assert r["kernel_ms"] <= d["ms_per_step"] * 1.001
    assert r["traffic"] is None or r["traffic"] >= 4 * n                       # the uint32 permutation at least
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4 * n and e["d2h_bytes_per_step"] > 0 and e["unit"] == "evals/s"
    assert e["value"] < d["value"] * 1.02                                      # the host-buffer call cannot beat the resident one
    assert d["gpu_launches"] >= d["steps"]
    assert d["check"]["e2e_matches_resident_bitwise"] is True
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
