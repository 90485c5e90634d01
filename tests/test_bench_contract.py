"""bench.py prints ONE JSON line with the keys the driver reads (bench contract of the task): the B200 arm on a small
workload (GPU), the reference arm on the host cores (CPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    b = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 600)
    assert b["impl"] == "reference" and b["unit"] == "env-steps/s" and b["value"] > 0 and b["higher_is_better"] is True
    assert b["cpu_baseline"]["kind"] in ("reference", "port") and b["cpu_baseline"]["cores"] >= 1 and b["cpu_baseline"]["value"] == b["value"]
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert b["gpu_launches"] == 0 and "workload" in b["config"]


@pytest.mark.gpu
def test_b200_arm_line():
    n, k = 1 << 16, 3
    b = _run(["--envs", str(n), "--ring", str(1 << 18), "--steps", str(k), "--warmup", "3", "--preroll", "20", "--no-cpu-baseline",
              "--e2e-preroll-s", "0.05"], 900)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "kernels", "config4", "sweep"):
        assert key in b, key
    assert b["n_gpus"] == 1 and b["steps"] == k and b["scaling"] == "weak" and b["data"] == "synthetic" and b["dtype"] == "f16"
    assert abs(b["value"] - n * k / (b["ms_per_step"] * k * 1e-3)) < 1e-6 * b["value"]
    assert b["gpu_launches"] == 2 * k                                   # two kernels per rollout iteration
    e = b["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 131601 * 4 and e["d2h_bytes_per_step"] == 4 * n + 4 * (n // 32) + 128
    r = b["roofline"]
    assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(b["kernels"]) == {"actor", "env_step"} and "workload" in b["config"] and "model" not in b["config"]
    assert b["config4"]["ms_per_step"] > 0 and [s["envs_per_gpu"] for s in b["sweep"]][:2] == [4096, 262144]
