"""The bench.py contract (one JSON line on stdout, the keys the driver reads): the CPU reference arm runs
everywhere; the CUDA arm is checked on the GPU box with a small batch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-clouds", "2"])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("points/sec through 4-layer GridConv @ N=8192, K=64")
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_cuda_arm_prints_one_json_line(cuda_dev):
    d = _run(["--steps", "3", "--warmup", "3", "--batch", "8", "--cpu-clouds", "1"])
    assert BASE_KEYS | {"roofline", "roofline_hbm", "clocks", "breakdown_ms"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 3 * 15 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 8192 * 16 + 8 * 4 and d["e2e"]["d2h_bytes_per_step"] > 0
    for r, bound in ((d["roofline"], "tensor"), (d["roofline_hbm"], "hbm")):
        assert r["bound"] == bound and r["achieved"] > 0 and r["peak"] > 0
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
