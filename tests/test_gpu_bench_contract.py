"""bench.py's output contract on a real GPU: ONE JSON line on stdout with the keys the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", *extra],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_line_small_batch():
    d = _run("--batch", "32", "--no-cpu-baseline")
    assert d["metric"].startswith("frame-pairs/sec") and d["unit"] == "frame-pairs/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 32 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] == d["launches_per_step"] * 3 and d["launches_per_step"] >= 30
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.02 and e["h2d_bytes_per_step"] == 4 * 32 * 3 * 256 * 4 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and 0 < r["frac_of_ceiling"] < 1.2 and r["hbm"]["peak_gbs"] > 1000
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert abs(sum(k["share"] for k in d["kernels"].values()) - 1.0) < 1e-6
    assert d["config"]["workload"].startswith("CMFlow forward") and "model" not in d["config"]


def test_bench_line_other_models():
    for model in ("cmflow_t", "raflow"):
        d = _run("--batch", "16", "--no-cpu-baseline", "--model", model)
        assert d["value"] > 0 and d["e2e"]["value"] > 0
