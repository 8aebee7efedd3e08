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
    assert d["gpu_launches"] == d["launches_per_step"] * 3 and d["launches_per_step"] >= 20
    for e in (d["e2e"], d["e2e_serial"]):
        assert 0 < e["value"] <= d["value"] * 1.10 and e["h2d_bytes_per_step"] == 4 * 32 * 3 * 256 * 4 and e["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] >= d["e2e_serial"]["value"] * 0.95                    # pipelined copies never cost throughput
    s = d["sustained"]
    assert s["seconds"] >= 5.0 and s["value"] > 0 and set(s["clocks"]) >= {"sm_mhz", "reasons"}
    lb = d["latency_b1"]
    assert lb["pairs"] == 1 and 0 < lb["device"]["p50_ms"] <= lb["host_eager"]["p50_ms"] and lb["host_graph"]["p50_ms"] > 0 and lb["graphs_cached"] >= 1
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and 0 < r["frac_of_ceiling"] < 1.2 and r["hbm"]["peak_gbs"] > 1000
    assert r["category"] in d["kernels"] and 0 < r["hbm"]["stage"]["algorithmic_frac"] < 1 and 0 < r["whole_step"]["frac"] < 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert abs(sum(k["share"] for k in d["kernels"].values()) - 1.0) < 1e-6
    assert d["config"]["workload"].startswith("CMFlow forward") and "model" not in d["config"]


def test_bench_line_other_models():
    for model in ("cmflow_t", "raflow"):
        d = _run("--batch", "16", "--no-cpu-baseline", "--no-extra-legs", "--model", model)
        assert d["value"] > 0 and d["e2e"]["value"] > 0


def test_reference_arm_line():
    """--impl reference: the unmodified reference Python on the host cores (staged under oracle/_ref/py), same config dict as our arm."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["pairs_per_gpu"] == 256
