"""CPU: the reference arm of bench.py (the CPU port of the reference callbacks on the host cores) runs without a GPU
and prints the JSON line the driver expects."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "crossvault_disc10",
                          "--steps", "1", "--warmup", "1", "--cpu-evals-per-core", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["config"]["workload"] == "crossvault_disc10"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_prints_the_contract_line():
    """`python bench.py` (small batch) on the GPU: every key the driver reads, roofline and e2e objects included."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--batch", "2048",
                          "--e2e-batch", "512", "--no-cpu-baseline", "--no-sweeps"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["unit"] == "evals/s" and line["dtype"] == "f64" and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["config"]["workload"] == "crossvault_disc20"
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert line["gpu_launches"] == 2 * 3                      # k_fill_funicular + k_onchip per step
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"]
    assert line["value"] > 1e5 and line["status_nonzero"] == 0
