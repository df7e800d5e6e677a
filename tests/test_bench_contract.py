"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run_bench(*args, timeout=900):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "reads/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_native_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--reads", "20000")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "parity"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    # default plan on 4096-sample reads: stats (+ its redo list) + lower-bound scan + windows, finalize, second-attempt windows,
    # finalize, full-length fallback per step
    assert d["plan"]["name"] == "two_pass" and r["kernel"] == "sqk_dtw_lb_kernel"
    assert r["kernel_ms_per_launch"] > 0 and d["gpu_launches"] == 8 * d["steps"]
    assert d["plan"]["full_length_fallback_reads_per_step"] <= 0.01 * 20000
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 20000 * 4096 * 2 + 20001 * 8
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0
    assert d["parity"]["indices_bit_exact"] and d["parity"]["dist_bit_exact"] and d["parity"]["e2e_matches_device_path"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
