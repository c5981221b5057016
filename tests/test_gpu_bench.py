"""bench.py on the GPU: the one-line JSON carries every key of the contract (structure only, no thresholds)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_arm_json_contract():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "300", "--warmup", "3", "--mix", "100",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 300 and d["scaling"] == "weak" and d["dtype"] == "u8"
    assert d["config"]["workload"].startswith("C2") and "model" not in d["config"]
    import argparse
    import bench
    assert d["config"] == bench.static_config(argparse.Namespace(workload="c2", envs_per_gpu=None, hardness="none"), 1)
    rs = d["run_stats"]
    # a 300-step block is ~10 ms: the region is repeated until it lasts >= 0.2 s, the median block is reported
    assert rs["blocks"] >= 5 and rs["timed_steps_total"] == 300 * rs["blocks"] and rs["timed_region_ms"] >= 150
    assert rs["ms_per_step_min"] <= d["ms_per_step"] <= rs["ms_per_step_max"]
    assert d["gpu_launches"] == rs["launches_per_step"] * 300 * rs["blocks"] and d["value"] > 0
    assert d["clocks"]["samples"] >= 5
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or (r["traffic"] > 0 and r["dram_frac"] > 0)
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4 * 4096 and e["d2h_bytes_per_step"] == 20 * 4096
    assert {"numpy_obs", "aux5_scaled_float", "aux5_uint8"} <= set(e["variants"])
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert {"rgb_only", "float_chw", "a2c_pass", "reference_run_config"} <= set(d["secondary"])
