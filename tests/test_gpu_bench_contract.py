"""The bench.py JSON contract, exercised on a small frame (both arms)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_bench_line_small_frame():
    d = _run(["--steps", "2", "--warmup", "1", "--height", "270", "--width", "480"])
    for k in REQUIRED + ["roofline", "cpu_baseline", "clocks"]:
        assert k in d, k
    assert d["metric"] == "geodesic_steps_per_s" and d["unit"] == "steps/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["gpu_launches"] == 2 and d["value"] > 1e9 and d["e2e"]["value"] > 1e9
    assert d["e2e"]["h2d_bytes_per_step"] == 270 * 480 * 3 and d["e2e"]["d2h_bytes_per_step"] == 270 * 480 * 4
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 20 < r["peak"] < 45 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # frac is a utilisation of the FP64 pipe computed in the run (loop bodies x FP64 instructions per body from the SASS of
    # this build), never above 1; the algorithmic figure lives in its own fields
    assert 0.3 < r["frac"] <= 1.0 and r["fp64_instr_per_loop_body"]["fast_path"] == 100 and r["algorithmic_flop_per_step"] == 312
    assert r["algorithmic_tflops"] > r["achieved"] * 0.9 and 0.9 < r["lane_utilisation"] <= 1.0
    assert "gather_check" not in d                                  # single GPU: nothing gathered
    oc = d["other_configs"]
    assert oc["config1_400x400"]["steps"] > 8e8 and oc["config3_1920x1080"]["frame_ms"] > oc["config1_400x400"]["frame_ms"]
    assert len(d["entry_point"]) == 2 and d["entry_point"][0]["nonblack_pixels"] > 100000
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 1e5 and "rows" in c["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    # the opt-in de-duplicated render is reported beside the headline, never in it
    dd = d["dedup_opt_in"]
    assert dd["identical_to_one_ray_per_pixel"] is True and 0 < dd["distinct_rays"] < dd["pixels"] / 2
    assert dd["pixel_steps"] == d["steps_per_frame"] and dd["steps_integrated"] < dd["pixel_steps"] / 2


def test_bench_reference_arm_small_frame():
    """`--impl reference` = the oracle port on the host cores (CPU-only; also runs in the build container)."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--height", "54", "--width", "96"])
    for k in REQUIRED + ["cpu_baseline", "impl"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
