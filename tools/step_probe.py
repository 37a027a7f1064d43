#!/usr/bin/env python
"""SM cycles per warp-step of the render kernel's loop body alone, for 1 and 2 rays per thread (schw_step_probe).

    python tools/step_probe.py [--iters 4000] [--out profiles/x.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
import schw_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30000)
ap.add_argument("--out", default="")
a = ap.parse_args()
hd = schw_b200.Handle(0)
info = hd.device_info()
f = info["clock_khz"] * 1e3
res = []
hd.fp64_peak_probe()            # warm-up: the SM clock has to be up before millisecond-long kernels are timed
hd.step_probe(106, a.iters)
for v in (106, 105, 104, 103, 102, 101, 204, 203, 202, 201, 106):
    r = hd.step_probe(v, a.iters)
    # an SM sub-partition hosts a quarter of the SM's warps: cycles it spends per warp-step of ONE ray
    warps_per_smsp = r["ctas_per_sm"] * 4 / 4.0
    cyc = r["ms"] * 1e-3 * f * 4 * info["sm_count"] / (r["warp_steps"] * r["rays_per_thread"])
    r.update(sm_cycles_per_warp_ray_step=cyc, fp64_pipe_utilisation=200.0 / cyc, warps_per_scheduler=warps_per_smsp)
    res.append(r)
    print(json.dumps(r))
if a.out:
    json.dump({"clock_hz": f, "note": "loop body only (100 FP64-pipe instructions per ray-step on the planar fast path = 200 pipe cycles "
                                      "per warp and ray-step); no queue, refill, epilogue or drain", "variants": res}, open(a.out, "w"), indent=1)
