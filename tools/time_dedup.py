#!/usr/bin/env python
"""Frame time of the opt-in de-duplicated render next to the one-ray-per-pixel render.

    python tools/time_dedup.py [--height 2160 --width 3840 --reps 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
import schw_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
os.environ["SCHW_DEDUP_REUSE"] = "0"          # every timed frame plans and integrates again
hd = schw_b200.Handle(0)
tex = np.random.default_rng(0).integers(0, 256, size=(a.height, a.width, 3), dtype=np.uint8)
sc = schw_b200.make_scene(height=a.height, width=a.width)
res = {}
for dedup in (False, True):
    best = 1e30
    for _ in range(a.reps):
        out = hd.render(sc, tex, dedup=dedup)
        best = min(best, hd.last_kernel_ms()[0])
    res[dedup] = (best, hd.last_total_steps(), out)
uniq, pixel_steps = hd.last_dedup_stats()
same = np.array_equal(res[False][2]["rgb"], res[True][2]["rgb"]) and np.array_equal(res[False][2]["cls"], res[True][2]["cls"])
print(json.dumps({"frame": f"{a.width}x{a.height}", "one_ray_per_pixel_ms": res[False][0], "dedup_ms": res[True][0],
                  "pixels": a.height * a.width, "distinct_rays": uniq, "steps_per_frame": res[False][1],
                  "steps_integrated_dedup": res[True][1], "pixel_steps_dedup": pixel_steps,
                  "speedup": res[False][0] / res[True][0], "identical_rgb_cls": bool(same)}))
