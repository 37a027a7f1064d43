#!/usr/bin/env python
"""A/B of the refill policy of the render queue (RenderJob::tile_mode): per-lane refill vs whole 32-pixel tiles per warp.

    python tools/ab_tile.py [--out profiles/x.json]

Frames: 400x400 (BASELINE config 1), 1920x1080 (config 3), one rank's shard of the 4K frame split 8 ways (rows 0::8) and
the whole 4K frame.  Reports kernel ms, lane utilisation (device counters) and that every output bit is the same.
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
ap.add_argument("--out", default="")
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
hd = schw_b200.Handle(0)
rng = np.random.default_rng(0)
res = []
for (h, w, stride) in ((400, 400, 1), (1080, 1920, 1), (2160, 3840, 8), (2160, 3840, 4), (2160, 3840, 2), (2160, 3840, 1)):
    tex = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    sc = schw_b200.make_scene(height=h, width=w)
    row = {"height": h, "width": w, "rows": f"0::{stride}"}
    outs = {}
    for mode, rays in (("0", "1"), ("1", "1"), ("2", "1"), ("2", "2")):
        os.environ["SCHW_TILE_MODE"] = mode
        os.environ["SCHW_TILE_RAYS"] = rays
        mode = {"0": "lane_refill", "1": "tiles_in_lane_kernel"}.get(mode, f"tile_kernel_r{rays}")
        best = 1e30
        for _ in range(a.reps):
            out = hd.render(sc, tex, 0, h, stride, extras=(h <= 1080))
            best = min(best, hd.last_kernel_ms()[0])
        steps, warp_iters, planar = hd.last_counters()
        outs[mode] = out
        row[f"{mode}_ms"] = best
        row[f"{mode}_lane_utilisation"] = steps / (32.0 * warp_iters)
        row["steps"] = steps
    row["identical"] = all(np.array_equal(outs["lane_refill"][k], o[k], equal_nan=(k == "final_q")) for o in outs.values() for k in o)
    row["speedup_tile_kernel_r2_over_lane_refill"] = row["lane_refill_ms"] / row["tile_kernel_r2_ms"]
    res.append(row)
    print(json.dumps(row), flush=True)
os.environ.pop("SCHW_TILE_MODE", None); os.environ.pop("SCHW_TILE_RAYS", None)
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
