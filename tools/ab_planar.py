#!/usr/bin/env python
"""A/B of the planar fast path on one frame: kernel time with SCHW_PLANAR=1 / 0 and a bitwise comparison of every output.

    python tools/ab_planar.py [--height 2160 --width 3840 --reps 3 --delta 0.01]
"""
import argparse
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
ap.add_argument("--delta", type=float, default=0.01)
a = ap.parse_args()

hd = schw_b200.Handle(0)
tex = np.random.default_rng(0).integers(0, 256, size=(a.height, a.width, 3), dtype=np.uint8)
sc = schw_b200.make_scene(height=a.height, width=a.width, delta=a.delta)
res = {}
for planar in ("1", "0", "1", "0"):
    os.environ["SCHW_PLANAR"] = planar
    best = 1e30
    for r in range(a.reps):
        out = hd.render(sc, tex, extras=(r == 0))
        ms, _ = hd.last_kernel_ms()
        best = min(best, ms)
        if r == 0:
            keep = out
    steps, iters, pl = hd.last_counters()
    print(f"SCHW_PLANAR={planar}: {a.width}x{a.height} kernel {best:.3f} ms, {steps} steps, {steps / best * 1e-6:.2f} Gsteps/s, "
          f"lane utilisation {steps / (32.0 * iters):.4f}, planar iterations {pl / iters:.5f}")
    res[planar] = keep
same = all(np.array_equal(np.ascontiguousarray(res["1"][k]).view(np.uint8), np.ascontiguousarray(res["0"][k]).view(np.uint8))
           for k in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"))
print("bitwise identical outputs (rgb, cls, final_q, ic, hit, uv, nsteps):", same)
sys.exit(0 if same else 1)
