#!/usr/bin/env python
"""Render one frame through the C ABI (for ncu captures and quick timings).

    python tools/profile_render.py [--height H --width W --reps R --delta D] [--batch N]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
import schw_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=540)
ap.add_argument("--width", type=int, default=960)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--delta", type=float, default=0.01)
ap.add_argument("--stride", type=int, default=1, help="render rows 0::stride only (one rank's shard of a stride-way split)")
ap.add_argument("--batch", type=int, default=0, help="also time integrate_batch on N random non-equatorial rays")
a = ap.parse_args()

hd = schw_b200.Handle(0)
rng = np.random.default_rng(0)
tex = rng.integers(0, 256, size=(a.height, a.width, 3), dtype=np.uint8)
sc = schw_b200.make_scene(height=a.height, width=a.width, delta=a.delta)
for r in range(a.reps):
    out = hd.render(sc, tex, 0, a.height, a.stride)
    ms, n = hd.last_kernel_ms()
    steps = hd.last_total_steps()
    print(f"render {a.width}x{a.height} rows 0::{a.stride}: kernel {ms:.3f} ms, {steps} steps, {steps / ms * 1e-6:.3f} Gsteps/s, "
          f"{312 * steps / ms * 1e-9:.2f} algorithmic TFLOP/s")
if a.batch:
    n = a.batch
    q0 = np.zeros((n, 4)); p0 = np.zeros((n, 4))
    r0 = rng.uniform(8, 40, n); th0 = rng.uniform(0.35, np.pi - 0.35, n)
    al = np.deg2rad(rng.uniform(0, 55, n)); be = rng.uniform(-np.pi, np.pi, n)
    pr = -np.cos(al) * np.sqrt(1 - 2 / r0); pth = np.sin(al) * np.sin(be) * r0; pph = np.sin(al) * np.cos(be) * r0 * np.sin(th0)
    f = 1 - 2 / r0
    C = f * pr * pr + pth * pth / r0 ** 2 + pph * pph / (r0 * r0 * np.sin(th0) ** 2)
    q0[:, 1] = r0; q0[:, 2] = th0; p0[:, 0] = np.sqrt(4 * C / f) / (2 / f); p0[:, 1] = pr; p0[:, 2] = pth; p0[:, 3] = pph
    for r in range(a.reps):
        hd.integrate_batch(q0, p0, 200000, a.delta, 2.0, 45.0, 1.0)
        ms, _ = hd.last_kernel_ms(); steps = hd.last_total_steps()
        print(f"integrate_batch {n} general rays: kernel {ms:.3f} ms, {steps} steps, {steps / ms * 1e-6:.3f} Gsteps/s")
t, ms = hd.fp64_peak_probe()
print(f"fp64 DFMA peak probe: {t:.2f} TFLOP/s ({ms:.3f} ms)")
