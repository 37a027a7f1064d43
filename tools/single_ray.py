#!/usr/bin/env python
"""BASELINE config 2 driver: the inputs of the reference's single_ray_cuda_test.py:250-321 through the drop-in
CUDASchwarzschildIntegrator; writes single_ray_cuda_test.csv with the reference's columns (t, r, theta[deg], phi[deg])
and, with --golden, compares against tests/golden/single_ray_k3.npz (the reference's own output).

    python tools/single_ray.py [--out single_ray_cuda_test.csv] [--golden]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
from simulation.cuda_geodesic import CUDASchwarzschildIntegrator  # noqa: E402
from simulation.utils import build_null_4momentum_ep_sph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="single_ray_cuda_test.csv")
ap.add_argument("--golden", action="store_true")
a = ap.parse_args()

mass_bh, r_max = 1.0, 50
steps, delta, omega = 200_000, 0.03, 0.01
pos_sph = np.array([35.0, np.pi / 2, 0.0])
p_direction = np.array([-0.026942690335328513, -0.028502831807219468, 0.06898831276132347])
# the script's local builder returns the NEGATIVE root for future=True (single_ray_cuda_test.py:219-222)
p0 = build_null_4momentum_ep_sph(p_direction, pos_sph, mass_bh=mass_bh, future=False)
q0 = np.array([0.0, *pos_sph])
integrator = CUDASchwarzschildIntegrator(steps=steps, delta=delta, mass=mass_bh, omega=omega, r_max=r_max)
traj = integrator.integrate_batch_full(q0[np.newaxis, :], p0[np.newaxis, :])[0]
safe = traj[:, 1] > 1.1 * 2.0 * mass_bh
if not np.all(safe):
    traj = traj[: np.argmax(~safe)]
print(f"Safe trajectory length: {len(traj)} steps")
import pandas as pd  # noqa: E402
df = pd.DataFrame(traj, columns=["t", "r", "theta", "phi"])
df["theta"], df["phi"] = np.degrees(df["theta"]), np.degrees(df["phi"])
df.to_csv(a.out, index=False)
print(f"wrote {a.out} ({os.path.getsize(a.out)} bytes; the reference's LFS pointer says 495057)")
if a.golden:
    g = np.load(os.path.join(ROOT, "tests", "golden", "single_ray_k3.npz"))["traj"]
    n = min(len(traj), len(g) - 1)
    err = np.abs(traj[:n] - g[:n]) / np.maximum(np.abs(g[:n]), 1.0)
    for lo, hi in ((0, 2500), (2500, 4500), (4500, 6000), (6000, n)):
        print(f"steps {lo:5d}..{hi:5d}: max rel. deviation from the reference {err[lo:hi].max():.2e}")
