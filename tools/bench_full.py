#!/usr/bin/env python
"""integrate_batch_full at scale (32 B per step per ray => <= 2.0e11 steps/s at 6543 GB/s of HBM writes).

    python tools/bench_full.py [--rays 262144 --steps 512 --delta 0.01 --kind camera|general] [--out profiles/x.json]

The trajectory buffer (rays x steps x 32 B) lives in HBM (schw_integrate_batch_full_dev); rays are chosen so that none
ends early (every row is stored).  Prints steps/s, the written GB/s against MEASURED_PEAKS.json's HBM figure and against
the FP64 bound of the same loop body (the render kernel's step rate), and checks a sample of rows against
integrate_batch (the final state of a ray after k steps = row k of its trajectory).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
import schw_b200  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=454656, help="default: whole waves of both kernels (148 SMs x 6 CTAs x 128 threads x 4 = 148 x 4 x 256 x 3)")
ap.add_argument("--steps", type=int, default=1024)
ap.add_argument("--delta", type=float, default=0.01)
ap.add_argument("--kind", default="camera", choices=["camera", "general"])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="")
a = ap.parse_args()

n, steps = a.rays, a.steps
rng = np.random.default_rng(0)
q0 = np.zeros((n, 4)); p0 = np.zeros((n, 4))
if a.kind == "camera":          # equatorial camera rays of the default scene (utils.py:148-194): the planar fast path applies
    al = np.deg2rad(rng.uniform(12.0, 50.0, n)); r0 = np.full(n, 30.0); th0 = np.full(n, np.pi / 2)
    pr = -np.cos(al) * np.sqrt(1 - 2 / r0); pth = -0.0 * r0; pph = np.sin(al) * r0
else:                           # tilted rays with real polar motion: the general flows, three sincos per step
    r0 = rng.uniform(40, 60, n); th0 = rng.uniform(0.6, np.pi - 0.6, n)
    al = np.deg2rad(rng.uniform(15, 55, n)); be = rng.uniform(-np.pi, np.pi, n)
    pr = -np.cos(al) * np.sqrt(1 - 2 / r0); pth = np.sin(al) * np.sin(be) * r0; pph = np.sin(al) * np.cos(be) * r0 * np.sin(th0)
f = 1 - 2 / r0
Cq = f * pr * pr + pth * pth / r0 ** 2 + pph * pph / (r0 * r0 * np.sin(th0) ** 2)
q0[:, 1] = r0; q0[:, 2] = th0
p0[:, 0] = np.sqrt(4 * Cq / f) / (2 / f); p0[:, 1] = pr; p0[:, 2] = pth; p0[:, 3] = pph

hd = schw_b200.Handle(0)
dq, dp = torch.from_numpy(q0).cuda(), torch.from_numpy(p0).cuda()
traj = torch.empty((n, steps, 4), dtype=torch.float64, device="cuda")
rows = torch.zeros(n, dtype=torch.int64, device="cuda")
st = torch.cuda.Stream()
best = 1e30
for r in range(a.reps):
    hd.integrate_batch_full_dev(dq.data_ptr(), dp.data_ptr(), n, steps, a.delta, 2.0, 1e6, 1.0, traj.data_ptr(), rows.data_ptr(),
                                stream=st.cuda_stream)
    hd.last_status()
    ms, launches = hd.last_kernel_ms()
    best = min(best, ms)
total, warp_iters, planar_iters = hd.last_counters()
assert int(rows.min()) == steps and int(rows.max()) == steps, (int(rows.min()), int(rows.max()))
assert total == n * (steps - 1)
# parity of a sample: row k of ray i == integrate_batch's final state after k steps
idx = rng.choice(n, 64, replace=False)
for k in (1, steps // 2, steps - 1):
    ref = hd.integrate_batch(q0[idx], p0[idx], k, a.delta, 2.0, 1e6, 1.0)
    got = traj[torch.from_numpy(idx).cuda(), k].cpu().numpy()
    assert np.array_equal(ref, got), k
# context: what a pure write stream reaches on this box (torch fill of the same buffer, CUDA events)
fill_best = 1e30
for r in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); traj.fill_(1.0); e1.record(); torch.cuda.synchronize()
    fill_best = min(fill_best, e0.elapsed_time(e1))
fill_gbs = traj.numel() * 8 / (fill_best * 1e-3) / 1e9
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
hbm = peaks.get("hbm_gbs", 6543.1)
rows_written = n * steps
gbs = rows_written * 32 / (best * 1e-3) / 1e9
res = {"what": f"integrate_batch_full_dev, {n} {a.kind} rays x {steps} rows, delta {a.delta}: {rows_written * 32 / 2**30:.2f} GiB of trajectory in HBM",
       "kernel_ms": best, "launches": launches, "rows_per_s": rows_written / (best * 1e-3), "steps_per_s": total / (best * 1e-3),
       "algorithmic_bytes": rows_written * 32, "achieved_gb_s": gbs, "hbm_peak_gb_s": hbm, "frac_of_hbm": gbs / hbm,
       "pure_write_gb_s_torch_fill_same_buffer": fill_gbs, "frac_of_pure_write": gbs / fill_gbs,
       "planar_fast_path_fraction": planar_iters / warp_iters if warp_iters else None,
       "rows_equal_integrate_batch_after_k_steps": True}
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
