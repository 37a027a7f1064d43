#!/usr/bin/env python
"""The reference's OWN Numba-CUDA kernels of the rows next to the hot path, on this GPU, beside ours
(TEST INFRASTRUCTURE, never imported by the product; companion of oracle/ref_numba_gpu.py).

    make -C oracle ref                                  # in the build container: stages the unmodified reference source
    python oracle/ref_numba_gpu_next_rows.py [--out gpurun_out/reference_numba_gpu_next_rows.json]       # on the GPU box

  a11  `CUDASchwarzschildIntegrator.integrate_batch_full` (reference `cuda_geodesic.py:604-625,647-653`) on 4096 camera rays
       of the default scene x 8192 rows (1 GiB of trajectories): wall time of the call with host arrays as the reference
       makes it, kernel time with device arrays, and the agreement of every stored row with ours.
  N1   `cuda_flat_raytrace` / `cuda_flat_raytrace_with_trajectories` (reference `cuda_geodesic.py:18-292`) on 2 M directions:
       wall time of the call and the agreement of the image and the sampled trajectories with ours.

Writes ONE JSON object: the numbers, or the exact failure.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--flat-rays", type=int, default=1920 * 1080)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_numba_gpu_next_rows.json"))
a = ap.parse_args()

res = {"what": "reference Numba-CUDA kernels of the rows next to the hot path (unmodified simulation/cuda_geodesic.py) vs "
               "libschw_b200 on the same B200, same inputs"}
try:
    os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
    os.environ.pop("NUMBA_ENABLE_CUDASIM", None)
    import numba
    from numba import cuda
    res["numba"] = numba.__version__
    ref_dir = os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(ref_dir, "simulation", "cuda_geodesic.py")):
        raise RuntimeError("oracle/_ref/simulation/cuda_geodesic.py is missing: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, ref_dir)
    import simulation.cuda_geodesic as refmod           # the reference's module, unmodified
    assert os.path.realpath(refmod.__file__).startswith(os.path.realpath(ref_dir)), refmod.__file__
    import schw_b200
    hd = schw_b200.Handle(0)

    # ---- a11: full trajectories ------------------------------------------------------------------
    sc = schw_b200.make_scene(height=400, width=400, has_texture=False)
    frame = hd.render(sc, None, extras=True)
    pick = np.linspace(0, 400 * 400 - 1, a.rays).astype(np.int64)
    n, rows = pick.size, a.rows
    q0s = np.tile(np.array([0.0, 30.0, np.pi / 2, 0.0]), (n, 1))
    q0s[:, 1] = np.hypot(np.hypot(sc.obs[0], sc.obs[1]), sc.obs[2])
    p0s = np.ascontiguousarray(frame["ic"][pick, 0:4])
    integ = refmod.CUDASchwarzschildIntegrator(steps=rows, delta=sc.delta, mass=sc.mass, r_max=sc.r_max)
    small = refmod.CUDASchwarzschildIntegrator(steps=16, delta=sc.delta, mass=sc.mass, r_max=sc.r_max)
    small.integrate_batch_full(q0s[:64], p0s[:64]); cuda.synchronize()                 # JIT
    t0 = time.perf_counter()
    ref_traj = integ.integrate_batch_full(q0s, p0s)                                    # the call the reference makes
    cuda.synchronize()
    ref_wall = time.perf_counter() - t0
    dq, dp_ = cuda.to_device(q0s), cuda.to_device(p0s)
    dt = cuda.to_device(np.zeros((n, rows, 4)))
    e0, e1 = cuda.event(), cuda.event()
    e0.record()
    refmod.fantasy_integrate_batch_ord2_full[(n + 31) // 32, 32](dq, dp_, rows, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0, dt)
    e1.record(); e1.synchronize()
    ref_kernel_ms = cuda.event_elapsed_time(e0, e1)
    del dt
    hd.integrate_batch_full(q0s[:64], p0s[:64], 16, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0)
    t0 = time.perf_counter()
    our_traj, our_rows = hd.integrate_batch_full(q0s, p0s, rows, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0, want_rows=True)
    our_wall = time.perf_counter() - t0
    our_kernel_ms, _ = hd.last_kernel_ms()
    stored = int(our_rows.sum())
    # rows the reference stored: everything up to and including the row that ended the ray (cuda_geodesic.py:619-623)
    ref_rows = np.zeros(n, np.int64)
    for k in range(n):
        r = ref_traj[k, :, 1]
        end = np.flatnonzero((r <= 1.1 * 2.0 * sc.mass) | (r >= sc.r_max))
        ref_rows[k] = end[0] + 1 if end.size else rows
    esc = frame["cls"].reshape(-1)[pick] >= 2
    same_rows = ref_rows == our_rows
    with np.errstate(invalid="ignore"):
        rel = np.abs(ref_traj - our_traj) / np.maximum(np.abs(our_traj), 1.0)
    ok = esc & same_rows
    res["a11_integrate_batch_full"] = {
        "rays": int(n), "rows_allocated": int(rows), "rows_stored": stored, "trajectory_bytes": int(n) * int(rows) * 32,
        "reference_gpu": {"call_wall_s_host_arrays": ref_wall, "kernel_ms": ref_kernel_ms,
                          "stored_steps_per_s_kernel": stored / (ref_kernel_ms * 1e-3)},
        "libschw_b200": {"call_wall_s_host_arrays": our_wall, "kernel_ms": our_kernel_ms,
                         "stored_steps_per_s_kernel": stored / (our_kernel_ms * 1e-3)},
        "speedup_kernel": ref_kernel_ms / our_kernel_ms, "speedup_call": ref_wall / our_wall,
        "agreement": {"escaped_rays": int(esc.sum()), "rays_with_equal_row_counts": int(same_rows.sum()),
                      "escaped_rays_with_equal_row_counts": int(ok.sum()),
                      "max_err_all_rows_of_those_t_r_th_ph": [float(np.nanmax(rel[ok][:, :, c])) for c in range(4)] if ok.any() else None,
                      "rows_after_the_last_stored_one_are_zero_in_both":
                          bool(all((ref_traj[k, ref_rows[k]:] == 0).all() and (our_traj[k, our_rows[k]:] == 0).all() for k in range(0, n, 37)))},
    }
    print(json.dumps(res["a11_integrate_batch_full"]), flush=True)
    del ref_traj, our_traj, rel

    # ---- N1: flat-space ray trace ------------------------------------------------------------------
    try:
        rng = np.random.default_rng(5)
        m = a.flat_rays
        obs = np.array([30.0, 0.0, 0.0])
        d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        bg = rng.integers(0, 256, size=(1024, 2048, 3), dtype=np.uint8)
        idx = np.linspace(0, m - 1, 20).astype(np.int32)
        flat = []
        for (R, pct, pcp, pst, psp, flt, flp) in ((31.0, np.pi / 2, np.pi, np.radians(60), np.radians(90), False, False),
                                                  (31.0, np.pi / 2, 0.0, np.radians(120), np.radians(200), True, True),
                                                  (20.0, 1.2, 4.0, np.radians(45), np.radians(45), False, True)):
            refmod.cuda_flat_raytrace(obs, d[:64], R, pct, pcp, pst, psp, flt, flp, bg); cuda.synchronize()
            t0 = time.perf_counter()
            ref_img = refmod.cuda_flat_raytrace(obs, d, R, pct, pcp, pst, psp, flt, flp, bg)
            cuda.synchronize()
            ref_wall = time.perf_counter() - t0
            hd.flat_raytrace(obs, d[:64], R, pct, pcp, pst, psp, flt, flp, bg)
            t0 = time.perf_counter()
            our_img = hd.flat_raytrace(obs, d, R, pct, pcp, pst, psp, flt, flp, bg)
            our_wall = time.perf_counter() - t0
            ref_img2, ref_tr = refmod.cuda_flat_raytrace_with_trajectories(obs, d, R, pct, pcp, pst, psp, flt, flp, bg, idx, n_points=100)
            our_img2, our_tr = hd.flat_raytrace(obs, d, R, pct, pcp, pst, psp, flt, flp, bg, sampled_indices=idx, n_points=100)
            diff = (ref_img != our_img).any(1)
            flat.append({"boundary_radius": R, "patch": [pct, pcp, pst, psp], "flips": [flt, flp], "rays": int(m),
                         "lit_pixels_reference": int((ref_img != 0).any(1).sum()), "lit_pixels_ours": int((our_img != 0).any(1).sum()),
                         "pixels_that_differ": int(diff.sum()),
                         "pixels_that_differ_with_trajectories_variant": int((ref_img2 != our_img2).any(1).sum()),
                         "max_abs_err_sampled_trajectories": float(np.abs(ref_tr - our_tr).max()),
                         "reference_call_wall_s": ref_wall, "our_call_wall_s": our_wall, "speedup_call": ref_wall / our_wall})
            print(json.dumps(flat[-1]), flush=True)
        res["N1_cuda_flat_raytrace"] = flat
    except Exception as e:                                # e.g. the reference kernel does not compile for a real GPU
        res["N1_cuda_flat_raytrace"] = {"status": "reference failed", "error": repr(e)[:1200]}
        t0 = time.perf_counter()
        our_img = hd.flat_raytrace(obs, d, 31.0, np.pi / 2, np.pi, np.radians(60), np.radians(90), False, False, bg)
        res["N1_cuda_flat_raytrace"]["ours_alone"] = {"rays": int(m), "call_wall_s": time.perf_counter() - t0,
                                                      "lit_pixels": int((our_img != 0).any(1).sum())}
        print(json.dumps(res["N1_cuda_flat_raytrace"])[:1500], flush=True)
    res["status"] = "ok"
except BaseException as e:                                # the outcome is the deliverable, failure included
    import traceback
    res["status"] = "failed"
    res["error"] = repr(e)
    res["traceback"] = traceback.format_exc()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res)[:3000])
