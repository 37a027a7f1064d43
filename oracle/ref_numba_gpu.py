#!/usr/bin/env python
"""Time the reference's OWN Numba-CUDA kernel on this GPU (TEST INFRASTRUCTURE, never imported by the product).

    make -C oracle ref                      # in the build container: stages the unmodified reference source
    python oracle/ref_numba_gpu.py [--frames 400x400 1920x1080] [--out profiles/r02_reference_numba_gpu.json]      # on the GPU box

Runs `CUDASchwarzschildIntegrator.integrate_batch` of the UNMODIFIED `simulation/cuda_geodesic.py` (reference
`cuda_geodesic.py:579-602,629-645`: 32-thread blocks, state in `cuda.local.array`) on the camera rays of the default
`config.py` scene (BASELINE config 1 = 400x400), exactly as `raytracing.py:118-120` calls it: host NumPy in, host
NumPy out, implicit transfers.  The rays (q0, p0) and their step counts come from this repository's render of the same
frame (`ic`, `nsteps` outputs; step counts are bit-equal to the oracle's, tests/test_gpu_parity.py).  Writes ONE JSON
object: either `reference_gpu: {steps_per_s, ...}` plus the agreement of its final states with ours, or the exact failure.
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--frames", nargs="+", default=["400x400"], help="WxH frames of the default scene")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_numba_gpu.json"))
a = ap.parse_args()

res = {"what": "reference fantasy_integrate_batch_ord2 (unmodified simulation/cuda_geodesic.py, Numba-CUDA) vs libschw_b200 "
               "on the same B200, same rays", "runs": []}
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
    dev = cuda.get_current_device()
    res["device"] = {"name": dev.name.decode() if isinstance(dev.name, bytes) else str(dev.name),
                     "compute_capability": list(dev.compute_capability)}
    from numba.cuda.cudadrv import nvvm
    res["numba_supported_ccs_max"] = list(max(nvvm.get_supported_ccs()))
    import simulation.cuda_geodesic as refmod           # the reference's module, unmodified
    assert os.path.realpath(refmod.__file__).startswith(os.path.realpath(ref_dir)), refmod.__file__
    import schw_b200
    hd = schw_b200.Handle(0)
    for frame in a.frames:
        fw, fh = [int(v) for v in frame.lower().split("x")]
        run = {"height": fh, "width": fw}
        sc = schw_b200.make_scene(height=fh, width=fw, has_texture=False)
        hd.render(sc, None)                              # warm-up (the Numba kernel gets one too)
        hd.render(sc, None)
        ours_ms, _ = hd.last_kernel_ms()
        ours = hd.render(sc, None, extras=True)          # per-ray outputs for the comparison (a little slower: 8x the stores)
        nsteps = ours["nsteps"].astype(np.int64)
        n = fw * fh
        q0s = np.tile(np.array([0.0, 30.0, np.pi / 2, 0.0]), (n, 1))
        q0s[:, 1] = np.hypot(np.hypot(sc.obs[0], sc.obs[1]), sc.obs[2])
        p0s = np.ascontiguousarray(ours["ic"][:, 0:4])
        integ = refmod.CUDASchwarzschildIntegrator(steps=sc.steps, delta=sc.delta, mass=sc.mass, r_max=sc.r_max)   # raytracing.py:118
        t0 = time.perf_counter()
        integ.integrate_batch(q0s[:64], p0s[:64])        # JIT compile + first launch
        cuda.synchronize()
        run["jit_first_call_s"] = time.perf_counter() - t0
        # (1) the call the reference makes: host arrays, implicit H2D/D2H (raytracing.py:120)
        t0 = time.perf_counter()
        out, _ = integ.integrate_batch(q0s, p0s)
        cuda.synchronize()
        wall = time.perf_counter() - t0
        # (2) kernel alone: device arrays, CUDA events
        dq, dp_, do = cuda.to_device(q0s), cuda.to_device(p0s), cuda.device_array((n, 4), np.float64)
        e0, e1 = cuda.event(), cuda.event()
        e0.record()
        refmod.fantasy_integrate_batch_ord2[(n + 31) // 32, 32](dq, dp_, sc.steps, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0, do)
        e1.record(); e1.synchronize()
        kern_ms = cuda.event_elapsed_time(e0, e1)
        steps = int(nsteps.sum())
        fq = ours["final_q"]
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(out[:, 1] - fq[:, 1]) / np.abs(fq[:, 1])
        esc = ours["cls"].reshape(-1) >= 2                # escaped rays (well-posed in the default scene)
        same_side = ((out[:, 1] >= sc.r_max) == (fq[:, 1] >= sc.r_max))
        # who the disagreeing rays are: band G = |b / (3 sqrt(3) M) - 1| < 1e-3 with b = |p_phi / p_t|; band H = the
        # trajectory reaches r < 2.5 M (the horizon blow-up band), measured on our own trajectory of each such ray
        # (integrate_batch_full).  And what they do to the image: classes of the reference kernel's final states (our
        # classify stage, raytracing.py:166-273) next to the frame's.
        ic = ours["ic"]
        b = np.abs(ic[:, 3]) / np.abs(ic[:, 0])
        in_g = np.abs(b / (3.0 * np.sqrt(3.0) * sc.mass) - 1.0) < 1e-3
        bad = ~same_side
        kb = np.flatnonzero(bad)[:4096]
        in_h = np.zeros(n, bool); rmin_bad = np.full(n, np.nan)
        if kb.size:
            cap = int(nsteps[kb].max()) + 1
            traj, rows = hd.integrate_batch_full(q0s[kb], p0s[kb], cap, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0, want_rows=True)
            for m, k in enumerate(kb):
                rmin_bad[k] = np.nanmin(traj[m, :max(int(rows[m]), 1), 1])
            in_h[kb] = rmin_bad[kb] < 2.5 * sc.mass
            del traj
        # the rays neither band explains: how do BOTH kernels react to a 1-ulp change of their initial momentum?
        left = kb[~(in_g[kb] | in_h[kb])]
        ulp_probe = []
        if left.size:
            _, first = np.unique(p0s[left], axis=0, return_index=True)
            for k in left[np.sort(first)][:4]:
                E = np.repeat(p0s[k][None, :], 9, 0)
                for m, (c, d_) in enumerate(((0, 1), (0, -1), (1, 1), (1, -1), (3, 1), (3, -1), (3, 2), (3, -2))):
                    for _ in range(abs(d_)):
                        E[m + 1, c] = np.nextafter(E[m + 1, c], np.inf * d_)
                Qe = np.repeat(q0s[k][None, :], 9, 0)
                mine_e = hd.integrate_batch(Qe, E, sc.steps, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0)
                ref_e, _ = integ.integrate_batch(Qe, E)
                ulp_probe.append({"b_over_bcrit": float(b[k] / (3.0 * np.sqrt(3.0) * sc.mass)),
                                  "pixels_with_this_ray": int((p0s[left] == p0s[k]).all(1).sum()),
                                  "min_r_of_our_stored_rows": float(rmin_bad[k]),
                                  "r_final_ours_over_1ulp_ensemble": [float(v) for v in mine_e[:, 1]],
                                  "r_final_reference_over_1ulp_ensemble": [float(v) for v in ref_e[:, 1]]})
        cls_ref = hd.classify(sc, out, ic[:, 4].copy(), ic[:, 5].copy(), None)["cls"]
        cls_ours = ours["cls"].reshape(-1)
        relq = np.abs(out - fq) / np.maximum(np.abs(fq), 1.0)
        run.update({
            "rays": n, "steps": steps,
            "reference_gpu": {"kernel_ms": kern_ms, "steps_per_s": steps / (kern_ms * 1e-3),
                              "call_wall_s_host_arrays": wall, "steps_per_s_host_arrays": steps / wall,
                              "launch": "blocks of 32 threads, one ray per thread (cuda_geodesic.py:642-644)"},
            "libschw_b200": {"kernel_ms": ours_ms, "steps_per_s": steps / (ours_ms * 1e-3)},
            "speedup_kernel": kern_ms / ours_ms,
            "agreement": {"escaped_rays": int(esc.sum()),
                          "max_rel_err_r_final_escaped": float(np.nanmax(rel[esc])) if esc.any() else None,
                          "max_err_final_state_escaped_t_r_th_ph": [float(np.nanmax(relq[esc, k])) for k in range(4)] if esc.any() else None,
                          "escape_side_mismatches_all_rays": int(bad.sum()),
                          "of_those_in_band_g": int((bad & in_g).sum()),
                          "of_those_examined": int(kb.size),
                          "of_those_in_band_h_trajectory_reaches_r_below_2.5M": int((bad & ~in_g & in_h).sum()),
                          "of_those_outside_both_bands": int(kb.size - (in_g[kb] | in_h[kb]).sum()),
                          "b_over_bcrit_range_of_those": [float((b[kb] / (3.0 * np.sqrt(3.0) * sc.mass)).min()),
                                                          float((b[kb] / (3.0 * np.sqrt(3.0) * sc.mass)).max())] if kb.size else None,
                          "nan_rays_ours_and_reference": [int((~(fq[:, 1] == fq[:, 1])).sum()), int((~(out[:, 1] == out[:, 1])).sum())],
                          "outside_both_bands_ulp_probe": ulp_probe,
                          "outside_both_bands_sample": [
                              {"b_over_bcrit": float(b[k] / (3.0 * np.sqrt(3.0) * sc.mass)), "r_final_ours": float(fq[k, 1]),
                               "r_final_reference": float(out[k, 1]), "nsteps_ours": int(nsteps[k]), "class_ours": int(cls_ours[k]),
                               "class_of_reference_state": int(cls_ref[k])}
                              for k in kb[~(in_g[kb] | in_h[kb])][:8]],
                          "class_mismatches_all_pixels": int((cls_ref != cls_ours).sum()),
                          "class_mismatches_outside_band_g": int(((cls_ref != cls_ours) & ~in_g).sum())},
        })
        res["runs"].append(run)
        print(json.dumps(run), flush=True)
    res["status"] = "ok"
except BaseException as e:                                # the outcome is the deliverable, failure included
    res["status"] = "failed"
    res["error"] = repr(e)
    res["traceback"] = traceback.format_exc()
    print(res["traceback"], file=sys.stderr)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "traceback"}))
