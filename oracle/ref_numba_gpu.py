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
        run.update({
            "rays": n, "steps": steps,
            "reference_gpu": {"kernel_ms": kern_ms, "steps_per_s": steps / (kern_ms * 1e-3),
                              "call_wall_s_host_arrays": wall, "steps_per_s_host_arrays": steps / wall,
                              "launch": "blocks of 32 threads, one ray per thread (cuda_geodesic.py:642-644)"},
            "libschw_b200": {"kernel_ms": ours_ms, "steps_per_s": steps / (ours_ms * 1e-3)},
            "speedup_kernel": kern_ms / ours_ms,
            "agreement": {"escaped_rays": int(esc.sum()),
                          "max_rel_err_r_final_escaped": float(np.nanmax(rel[esc])) if esc.any() else None,
                          "escape_side_mismatches_all_rays": int((~same_side).sum())},
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
