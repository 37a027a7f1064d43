#!/usr/bin/env python
"""One rank's share of the de-duplicated 4K frame (every world-th distinct ray) on ONE GPU, by residency of the tile kernel.

    python tools/ab_dedup_share.py [--world 8]

Times schw_dedup_integrate_dev (the stage a rank of an N-rank job runs between the plan and the all_gather) with CUDA events
for SCHW_TILE_RAYS x SCHW_TILE_CTAS; the environment switches are read on every call and change no result bit.
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
ap.add_argument("--world", type=int, nargs="+", default=[8, 4, 2])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
hd = schw_b200.Handle(0)
sc = schw_b200.make_scene(height=2160, width=3840, has_texture=False)
st = torch.cuda.Stream()
out = []
for world in a.world:
    os.environ.pop("SCHW_TILE_RAYS", None); os.environ.pop("SCHW_TILE_CTAS", None)
    nu, chunk = hd.dedup_plan_dev(sc, world, stream=st.cuda_stream)
    fq = torch.zeros((world * chunk, 4), dtype=torch.float64, device="cuda")
    ns = torch.zeros(world * chunk, dtype=torch.int32, device="cuda")
    ref = None
    for rays, ctas in ((None, None), (1, 6), (1, 4), (1, 3), (1, 2), (2, 4), (2, 3), (2, 2)):
        for k, v in (("SCHW_TILE_RAYS", rays), ("SCHW_TILE_CTAS", ctas)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st)
                hd.dedup_integrate_dev(sc, 0, fq.data_ptr(), ns.data_ptr(), stream=st.cuda_stream)
                e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1))
        got = (fq[:chunk].cpu().numpy().view(np.int64).copy(), ns[:chunk].cpu().numpy().copy())
        if ref is None:
            ref = got
        same = np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
        out.append({"world": world, "rays_of_rank0": int(-(-nu // world)), "tile_rays": rays, "tile_ctas": ctas, "integrate_ms": best,
                    "max_steps": int(got[1].max()), "identical_to_default": bool(same)})
        print(json.dumps(out[-1]), flush=True)
