#!/usr/bin/env python
"""BASELINE config 5: 360-frame azimuth camera sweep, frame-parallel over the ranks of a torchrun job.

The reference has no camera sweep; because its observer is hard-wired to the +x axis and the metric is spherically
symmetric, an observer azimuth of k degrees is `--bg-patch-center-phi-relobs k` in the reference's own terms
(SURVEY §8d).  Every frame is a full independent render (the reference re-integrates every run); "fine step" =
delta 0.005.  Frame k goes to rank k mod N; no collective.  Two handles (streams) per rank alternate so the
end-of-frame tail of one launch overlaps the start of the next.

    python tools/bench_sweep.py [--frames 360 --height 1080 --width 1920 --delta 0.005]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_sweep.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
sys.path.insert(0, ROOT)
import schw_b200  # noqa: E402
from bench import synthetic_sky  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=360)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--delta", type=float, default=0.005)
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--reuse", action="store_true",
                help="geodesic-reuse variant (reported separately): de-duplicated render, the geodesics integrated once per "
                     "handle and every further frame of the sweep is a per-pixel lookup (schw_render_dedup_dev)")
a = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h, w = a.height, a.width
d_tex = torch.from_numpy(synthetic_sky(h, w)).to(dev)
handles = [schw_b200.Handle(local) for _ in range(a.streams)]
streams = [torch.cuda.Stream(device=dev) for _ in range(a.streams)]
d_rgb = [torch.empty((h, w, 3), dtype=torch.uint8, device=dev) for _ in range(a.streams)]
d_cls = [torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(a.streams)]
mine = list(range(rank, a.frames, world))
host = torch.empty((len(mine), h, w, 3), dtype=torch.uint8).pin_memory()


def scene(k):
    return schw_b200.make_scene(height=h, width=w, delta=a.delta, patch_center_phi=np.deg2rad((180 + k) % 360))


def run(frames):
    for n, k in enumerate(frames):
        s = n % a.streams
        with torch.cuda.stream(streams[s]):
            render = handles[s].render_dedup_dev if a.reuse else handles[s].render_dev
            render(scene(k), d_tex.data_ptr(), 0, h, 1, d_rgb[s].data_ptr(), d_cls[s].data_ptr(), stream=streams[s].cuda_stream)
            host[n].copy_(d_rgb[s], non_blocking=True)
    torch.cuda.synchronize()


if a.reuse:
    os.environ["SCHW_DEDUP_REUSE"] = "0"        # warm up with full frames (page-locked buffers, kernels), then drop the
    run(mine[:a.streams])                       # cached geodesics so that the timed sweep integrates them itself
    steps_per_frame = handles[0].last_dedup_stats()[1]
    os.environ["SCHW_DEDUP_REUSE"] = "1"
    for hd_ in handles:
        hd_.render_dedup_dev(schw_b200.make_scene(height=8, width=8), d_tex.data_ptr(), 0, 8, 1, d_rgb[0].data_ptr(), d_cls[0].data_ptr())
    torch.cuda.synchronize()
else:
    run(mine[:2])                               # warm-up (also builds the cached tile order)
    steps_per_frame = handles[0].last_total_steps()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run(mine)
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    dt = t.item()
    print(json.dumps({"config": f"{a.frames}-frame azimuth sweep, {w}x{h}, delta {a.delta}, frame-parallel over {world} GPU(s), "
                                f"{a.streams} streams per GPU" + (", geodesic reuse (integrate once per handle, then lookups)" if a.reuse else ""),
                      "n_gpus": world, "seconds": dt, "frames_per_s": a.frames / dt,
                      "steps_per_frame": steps_per_frame, "geodesic_steps_per_s": steps_per_frame * a.frames / dt,
                      "distinct_frames_checksum": int(host[:, ::97, ::89].to(torch.int64).sum())}))
if world > 1:
    dist.destroy_process_group()
