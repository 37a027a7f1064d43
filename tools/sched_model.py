#!/usr/bin/env python
"""Event model of the render launch: how close is longest-first list scheduling of indivisible tiles to the machine's rate,
and what would migrating unfinished tiles between schedulers at the end of the launch gain?  ("the end of the launch")

    python tools/sched_model.py --dump      # on a B200: per-tile step counts of the 4K / 1080p / 400x400 frames -> profiles/r02_tile_steps.npz
    python tools/sched_model.py             # anywhere: the model, from that file and profiles/r02_step_probe.json

Model.  148 SMs x 4 warp schedulers.  A scheduler hosts up to `ctas` resident warps of the tile kernel, each working on a
job = R consecutive tiles of the queue (R rays per thread; the job lasts as long as its longest ray).  With k resident warps
a scheduler advances every one of them by one loop body per T[k] SM cycles, T measured by `schw_step_probe` on this GPU
(loop body only, per residency; profiles/r02_step_probe.json).  A warp that finishes takes the next job of the queue (tiles
sorted longest first, as build_tile_order does); when the queue is empty it retires.  `--calibrate` scales T by one factor so
that the modelled full 4K frame takes the measured time (the probe is 6-8 % slower than the kernel's steady state).
"Migration" = once the queue is empty, after every time slice the unfinished jobs are re-dealt over the schedulers as evenly
as possible, at no cost: an upper bound of what any hand-off scheme inside the kernel could gain.
"""
import argparse
import heapq
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NPZ = os.path.join(ROOT, "profiles", "r02_tile_steps.npz")
FRAMES = {"4k": (2160, 3840), "1080p": (1080, 1920), "400": (400, 400)}

ap = argparse.ArgumentParser()
ap.add_argument("--dump", action="store_true")
ap.add_argument("--calibrate", type=float, default=292.5, help="measured 4K frame, ms (0: use the probe's rates as they are)")
ap.add_argument("--out", default="")
a = ap.parse_args()

if a.dump:
    sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
    import schw_b200
    hd = schw_b200.Handle(0)
    out = {}
    for name, (h, w) in FRAMES.items():
        ns = hd.render(schw_b200.make_scene(height=h, width=w, has_texture=False), None, extras=True)["nsteps"].reshape(h, w)
        tpr = (w + 31) // 32
        pad = np.zeros((h, tpr * 32), np.int32); pad[:, :w] = ns
        out[name] = pad.reshape(h, tpr, 32).max(2).astype(np.int32)          # a tile lasts as long as its longest ray
    np.savez_compressed(NPZ, **out)
    print("wrote", NPZ, {k: v.shape for k, v in out.items()})
    sys.exit(0)

probe = json.load(open(os.path.join(ROOT, "profiles", "r02_step_probe.json")))
F = probe["clock_hz"]
T = {1: {}, 2: {}}                    # T[R][k] = SM cycles per loop body of one warp with k resident warps per scheduler
for v in probe["variants"]:
    R, k = int(v["rays_per_thread"]), int(round(v["warps_per_scheduler"]))
    if R in T and "sm_cycles_per_warp_ray_step" in v:
        T[R][k] = v["sm_cycles_per_warp_ray_step"] * k * R
tiles = np.load(NPZ)
NSCHED = 148 * 4


def simulate(jobs, Tk, ctas, migrate_slice_s=0.0):
    """jobs: loop bodies per job in queue order.  Returns the makespan in seconds."""
    qi = 0
    rem = [[] for _ in range(NSCHED)]
    for _ in range(ctas):
        for s in range(NSCHED):
            if qi < len(jobs):
                rem[s].append(float(jobs[qi])); qi += 1
    last = [0.0] * NSCHED
    heap = []

    def push(s, now):
        if rem[s]:
            heapq.heappush(heap, (now + min(rem[s]) * Tk[len(rem[s])] / F, s, len(rem[s]), min(rem[s])))

    for s in range(NSCHED):
        push(s, 0.0)
    t = 0.0
    while heap:
        tf, s, k, m = heapq.heappop(heap)
        if len(rem[s]) != k or not rem[s] or abs(min(rem[s]) - m) > 1e-9 * max(1.0, m):
            continue                                           # stale event
        adv = (tf - last[s]) * F / Tk[k]
        rem[s] = [r - adv for r in rem[s]]
        last[s] = tf
        nxt = []
        for r in rem[s]:
            if r <= 1e-6:
                if qi < len(jobs):
                    nxt.append(float(jobs[qi])); qi += 1
            else:
                nxt.append(r)
        rem[s] = nxt
        t = max(t, tf)
        if qi >= len(jobs) and migrate_slice_s > 0.0:
            break
        push(s, tf)
    if migrate_slice_s > 0.0 and any(rem):
        # bring every scheduler to the common time t, then advance in slices with a free, even re-deal after each
        for s in range(NSCHED):
            if rem[s]:
                adv = (t - last[s]) * F / Tk[len(rem[s])]
                rem[s] = [r - adv for r in rem[s] if r - adv > 1e-6]
        while True:
            alive = sorted((r for rs in rem for r in rs), reverse=True)
            if not alive:
                break
            rem = [[] for _ in range(NSCHED)]
            for i, r in enumerate(alive):                      # even deal, longest first
                rem[i % NSCHED].append(r)
            dt = migrate_slice_s
            # do not overshoot the end: the slice ends when the first scheduler would run dry
            dt = min(dt, min(min(rs) * Tk[len(rs)] / F for rs in rem if rs))
            for s in range(NSCHED):
                if rem[s]:
                    adv = dt * F / Tk[len(rem[s])]
                    rem[s] = [r - adv for r in rem[s] if r - adv > 1e-6]
            t += dt
    return t


def jobs_for(frame, stride, R):
    d = np.sort(tiles[frame][0::stride].reshape(-1))[::-1].astype(np.float64)
    if R == 2:
        if len(d) % 2:
            d = np.append(d, 0.0)
        d = d.reshape(-1, 2).max(1)                            # two consecutive tiles of the queue share a warp
    return d


scale = 1.0
if a.calibrate > 0:
    t4k = simulate(jobs_for("4k", 1, 2), T[2], 4)
    scale = a.calibrate * 1e-3 / t4k
    for R in T:
        for k in T[R]:
            T[R][k] *= scale
rows = []
for frame, stride, label in (("4k", 1, "3840x2160"), ("4k", 8, "3840x2160, rows 0::8 (one of 8 ranks)"), ("1080p", 1, "1920x1080"),
                             ("400", 1, "400x400")):
    work = tiles[frame][0::stride].astype(np.float64).sum()
    for R, ctas in ((1, 6), (2, 4)):
        full_rate = work * (T[R][ctas] / ctas / R) / F / NSCHED          # every scheduler full from start to end
        j = jobs_for(frame, stride, R)
        lpt = simulate(j, T[R], ctas)
        mig = simulate(j, T[R], ctas, migrate_slice_s=20e-6)
        rows.append({"frame": label, "tiles": int(tiles[frame][0::stride].size), "rays_per_thread": R, "resident_ctas": ctas,
                     "ms_at_full_rate": full_rate * 1e3, "ms_model": lpt * 1e3, "ms_model_with_free_migration": mig * 1e3,
                     "end_of_launch_cost_ms": (lpt - full_rate) * 1e3, "migration_would_gain_ms": (lpt - mig) * 1e3})
        print(json.dumps(rows[-1]), flush=True)
res = {"what": "event model of the tile kernel's launch (tools/sched_model.py): longest-first list scheduling of indivisible tiles "
               "vs the same with free migration of unfinished tiles after the queue is empty",
       "rate_scale_applied_to_probe": scale, "rows": rows}
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
