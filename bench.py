#!/usr/bin/env python
"""bench.py — headline measurement of the per-ray geodesic hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the configuration its metric is quoted on): one 3840x2160 frame of the
default config.py scene (observer (30,0,0), fov 80 deg, r_s = 2, r_max = 31, delta = 0.01, step cap 200 000,
omega_eff = 1.0, full-sky patch) over a synthetic sky texture.  A "step" is one frame.  With N ranks the
pixel rows are sharded row-cyclically (fixed total work => "strong" scaling) and each step ends with ONE NCCL
collective of rgb+class (all_gather in the device-timed leg, gather to rank 0 in the end-to-end leg).

One JSON line on rank 0.  `value` = geodesic steps/s with inputs resident in HBM (CUDA events on the launch
stream, max over ranks); `e2e` = the same metric through the host-buffer C-ABI call (pinned host texture in,
rgb/class out, copies inside the timed region); `roofline` against the FP64 FMA peak measured in this run;
`cpu_baseline` = the CPU oracle (oracle/, a port of the reference's arithmetic) on a bounded row sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

def csrc_sha():
    """SHA-256 over the kernel sources: the key of every build-time / ncu artefact bench.py quotes."""
    import hashlib
    hsh = hashlib.sha256()
    for f in ("schw_kernels.cu", "schw_device.cuh", "Makefile"):
        hsh.update(open(os.path.join(PKG, "csrc", f), "rb").read())
    return hsh.hexdigest()[:16]


def keyed_json(path):
    """A committed artefact, or ({}, reason) when it was made from other kernel sources than the ones in the tree."""
    try:
        d = json.load(open(path))
    except (OSError, ValueError) as e:
        return {}, f"{os.path.relpath(path, ROOT)} unreadable ({e.__class__.__name__})"
    if d.get("csrc_sha") != csrc_sha():
        return {}, f"{os.path.relpath(path, ROOT)} is stale: made for csrc {d.get('csrc_sha')}, tree is {csrc_sha()}"
    return d, None


ALGO_FLOP_PER_STEP = 312        # 268 add/mul (FMA = 2) + 36 div + 8 sin/cos
NOMINAL_FP64_TFLOPS = 37.2      # 148 SM x 64 lanes x 2 x 1.965 GHz


def synthetic_sky(h, w, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    t = np.empty((h, w, 3), np.uint8)
    t[..., 0] = (xx * 255 // max(w - 1, 1)).astype(np.uint8)
    t[..., 1] = (yy * 255 // max(h - 1, 1)).astype(np.uint8)
    t[..., 2] = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    return t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def sample_rows(h, nrows):
    """Evenly spaced rows of the frame (a representative mix of captured / near-critical / corner rays)."""
    nrows = max(1, min(h, nrows))
    return np.unique(np.floor((np.arange(nrows) + 0.5) * h / nrows).astype(int))


def cpu_reference_leg(args, nthreads, target_seconds):
    """The reference's CPU arithmetic (oracle port, OpenMP over all host cores) on a bounded row sample."""
    from oracle import oracle as O
    O.build()
    nthreads = nthreads or (os.cpu_count() or 1)
    sc = O.make_scene(height=args.height, width=args.width)
    tex = synthetic_sky(args.height, args.width)
    # ~1.3 us per step per core for the literal port, ~5.9 k steps per ray
    rays = target_seconds * nthreads / (1.3e-6 * 5900)
    rows = sample_rows(args.height, int(round(rays / args.width)))
    def one():
        t0 = time.perf_counter()
        steps = 0
        for r in rows:
            out = O.render(sc, tex, int(r), int(r) + 1, 1, nthreads=nthreads, extras=False)
            steps += int(out["nsteps"].sum())
        return steps, time.perf_counter() - t0
    return one, rows, nthreads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    one, rows, nthreads = cpu_reference_leg(args, 0, target_seconds=8.0)
    for _ in range(args.warmup):
        one()
    tot_steps, tot_t = 0, 0.0
    for _ in range(args.steps):
        s, t = one()
        tot_steps += s; tot_t += t
    value = tot_steps / tot_t
    sample = f"{len(rows)} evenly spaced rows of the {args.width}x{args.height} frame per step ({len(rows) * args.width} rays)"
    line = {
        "impl": "reference", "metric": "geodesic_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "mrays_per_s": 1e-6 * len(rows) * args.width * args.steps / tot_t,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": f"{args.width}x{args.height} Schwarzschild lensing frame, default config.py scene "
                        f"(obs (30,0,0), fov 80deg, r_max 31, delta 0.01, steps cap 200000, omega_eff 1.0, full-sky patch), "
                        f"synthetic sky texture", "height": args.height, "width": args.width,
            "sharding": (f"row-cyclic over {n} rank(s); value: one all_gather of rgb+class per frame; e2e: every rank uploads "
                         f"the texture, one gather of rgb+class to rank 0, one D2H there") if n > 1 else "single GPU",
            "l2": "L2 flushed (256 MiB memset) between timed frames"}


def time_frame(hd, h, w, reps=3):
    """Kernel time (CUDA events inside the library, on its stream) and steps of one default-scene frame, texture resident."""
    import schw_b200
    sc = schw_b200.make_scene(height=h, width=w)
    tex = synthetic_sky(h, w)
    best = 1e30
    for _ in range(reps):
        hd.render(sc, tex)
        best = min(best, hd.last_kernel_ms()[0])
    steps = hd.last_total_steps()
    return {"height": h, "width": w, "frame_ms": best, "steps": steps, "steps_per_s": steps / (best * 1e-3),
            "mrays_per_s": 1e-6 * h * w / (best * 1e-3)}


def time_entry_point(h, w, tmpdir):
    """Wall time of the reference's own entry point, run_manual_simulation (simulation/raytracing.py:16), as main.py calls
    it (n_samples = 20, per-ray diagnostics gathered, sampled trajectories), without the PNG/CSV side effects."""
    import random
    from PIL import Image
    from simulation.blackhole import BlackHole, Observer
    from simulation.raytracing import run_manual_simulation
    path = os.path.join(tmpdir, f"sky_{w}x{h}.png")
    Image.fromarray(synthetic_sky(h // 2 or 1, w // 2 or 1)).save(path)
    obs = Observer(position=np.array([30.0, 0.0, 0.0]), fov=np.radians(80.0), image_size=(h, w))
    kw = dict(steps=200000, delta=0.01, background_path=path, use_cuda=True, boundary_radius=31.0,
              patch_center_theta=np.pi / 2, patch_center_phi=np.pi, patch_size_theta=np.pi, patch_size_phi=2 * np.pi,
              n_samples=20, write_outputs=False)
    import contextlib
    import io
    best = 1e30
    for _ in range(2):
        random.seed(0)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            img, _ = run_manual_simulation(BlackHole(mass=1.0), obs, **kw)
        best = min(best, time.perf_counter() - t0)
    return {"height": h, "width": w, "wall_s": best, "nonblack_pixels": int(np.any(img != 0, axis=2).sum()),
            "what": "run_manual_simulation(..., n_samples=20, write_outputs=False): texture load + LANCZOS resize, fused render "
                    "with all per-ray diagnostics copied to the host, 20 decimated trajectories; best of 2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / entry_point / dedup_opt_in (quick runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import schw_b200
    from schw_b200.multigpu import render_frame, shard_rows, rows_of

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, w = args.height, args.width
    hd = schw_b200.Handle(local_rank)
    info = hd.device_info()
    scene = schw_b200.make_scene(height=h, width=w)
    tex_host = torch.from_numpy(synthetic_sky(h, w)).pin_memory()
    rb, re, rs = shard_rows(h, rank, world)
    max_rows = (h + world - 1) // world
    d_tex = tex_host.to(dev, non_blocking=False)
    # this rank's block (rows, 4w): rgb and class side by side, written by the kernel itself (schw_render_dev_pitched) --
    # the message of the single all_gather
    d_out = torch.zeros((max_rows, 4 * w), dtype=torch.uint8, device=dev)
    d_all = torch.empty((world * max_rows, 4 * w), dtype=torch.uint8, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # A dedicated (non-default) torch stream: the C ABI treats a NULL stream as "the handle's own stream", and
    # torch.cuda.Event only sees work on torch's current stream, so kernels and events must share this one.
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device(mid=None):
        hd.render_dev_pitched(scene, d_tex.data_ptr(), rb, re, rs, d_out.data_ptr(), 4 * w, d_out.data_ptr() + 3 * w, 4 * w,
                              stream=stream)
        if mid is not None:
            mid.record()
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 FMA rate of this device, measured before the run (register-resident DFMA chains)
    probe_tflops, _ = hd.fp64_peak_probe()

    for _ in range(args.warmup):
        step_device()
    barrier()
    hd.last_status()
    steps_per_frame_local, warp_iters, planar_iters = hd.last_counters()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_mid = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]      # between this rank's kernel and the collective
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                   # L2 flush, outside the per-step events
        ev[k][0].record()
        step_device(ev_mid[k])
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    own_kernel_ms = sum(a.elapsed_time(m) for (a, _), m in zip(ev, ev_mid)) / args.steps
    wait_ms = sum(m.elapsed_time(b) for (_, b), m in zip(ev, ev_mid)) / args.steps
    kern_ms, _ = hd.last_kernel_ms()
    hd.last_status()

    # ---- e2e: the host-buffer C-ABI call a user makes (texture H2D + outputs D2H inside the timed region) ----
    def step_e2e():
        # the call a user of the package makes: host texture in, host frame out (sharded + gathered when world > 1)
        # (world > 1: root=0 = the single NCCL gather of rgb + class to rank 0, which copies the frame to the host once)
        return render_frame(scene, tex_host.numpy(), handle=hd, root=0)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    frame0 = None
    for _ in range(args.steps):
        frame0 = None                   # (drop the previous frame first: its page-locked arrays go back to the pool)
        frame0 = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    # ---- multi-GPU correctness in the record: rank 0 re-renders sample rows on its own GPU and compares them bit for bit
    #      with the gathered frame of the device leg (all_gather) and of the e2e leg (gather to rank 0) ----
    gather_check = None
    if world > 1:
        step_device()
        torch.cuda.synchronize()
        if rank == 0:
            nat = d_all.view(world, max_rows, 4 * w).permute(1, 0, 2).reshape(world * max_rows, 4 * w)[:h].cpu().numpy()
            # >= 16 rows, every rank's residue at least twice, top / centre (shadow) / bottom of the frame
            rows = sorted({int(v) for v in np.linspace(0, h - 1, max(16, 2 * world)).round()} |
                          {h // 2 + r for r in range(world) if h // 2 + r < h} | {r for r in range(world) if r < h})
            ident = True
            for r in rows:
                one = hd.render(scene, tex_host.numpy(), r, r + 1, 1)
                ident &= bool(np.array_equal(one["rgb"].reshape(-1), nat[r, :3 * w]) and np.array_equal(one["cls"].reshape(-1), nat[r, 3 * w:]))
                ident &= bool(np.array_equal(one["rgb"][0], frame0["rgb"][r]) and np.array_equal(one["cls"][0], frame0["cls"][r]))
            gather_check = {"rows": len(rows), "residues_covered": len({r % world for r in rows}), "identical": bool(ident),
                            "what": "rank 0 re-rendered these rows alone (schw_render, one row per call) and compared rgb + class bit "
                                    "for bit with the all_gathered frame of the device leg and the root-gathered frame of the e2e leg"}
        barrier()

    # ---- opt-in geodesic de-duplication, reported beside the headline, never in it ----
    dedup = None
    if not args.no_extras:
        try:
            tex_np = tex_host.numpy()
            os.environ["SCHW_DEDUP_REUSE"] = "0"            # every timed frame plans and integrates again
            if world == 1:
                ref = hd.render(scene, tex_np)
                best = 1e30
                for _ in range(max(2, args.steps)):
                    out = hd.render(scene, tex_np, dedup=True)
                    best = min(best, hd.last_kernel_ms()[0])
                uniq, pixel_steps = hd.last_dedup_stats()
                dedup = {"frame_ms": best, "distinct_rays": uniq, "pixels": h * w, "steps_integrated": hd.last_total_steps(),
                         "pixel_steps": pixel_steps, "pixel_steps_per_s": pixel_steps / (best * 1e-3),
                         "identical_to_one_ray_per_pixel": bool(np.array_equal(out["rgb"], ref["rgb"]) and np.array_equal(out["cls"], ref["cls"])),
                         "note": "schw_render_dedup: pixels whose camera angle agrees bitwise share one integration; same output "
                                 "bits; kernel time of the whole pipeline (keys, sort, integrate, per-pixel epilogue), texture resident"}
            else:
                render_frame(scene, tex_np, handle=hd, root=0, dedup=True)
                best = 1e30
                for _ in range(max(2, args.steps)):
                    barrier()
                    t0 = time.perf_counter()
                    out = render_frame(scene, tex_np, handle=hd, root=0, dedup=True)
                    barrier()
                    best = min(best, time.perf_counter() - t0)
                t = torch.tensor([best], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                st = torch.tensor([float(hd.last_total_steps())], dtype=torch.float64, device=dev)
                dist.all_reduce(st, op=dist.ReduceOp.SUM)
                if rank == 0:
                    dedup = {"frame_ms_e2e": 1e3 * t.item(), "pixels": h * w, "steps_integrated": st.item(),
                             "identical_to_one_ray_per_pixel": bool(np.array_equal(out["rgb"], frame0["rgb"]) and np.array_equal(out["cls"], frame0["cls"])),
                             "note": "render_frame(dedup=True, root=0): every rank plans the frame, the ranks share the integration of "
                                     "the distinct rays (one all_gather of final states + step counts), per-rank epilogue, frame gather "
                                     "to rank 0, D2H; wall time incl. the texture upload, max over ranks, best of the timed frames"}
        except Exception as e:                         # the headline must not depend on the opt-in mode
            dedup = {"error": repr(e)}
        os.environ.pop("SCHW_DEDUP_REUSE", None)

    stats = torch.tensor([dev_ms, e2e_s, float(steps_per_frame_local), kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s = mx[0].item(), mx[1].item()
        steps_per_frame = sm[2].item()
    else:
        steps_per_frame = float(steps_per_frame_local)

    per_rank = None
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, (round(own_kernel_ms, 3), round(wait_ms, 3)))
    if rank == 0:
        value = steps_per_frame * args.steps / (dev_ms * 1e-3)
        e2e_value = steps_per_frame * args.steps / e2e_s
        # ---- roofline of the dominant kernel (persistent_rays<RenderJob>), per launch, this rank's shard ----
        # FP64-pipe utilisation, computed in this run: loop bodies executed (device counter) x FP64-pipe instructions per
        # loop body (counted in the SASS of THIS build by tools/loop_body_fp64.py) x 2 issue cycles per warp instruction
        # (16 FP64 lanes per SM sub-partition) / (kernel time x SM clock x 4 sub-partitions x SMs).
        sm_mhz = (clocks or {}).get("sm_mhz") or info["clock_khz"] * 1e-3
        body, body_stale = keyed_json(os.path.join(PKG, "lib", "loop_body.json"))
        ncu, ncu_stale = keyed_json(os.path.join(ROOT, "profiles", "ncu_summary.json"))
        if (h, w) != (2160, 3840) or world != 1:
            ncu, ncu_stale = {}, "the ncu capture is of the 1-GPU 3840x2160 launch"
        peak = info["sm_count"] * 64 * 2 * sm_mhz * 1e6 / 1e12
        roof = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak,
                "peak_source": f"FP64 pipe: {info['sm_count']} SMs x 64 lanes x 2 (FMA) x {sm_mhz:.0f} MHz (SM clock sampled under load); "
                               "MEASURED_PEAKS.json has no FP64 entry; a pure register-resident DFMA stream measured in this run "
                               "(schw_fp64_peak_probe) is dfma_probe_tflops",
                "dfma_probe_tflops": probe_tflops, "kernel": body.get("kernel", "render kernel") if body else "render kernel", "kernel_ms": kern_ms,
                "traffic": ncu.get("dram_bytes_per_launch"), "traffic_note": ncu_stale or ncu.get("source"),
                "algorithmic_bytes_per_launch": 7 * h * w,
                "lane_utilisation": steps_per_frame_local / (32.0 * warp_iters) if warp_iters else None,
                "planar_fast_path_fraction": planar_iters / warp_iters if warp_iters else None,
                "fp64_pipe_pct_of_peak_ncu": ncu.get("fp64_pipe_pct"),
                "executed_flop_per_step_ncu": ncu.get("executed_flop_per_step")}
        algo = ALGO_FLOP_PER_STEP * steps_per_frame_local / (kern_ms * 1e-3) / 1e12
        roof["algorithmic_tflops"] = algo
        roof["algorithmic_flop_per_step"] = ALGO_FLOP_PER_STEP
        roof["algorithmic_speedup_vs_survey_count"] = algo / probe_tflops
        if body and warp_iters:
            fp64_instr = planar_iters * body["fp64_per_body_fast_path"] + (warp_iters - planar_iters) * body["fp64_per_body_general"]
            achieved = fp64_instr * 64 / (kern_ms * 1e-3) / 1e12        # one warp instruction = 32 lanes x 2 (FMA-equivalent)
            roof.update({"achieved": achieved, "frac": achieved / peak, "frac_of_dfma_probe": achieved / probe_tflops,
                         "fp64_instr_per_loop_body": {"fast_path": body["fp64_per_body_fast_path"], "general": body["fp64_per_body_general"],
                                                      "source": f"SASS of this build (lib/loop_body.json, csrc {body['csrc_sha']})"},
                         "sm_cycles_per_warp_step": kern_ms * 1e-3 * sm_mhz * 1e6 * 4 * info["sm_count"] / warp_iters,
                         "note": "achieved = FP64-pipe warp instructions issued by the step loop x 64 flop-equivalents / kernel time, so "
                                 "frac is the utilisation of the FP64 pipe (what ncu reports as sm__pipe_fp64_cycles_active); the "
                                 "reference's algorithmic work (312 flop/step) is algorithmic_tflops -- the kernel executes far fewer "
                                 "flops per step, so that figure can exceed the peak (algorithmic_speedup_vs_survey_count)"})
        else:
            roof.update({"achieved": None, "frac": None, "note": body_stale or "no loop bodies counted"})
        line = {
            "metric": "geodesic_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "frame_ms": dev_ms / args.steps, "mrays_per_s": 1e-6 * h * w * args.steps / (dev_ms * 1e-3),
            "steps_per_frame": steps_per_frame, "wall_s_timed_region": t_wall,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(h * w * 3) * world,
                    "d2h_bytes_per_step": int(4 * w * h), "frame_ms": 1e3 * e2e_s / args.steps},
            "gpu_launches": args.steps * world,
            "roofline": roof,
            "clocks": clocks,
        }
        if per_rank is not None:
            # where a multi-rank frame goes: every rank's own kernel, then the collective (which includes waiting for the
            # slowest rank)
            line["per_rank"] = {"kernel_ms": [p[0] for p in per_rank], "all_gather_incl_wait_ms": [p[1] for p in per_rank]}
        if gather_check is not None:
            line["gather_check"] = gather_check
        if dedup is not None:
            line["dedup_opt_in"] = dedup
        if not args.no_extras and world == 1:
            # BASELINE configs 1 and 3 (kernel time, texture resident) and the reference's entry point end to end
            try:
                import tempfile
                with tempfile.TemporaryDirectory() as td:
                    line["other_configs"] = {"config1_400x400": time_frame(hd, 400, 400), "config3_1920x1080": time_frame(hd, 1080, 1920)}
                    line["entry_point"] = [time_entry_point(400, 400, td), time_entry_point(h, w, td)]
            except Exception as e:
                line["other_configs"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            one, rows, nthreads = cpu_reference_leg(args, 0, target_seconds=15.0)
            s, t = one()
            line["cpu_baseline"] = {"value": s / t, "unit": "steps/s", "cores": nthreads, "kind": "port",
                                    "sample": f"{len(rows)} evenly spaced rows of the same frame ({len(rows) * w} rays, {s} steps, {t:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if gather_check is not None and not gather_check["identical"]:
        raise SystemExit("gather_check failed: the gathered frame differs from rank 0's own render of the sample rows")


if __name__ == "__main__":
    main()
