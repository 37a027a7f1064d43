#!/usr/bin/env python
"""profiles/<tag>_scaling.md + the per-N bench lines from the files tools/gpu_scaling.sh leaves in gpurun_out/ (runs here, no GPU).

    python tools/make_scaling_md.py r02
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = []
for n in (1, 2, 4, 8):
    src = os.path.join(ROOT, "gpurun_out", f"scale_{tag}_n{n}.json")
    if not os.path.exists(src):
        continue
    lines = [l for l in open(src) if l.startswith("{")]
    if not lines:
        continue
    d = json.loads(lines[-1])
    shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_bench_line_n{n}_scaling_box.json"))
    rows.append((n, d))
one = rows[0][1]
out = [f"# Strong scaling of one 3840x2160 frame (bench.py, round-2 kernels, one 8xB200 box, back to back)", "",
       "| GPUs | steps/s (device-timed) | frame ms | speed-up | efficiency | e2e frame ms (host buffers in/out) | e2e speed-up | gather_check (rows, identical) | "
       "de-duplicated frame, ms (opt-in) | SM MHz | rank-0 power W |", "|---|---|---|---|---|---|---|---|---|---|---|"]
for n, d in rows:
    gc = d.get("gather_check")
    dd = d.get("dedup_opt_in") or {}
    ddms = dd.get("frame_ms_e2e", dd.get("frame_ms"))
    ddtxt = "—" if ddms is None else f"{ddms:.2f}{' e2e' if 'frame_ms_e2e' in dd else ' kernel'}{'' if dd.get('identical_to_one_ray_per_pixel') else ' (NOT identical)'}"
    out.append(f"| {n} | {d['value']:.3e} | {d['frame_ms']:.2f} | {one['frame_ms'] / d['frame_ms']:.2f}x | {one['frame_ms'] / d['frame_ms'] / n:.3f} | "
               f"{d['e2e']['frame_ms']:.2f} | {one['e2e']['frame_ms'] / d['e2e']['frame_ms']:.2f}x | "
               f"{'—' if not gc else str(gc['rows']) + ', ' + str(gc['identical']).lower()} | {ddtxt} | "
               f"{(d.get('clocks') or {}).get('sm_mhz')} | {(d.get('clocks') or {}).get('power_w_max')} |")
out += ["", "Launch: `bash tools/gpu_scaling.sh " + tag + "` = `python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 "
        "bench.py --gpus N --steps 3 --warmup 3 --no-cpu-baseline` (the JSON lines are `" + tag + "_bench_line_n{1,2,4,8}_scaling_box.json`).  "
        "Row-cyclic shards; every rank's kernel writes rgb and class side by side into its (rows, 4w) block (`schw_render_dev_pitched`); `value`: one "
        "all_gather of those blocks per frame inside the timed region; `e2e`: every rank uploads the texture (overlapped with the start of "
        "its kernel), one NCCL gather to rank 0, one D2H of the frame into page-locked arrays there.  `gather_check`: rank 0 re-rendered "
        "that many rows alone and compared rgb + class bit for bit with both gathered frames.  De-duplicated frame (opt-in): the ranks "
        "share the integration of the distinct camera rays (one extra all_gather of final states + step counts), per-rank epilogue, frame "
        "gather, D2H; wall time, max over ranks, incl. the texture upload and the plan (keys + sort) on every rank."]
sw = os.path.join(ROOT, "gpurun_out", f"sweep_{tag}_n8.json")
if os.path.exists(sw):
    lines = [l for l in open(sw) if l.startswith("{")]
    if lines:
        shutil.copy(sw, os.path.join(ROOT, "profiles", f"{tag}_sweep_n8.json"))
        s = json.loads(lines[-1])
        out += ["", f"Config 5 (`tools/bench_sweep.py` under torchrun on 8 GPUs, `{tag}_sweep_n8.json`): " + json.dumps({k: s[k] for k in s if k in ("frames", "height", "width", "delta", "seconds", "frames_per_s", "steps_per_s", "wall_s")})]
bd = os.path.join(ROOT, "profiles", f"{tag}_n8_breakdown.json")
if os.path.exists(bd):
    b = json.loads([l for l in open(bd) if l.startswith("{")][-1])
    out += ["", f"Where an 8-rank frame goes (`{tag}_n8_breakdown.json`: `bench.py --gpus 8 --steps 5`, frame {b['frame_ms']:.2f} ms): own kernel per rank "
            f"{b['per_rank']['kernel_ms']} ms, then all_gather incl. waiting for the slowest rank {b['per_rank']['all_gather_incl_wait_ms']} ms."]
mr = os.path.join(ROOT, "gpurun_out", f"pytest_multirank_{tag}.log")
if os.path.exists(mr):
    shutil.copy(mr, os.path.join(ROOT, "profiles", f"{tag}_pytest_multirank.log"))
open(os.path.join(ROOT, "profiles", f"{tag}_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
