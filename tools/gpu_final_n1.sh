#!/bin/bash
# Round record on one B200: GPU suite, smoke, bench (+ reference arm), ncu launch list + --set full captures (render kernel,
# trajectory kernel, general-ray batch kernel), probes, A/Bs, the reference's own Numba kernel.   gpurun -- 'bash tools/gpu_final_n1.sh r02'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_line_n1_$TAG.json 2> gpurun_out/bench_line_n1_$TAG.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_line_n1_$TAG.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_line_reference_arm_$TAG.json 2>/dev/null; echo "ref arm rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_line_n1_$TAG.json").read().strip().splitlines()[-1])
print("frame", d["frame_ms"], "e2e", d["e2e"]["frame_ms"], "frac", d["roofline"]["frac"], d["roofline"]["sm_cycles_per_warp_step"], [ (k, v["frame_ms"]) for k, v in d.get("other_configs", {}).items()], (d.get("dedup_opt_in") or {}).get("frame_ms"), [e["wall_s"] for e in d.get("entry_point", [])])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_render.py --height 2160 --width 3840 --reps 2 --batch 400000 > gpurun_out/plain_$TAG.log 2>&1 && tail -4 gpurun_out/plain_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:persistent_tiles -c 1 -f -o gpurun_out/prof_render_4k_$TAG \
    python tools/profile_render.py --height 2160 --width 3840 --reps 1 > gpurun_out/ncu_render_$TAG.log 2>&1; echo "ncu render rc=$?"
# general (tilted) rays through integrate_batch: launches = probe, fast-path kernel (exits at once: the probe says "not planar"),
# the per-lane-refill kernel without fast path and trig memo -- the second persistent_rays launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:persistent_rays -s 1 -c 1 -f -o gpurun_out/prof_batch_general_$TAG \
    python tools/profile_render.py --height 64 --width 64 --reps 1 --batch 400000 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu batch rc=$?"
for kind in camera general; do
  timeout 300 python tools/bench_full.py --kind $kind --out gpurun_out/full_${kind}_$TAG.json 2>&1 | tail -1 | cut -c1-330
done
SCHW_FULL_RAYS=1 timeout 300 python tools/bench_full.py --kind camera --out gpurun_out/full_camera_one_ray_per_thread_$TAG.json 2>&1 | tail -1 | cut -c1-330
timeout 300 python tools/bench_full.py --kind camera --rays 4096 --steps 8192 --out gpurun_out/full_camera_4096x8192_$TAG.json 2>&1 | tail -1 | cut -c1-330
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_full_planar_kernel -c 1 -f -o gpurun_out/prof_full_$TAG \
    python tools/bench_full.py --kind camera --reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/step_probe.py --out gpurun_out/step_probe_$TAG.json > /dev/null 2>&1; echo "probe rc=$?"
timeout 600 python tools/ab_tile.py --out gpurun_out/ab_tile_$TAG.json > /dev/null 2>&1; echo "ab_tile rc=$?"
timeout 300 python tools/time_dedup.py --reps 2 > gpurun_out/dedup_4k_$TAG.json 2>/dev/null; echo "dedup rc=$?"; tail -1 gpurun_out/dedup_4k_$TAG.json | cut -c1-300
bash tools/gpu_ref_numba.sh
ls -la gpurun_out/*.ncu-rep | cut -c30-
