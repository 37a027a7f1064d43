#!/bin/bash
# full GPU suite, bench N=1, ncu launch list + --set full of the render kernel and the trajectory kernel
cd "$(dirname "$0")/.."
TAG=${1:-r02g}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${TAG}_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n1.json").read().strip().splitlines()[-1])
print("frame", d["frame_ms"], "e2e", d["e2e"]["frame_ms"], "frac", d["roofline"]["frac"], d["roofline"]["sm_cycles_per_warp_step"], d.get("other_configs"), d.get("dedup_opt_in",{}).get("frame_ms"), d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2>/dev/null; echo "ref rc=$?"
# ncu: launch list of the bench command (after it exited 0 without ncu), then one --set full capture of the render kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_render.py --height 2160 --width 3840 --reps 2 > gpurun_out/plain_$TAG.log 2>&1 && tail -2 gpurun_out/plain_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:persistent_tiles -c 1 -f -o gpurun_out/prof_render_$TAG \
    python tools/profile_render.py --height 2160 --width 3840 --reps 1 > gpurun_out/ncu_render_$TAG.log 2>&1; echo "ncu render rc=$?"
timeout 300 python tools/bench_full.py --kind camera --steps 2048 --out gpurun_out/full_camera_$TAG.json 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_full_kernel -c 1 -f -o gpurun_out/prof_full_$TAG \
    python tools/bench_full.py --kind camera --steps 2048 --reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
