#!/bin/bash
# After the last change to csrc/: the three --set full captures and the bench line against the final kernels
# (the artefacts are keyed by the SHA of csrc/; bench.py refuses a stale one).   gpurun -- 'bash tools/gpu_recapture.sh r02'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
python tools/profile_render.py --height 2160 --width 3840 --reps 2 --batch 400000 > gpurun_out/plain_$TAG.log 2>&1 && tail -3 gpurun_out/plain_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:persistent_tiles -c 1 -f -o gpurun_out/prof_render_4k_$TAG \
    python tools/profile_render.py --height 2160 --width 3840 --reps 1 > gpurun_out/ncu_render_$TAG.log 2>&1; echo "ncu render rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:persistent_rays -s 1 -c 1 -f -o gpurun_out/prof_batch_general_$TAG \
    python tools/profile_render.py --height 64 --width 64 --reps 1 --batch 400000 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu batch rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_full_planar_kernel -c 1 -f -o gpurun_out/prof_full_$TAG \
    python tools/bench_full.py --kind camera --reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
