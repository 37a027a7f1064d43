#!/bin/bash
# ncu captures of the shipped render kernel on the bench workload (run on the GPU box); tag = $1
cd "$(dirname "$0")/.."
TAG=${1:-r01_planar_4k}
set -x
python tools/profile_render.py --height 2160 --width 3840 --reps 2 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
cat gpurun_out/plain_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:persistent_rays -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/profile_render.py --height 2160 --width 3840 --reps 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
