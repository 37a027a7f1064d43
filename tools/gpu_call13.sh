#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r02i}
V=$PWD/schwarzchild-einsteinpy-raytracing_b200/lib/variants
echo "== trajectory kernel: 6 / 5 / 4 CTAs per SM"
python tools/bench_full.py --kind camera --steps 2048 2>&1 | tail -1 | cut -c100-330
SCHW_B200_LIB=$V/libschw_b200_fullmb5.so python tools/bench_full.py --kind camera --steps 2048 2>&1 | tail -1 | cut -c100-330
SCHW_B200_LIB=$V/libschw_b200_fullmb4.so python tools/bench_full.py --kind camera --steps 2048 2>&1 | tail -1 | cut -c100-330
echo "== 400x400 with fewer CTAs per SM (tile kernel R=1)"
for c in 6 5 4 3; do echo "ctas $c"; SCHW_TILE_CTAS=$c python tools/profile_render.py --height 400 --width 400 --reps 4 | grep render | tail -1; done
echo "== 1/8 shard with fewer CTAs"
for c in 6 5 4; do echo "ctas $c"; SCHW_TILE_CTAS=$c python tools/profile_render.py --height 2160 --width 3840 --stride 8 --reps 3 | grep render | tail -1; done
echo "== general batch with / without the trig memo"
python tools/profile_render.py --height 64 --width 64 --reps 1 --batch 400000 | grep integrate_batch | tail -1
SCHW_TRIG_MEMO=0 python tools/profile_render.py --height 64 --width 64 --reps 1 --batch 400000 | grep integrate_batch | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:persistent_rays -s 1 -c 1 -f -o gpurun_out/prof_batch_general_$TAG \
    python tools/profile_render.py --height 64 --width 64 --reps 1 --batch 400000 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu batch rc=$?"; tail -3 gpurun_out/ncu_batch_$TAG.log | cut -c1-200
