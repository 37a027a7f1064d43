#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r02h}
timeout 900 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -k "full or rows or ragged or k3 or kats or integrator_api or manual_simulation" 2>&1 | tail -4 | cut -c1-300
for kind in camera general; do
  timeout 300 python tools/bench_full.py --kind $kind --steps 2048 --out gpurun_out/full_${kind}_$TAG.json 2>&1 | tail -1
done
timeout 300 python tools/bench_full.py --kind camera --rays 4096 --steps 8192 --out gpurun_out/full_camera_4096x8192_$TAG.json 2>&1 | tail -1
