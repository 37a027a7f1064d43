#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r02d}
timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py -m gpu -q -k "full or robust or rows or ragged or k3 or kats" 2>&1 | tail -4 | cut -c1-300
timeout 600 python tools/ab_tile.py --out gpurun_out/ab_tile_$TAG.json 2>&1 | tail -8
timeout 300 python tools/step_probe.py --out gpurun_out/step_probe_$TAG.json 2>&1 | tail -12 | cut -c1-260
for kind in camera general; do
  timeout 300 python tools/bench_full.py --kind $kind --rays 262144 --steps 2048 --out gpurun_out/full_${kind}_$TAG.json 2>&1 | tail -1
done
