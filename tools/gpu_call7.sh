#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r02f}
timeout 900 python -m pytest tests/test_gpu_planar.py tests/test_gpu_robustness.py -m gpu -q 2>&1 | tail -5 | cut -c1-300
timeout 600 python tools/ab_tile.py --out gpurun_out/ab_tile_$TAG.json 2>&1 | tail -8 | cut -c1-900
