#!/bin/bash
# A/B of the planar fast path on caller-supplied tilted rays (integrate_batch) and on the 4K frame; run on the GPU box.
cd "$(dirname "$0")/.."
for pl in 1 0; do
echo "== SCHW_PLANAR=$pl"
SCHW_PLANAR=$pl python tools/profile_render.py --height 2160 --width 3840 --reps 2 --batch 400000 | grep -v probe
done
python -m pytest tests/test_gpu_planar.py -q -x 2>&1 | tail -3
