set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_planar.py --reps 3 > gpurun_out/ab_planar_4k.log 2>&1; cat gpurun_out/ab_planar_4k.log
timeout 300 python tools/ab_planar.py --height 1080 --width 1920 --delta 0.005 --reps 2 > gpurun_out/ab_planar_1080_fine.log 2>&1; cat gpurun_out/ab_planar_1080_fine.log
