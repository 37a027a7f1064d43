#!/bin/bash
# Round-end single-GPU evidence (run on the GPU box): tests, smoke, bench lines, ncu launch list + full capture.  tag = $1
cd "$(dirname "$0")/.."
TAG=${1:-r01_final}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference_arm.json 2>> gpurun_out/bench.err; echo "ref rc=$?"; cat gpurun_out/bench_${TAG}_reference_arm.json
timeout 300 python tools/ab_planar.py --reps 2 > gpurun_out/ab_planar_${TAG}.log 2>&1; cat gpurun_out/ab_planar_${TAG}.log
# launch list of the bench command (per-launch device time; cold-cache and serialised under ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1; echo "ncu launch list rc=$?"
bash tools/gpu_ncu_render.sh ${TAG}_4k
