#!/bin/bash
# Round check on a 2-GPU box: GPU test suite (incl. the two-rank NCCL tests), smoke, the bench line at N=1 and N=2, X1.
# gpurun --gpus 2 -- 'bash tools/gpu_check2.sh r02'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}_n1.err; cat gpurun_out/bench_${TAG}_n1.json
NG=$(nvidia-smi -L | wc -l)
if [ $NG -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err
  echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}_n2.err; grep "^{" gpurun_out/bench_${TAG}_n2.json
fi
bash tools/gpu_ref_numba.sh
