#!/bin/bash
# 1/2/4/8-GPU strong scaling of the bench frame + the two-rank NCCL test + the 360-frame sweep (config 5).
# Run on an N-GPU box: gpurun --gpus 8 -- 'bash tools/gpu_scaling.sh r01'
cd "$(dirname "$0")/.."
TAG=${1:-r01}
NG=$(nvidia-smi -L | wc -l)
# the two-rank NCCL tests (gathered frame == single-GPU frame, with and without de-duplication): the log is kept
[ -n "$SKIP_TESTS" ] || { timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA > gpurun_out/pytest_multirank_${TAG}.log 2>&1; echo "multirank pytest rc=$?" >> gpurun_out/pytest_multirank_${TAG}.log; tail -3 gpurun_out/pytest_multirank_${TAG}.log; }
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/scale_${TAG}_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print($N, d["value"], d["frame_ms"], d["e2e"]["frame_ms"], d.get("gather_check"), (d.get("dedup_opt_in") or {}).get("frame_ms_e2e"), d["clocks"])
PY
done
if [ $NG -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 \
    tools/bench_sweep.py > gpurun_out/sweep_${TAG}_n$NG.json 2> gpurun_out/sweep_${TAG}_n$NG.err
  echo "sweep rc=$?"; grep "^{" gpurun_out/sweep_${TAG}_n$NG.json
fi
