#!/bin/bash
# pytest (rest of the suite), step probe, trajectory kernel at scale, bench n1
cd "$(dirname "$0")/.."
TAG=${1:-r02c}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python tools/step_probe.py --out gpurun_out/step_probe_$TAG.json 2>&1 | tail -8
for kind in camera general; do
  timeout 300 python tools/bench_full.py --kind $kind --rays 262144 --steps 512 --out gpurun_out/full_${kind}_$TAG.json 2>&1 | tail -2
done
timeout 300 python tools/bench_full.py --kind camera --rays 4096 --steps 8192 --out gpurun_out/full_camera_4096x8192_$TAG.json 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${TAG}_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n1.json").read().strip().splitlines()[-1])
print("frame", d["frame_ms"], "e2e", d["e2e"]["frame_ms"], "frac", d["roofline"]["frac"], d.get("other_configs"))
PY
