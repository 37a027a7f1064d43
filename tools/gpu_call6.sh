#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r02e}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python tools/step_probe.py --out gpurun_out/step_probe_$TAG.json 2>&1 | tail -12 | cut -c1-330
for kind in camera general; do
  timeout 300 python tools/bench_full.py --kind $kind --steps 2048 --out gpurun_out/full_${kind}_$TAG.json 2>&1 | tail -1
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${TAG}_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n1.json").read().strip().splitlines()[-1])
print("frame", d["frame_ms"], "e2e", d["e2e"]["frame_ms"], "frac", d["roofline"]["frac"], d["roofline"]["sm_cycles_per_warp_step"], d.get("other_configs"), d.get("dedup_opt_in",{}).get("frame_ms"))
PY
