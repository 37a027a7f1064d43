#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -k "full or rows or ragged or k3 or kats or integrator_api or manual_simulation" 2>&1 | tail -3 | cut -c1-300
echo "== trajectory kernel R=2 vs R=1, 454656 camera rays x 1024 rows"
python tools/bench_full.py --kind camera --rays 454656 --steps 1024 2>&1 | tail -1 | cut -c90-420
SCHW_FULL_RAYS=1 python tools/bench_full.py --kind camera --rays 454656 --steps 1024 2>&1 | tail -1 | cut -c90-420
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "schwarzchild-einsteinpy-raytracing_b200")
import schw_b200
hd = schw_b200.Handle(0)
sc = schw_b200.make_scene(height=2160, width=3840, has_texture=False)
out = hd.render(sc, None, 0, 2160, 1, extras=True)
ns = out["nsteps"]
print("4K frame: max steps of a ray", int(ns.max()), "99.99th pct", int(np.percentile(ns, 99.99)), "rays > 20000 steps", int((ns > 20000).sum()), "> 10000", int((ns > 10000).sum()))
PY
