#!/bin/bash
# X1: the reference's own Numba kernel on the B200 next to ours (oracle/ref_numba_gpu.py).  Run: gpurun -- 'bash tools/gpu_ref_numba.sh'
cd "$(dirname "$0")/.."
timeout 900 python oracle/ref_numba_gpu.py --frames 400x400 1920x1080 3840x2160 --out gpurun_out/reference_numba_gpu.json > gpurun_out/ref_numba.log 2> gpurun_out/ref_numba.err
echo "rc=$?"; tail -5 gpurun_out/ref_numba.log; tail -5 gpurun_out/ref_numba.err
timeout 900 python oracle/ref_numba_gpu_next_rows.py --out gpurun_out/reference_numba_gpu_next_rows.json > gpurun_out/ref_numba_next.log 2> gpurun_out/ref_numba_next.err
echo "next rows rc=$?"; tail -c 1500 gpurun_out/ref_numba_next.log; tail -3 gpurun_out/ref_numba_next.err
