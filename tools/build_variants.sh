#!/bin/bash
# Build launch-configuration variants of libschw_b200.so for A/B timing (tools/profile_render.py with SCHW_B200_LIB=...).
set -e
cd "$(dirname "$0")/../schwarzchild-einsteinpy-raytracing_b200/csrc"
mkdir -p ../lib/variants
for v in "128 3" "128 4" "128 5" "256 2" "64 8"; do
  set -- $v
  out=../lib/variants/libschw_b200_b$1_m$2.so
  /usr/local/cuda/bin/nvcc -DSCHW_BLOCK=$1 -DSCHW_MINBLOCKS=$2 -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v -shared -o $out schw_kernels.cu -lcudart 2>&1 | grep -A1 "persistent_raysINS_9RenderJob" | grep -E "Used" | sed "s/^/b$1 m$2: /"
done
