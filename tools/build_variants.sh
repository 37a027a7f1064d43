#!/bin/bash
# Build launch-configuration variants of libschw_b200.so for A/B timing (tools/profile_render.py with SCHW_B200_LIB=...).
# usage: tools/build_variants.sh "128 4" "128 5 1" ...   (threads-per-CTA  min-CTAs-per-SM  [trig memo in shared memory 0/1])
set -e
cd "$(dirname "$0")/../schwarzchild-einsteinpy-raytracing_b200/csrc"
mkdir -p ../lib/variants
[ $# -eq 0 ] && set -- "128 4" "128 5" "128 6" "96 7" "64 10" "64 12"
for v in "$@"; do
  set -- $v
  out=../lib/variants/libschw_b200_b$1_m$2_s${3:-0}.so
  /usr/local/cuda/bin/nvcc -DSCHW_BLOCK=$1 -DSCHW_MINBLOCKS=$2 -DSCHW_MEMO_SMEM=${3:-0} -Xptxas -v -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC --expt-relaxed-constexpr -shared -o $out schw_kernels.cu -lcudart 2>&1 | grep -A2 "persistent_raysINS_9RenderJobELb1ELb1" | grep -E "Used|spill" | sed -e 's/ptxas info    ://'
  echo "built $out"
done
