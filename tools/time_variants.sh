#!/bin/bash
# Time every built launch variant (tools/build_variants.sh) on one frame; run on the GPU box.
# usage: tools/time_variants.sh [height width]
cd "$(dirname "$0")/.."
H=${1:-2160}; W=${2:-3840}
for so in schwarzchild-einsteinpy-raytracing_b200/lib/libschw_b200.so schwarzchild-einsteinpy-raytracing_b200/lib/variants/*.so; do
  echo "== $so"
  SCHW_B200_LIB=$PWD/$so timeout 300 python tools/profile_render.py --height $H --width $W --reps 3 | grep render | tail -2
done
