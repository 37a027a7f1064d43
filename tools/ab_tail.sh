#!/bin/bash
# Rays per thread of the tile kernel against launch size on one GPU (the data behind tile_rays' threshold)
cd "$(dirname "$0")/.."
for cfg in "2160 3840 8" "2160 3840 12" "2160 3840 16" "1080 1920 2" "1080 1920 4" "1440 2560 8" "400 400 1"; do
  set -- $cfg
  for r in 1 2; do
    echo -n "h=$1 w=$2 stride=$3 R=$r: "
    SCHW_TILE_RAYS=$r python tools/profile_render.py --height $1 --width $2 --stride $3 --reps 5 | grep render | sed 's/.*kernel \([0-9.]*\) ms.*/\1/' | sort -n | head -1
  done
done
