#!/bin/bash
# A/B of library variants under variants/*.so against the default build:  bash scripts/gpu_ab.sh <tag> "<quick_gpu stages>"
TAG=${1:-ab}; STAGES=${2:-"stolt filters"}
O=gpurun_out; mkdir -p $O
{
echo "== default"; timeout 300 python scripts/quick_gpu.py $STAGES 2>&1 | grep -v Warning
for lib in variants/*.so; do
  echo "== $lib"; IMPDAR_B200_LIB=$PWD/$lib timeout 300 python scripts/quick_gpu.py $STAGES 2>&1 | grep -v Warning
done
} | tee $O/ab_$TAG.txt
