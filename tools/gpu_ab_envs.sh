#!/bin/bash
# A/B of one build under two environments with tools/probe_ab_clip.py (alternating processes), then an optional
# test subset under environment B: tools/gpu_ab_envs.sh <tag> "<envA>" "<envB>" [rounds] [pytest -k expression]
TAG=$1; EA=$2; EB=$3; N=${4:-2}; K=${5:-}
mkdir -p gpurun_out
for i in $(seq 1 $N); do
  for L in A B; do
    E=$EA; [ $L = B ] && E=$EB
    env $E timeout 200 python tools/probe_ab_clip.py 2>gpurun_out/${TAG}_$L$i.err | tail -1 | tee gpurun_out/${TAG}_$L$i.json | cut -c1-330
  done
done
if [ -n "$K" ]; then env $EB timeout 600 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -4; fi
