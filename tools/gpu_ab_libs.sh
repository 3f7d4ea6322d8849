#!/bin/bash
# A/B of two builds on one box with tools/probe_ab_clip.py (alternating processes), then a test subset on the
# default library: tools/gpu_ab_libs.sh <tag> <libA> <libB> [rounds] [pytest -k expression]
TAG=$1; LA=$2; LB=$3; N=${4:-2}; K=${5:-}
mkdir -p gpurun_out
for i in $(seq 1 $N); do
  for L in A B; do
    LIB=$LA; [ $L = B ] && LIB=$LB
    BSVD_B200_LIB=$PWD/$LIB timeout 200 python tools/probe_ab_clip.py 2>gpurun_out/${TAG}_$L$i.err | tail -1 | tee gpurun_out/${TAG}_$L$i.json | cut -c1-330
  done
done
if [ -n "$K" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -4; fi
