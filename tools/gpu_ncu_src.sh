#!/bin/bash
# source-level (SASS + stall sampling) ncu pages of selected launches of the second forward
# usage: tools/gpu_ncu_src.sh <tag> <launch idx 0..31> [...]
TAG=$1; shift
mkdir -p gpurun_out
for L in "$@"; do
  SKIP=$((32 + L))
  T=10 N=2 timeout 600 ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count 1 -f -o /tmp/${TAG}_l${L} python tools/run_once.py > gpurun_out/${TAG}_l${L}.log 2>&1; echo "ncu l$L exit $?"
  ncu -i /tmp/${TAG}_l${L}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_l${L}_source.csv 2>/dev/null
  ncu -i /tmp/${TAG}_l${L}.ncu-rep --page raw --csv > gpurun_out/${TAG}_l${L}_raw.csv 2>/dev/null
  ls -la gpurun_out/${TAG}_l${L}_source.csv
done
