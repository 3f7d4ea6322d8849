#!/bin/bash
# same-box A/B of the c32 configuration (tools/probe_c32_stages.py) and of the BSVD-64 clip between a reference
# build (bsvd_b200/lib/libA.so) and the current library, then a test subset: tools/gpu_c32ab.sh [pytest -k]
mkdir -p gpurun_out
for i in 1 2; do
  BSVD_B200_LIB=$PWD/bsvd_b200/lib/libA.so timeout 120 python tools/probe_c32_stages.py 2>/dev/null | tail -1 | tee gpurun_out/c32ab_A$i.json | cut -c1-330
  timeout 120 python tools/probe_c32_stages.py 2>/dev/null | tail -1 | tee gpurun_out/c32ab_B$i.json | cut -c1-330
done
BSVD_B200_LIB=$PWD/bsvd_b200/lib/libA.so timeout 200 python tools/probe_ab_clip.py 2>/dev/null | tail -1 | tee gpurun_out/c64ab_A.json | cut -c1-330
timeout 200 python tools/probe_ab_clip.py 2>/dev/null | tail -1 | tee gpurun_out/c64ab_B.json | cut -c1-330
if [ -n "$1" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -3; fi
