#!/bin/bash
# A/B one build under two environments, alternating: tools/gpu_ab_env.sh <tag> "<envA>" "<envB>" [rounds]
TAG=$1; EA=$2; EB=$3; N=${4:-2}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "fps %.1f e2e %.1f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
print("  stage_ms", d["stage_ms"])
PY
}
for i in $(seq 1 $N); do
  for L in A B; do
    E=$EA; [ $L = B ] && E=$EB
    env $E timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${L}${i}.json 2> gpurun_out/${TAG}_${L}${i}.err; show gpurun_out/${TAG}_${L}${i}.json
  done
done
