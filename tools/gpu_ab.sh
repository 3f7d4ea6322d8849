#!/bin/bash
# quick A/B on the GPU box: tests, then bench lines under different debug env switches
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "fps %.1f e2e %.1f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
print("  stage_ms", d["stage_ms"])
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_a.json 2> gpurun_out/${TAG}_bench_a.err; show gpurun_out/${TAG}_bench_a.json
BSVD_B200_NO_SKIP_PF=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_b.json 2> gpurun_out/${TAG}_bench_b.err; show gpurun_out/${TAG}_bench_b.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c.json 2> gpurun_out/${TAG}_bench_c.err; show gpurun_out/${TAG}_bench_c.json
