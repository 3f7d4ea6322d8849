#!/bin/bash
# One GPU-box visit: GPU tests, bench (both arms), ncu launch list + full capture of one forward.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python tools/bench_extra.py u8 --steps 10 2>/dev/null | tail -1 > gpurun_out/${TAG}_u8.json; cut -c1-200 gpurun_out/${TAG}_u8.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cut -c1-400 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; echo "ref exit $?"
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-300
T=10 N=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_launches_T10.csv python tools/run_once.py > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu launches exit $?"
T=10 N=2 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 32 --launch-count 32 -f -o /tmp/${TAG}_full python tools/run_once.py > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_T10.csv 2>/dev/null
python tools/ncu_summarize.py gpurun_out/${TAG}_ncu_raw_T10.csv gpurun_out/${TAG}_ncu_summary_T10.md gpurun_out/${TAG}_ncu_traffic.json | tail -2
