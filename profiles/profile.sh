#!/usr/bin/env bash
# Profiling recipe (run under gpurun on one B200):  bash profiles/profile.sh <tag>
# 1. launch list with device times of every kernel of one dev-workload step (cold-cache, serialised: compare SHARES)
# 2. one `--set full` capture each of the top kernels
set -uo pipefail
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --workload llama3-8b-dev --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --mode exact"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
for K in gptq_layer_kernel hessian_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/${TAG}_$K $BENCH > $OUT/${TAG}_$K.log 2>&1
done
ls -la $OUT
