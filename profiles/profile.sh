#!/usr/bin/env bash
# Profiling recipe (run under gpurun on one B200):  bash profiles/profile.sh <tag>
# 1. launch list with device times of every kernel of one dev-workload step (cold-cache, serialised: compare SHARES)
# 2. `--set full` captures of the hot kernels from small stand-alone launches (profiles/ncu_targets.py)
set -uo pipefail
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# NOTE on cost: under ncu every launch takes ~0.17 s; the two-block dev workload (4565 launches) needed 13 minutes of box time
# in round 1.  The one-block workload below is the default; pass WORKLOAD=llama3-8b-dev for the two-block list.
BENCH="python bench.py --workload ${WORKLOAD:-llama3-8b-dev1} --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --mode exact"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'exact_update_kernel|gptq_layer_kernel' -s 8 -c 6 -f -o $OUT/${TAG}_colloop \
    python profiles/ncu_targets.py gptq > $OUT/${TAG}_ncu_colloop.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'chol_diag_v2_kernel|hessian_tc_kernel' -s 3 -c 3 -f -o $OUT/${TAG}_linalg \
    python profiles/ncu_targets.py prepare hessian > $OUT/${TAG}_ncu_linalg.log 2>&1
ls -la $OUT
