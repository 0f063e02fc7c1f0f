#!/bin/bash
# Round 2, session 3, call 7: ncu launch list of the one-block development workload (exact mode) and --set full captures of the
# column-loop kernels (panel launch after the rework, exact_update_kernel).
tag=${1:-r02ag}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --workload llama3-8b-dev1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${tag}_launches_exact.csv $BENCH --mode exact > $OUT/${tag}_launches_exact.log 2>&1; tail -1 $OUT/${tag}_launches_exact.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'exact_update_kernel|gptq_layer_kernel' -s 8 -c 4 -f -o $OUT/${tag}_colloop python profiles/ncu_targets.py gptq > $OUT/${tag}_ncu_colloop.log 2>&1; tail -2 $OUT/${tag}_ncu_colloop.log
ls -la $OUT | grep ${tag}
