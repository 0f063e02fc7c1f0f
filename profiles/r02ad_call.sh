#!/bin/bash
# Round 2, session 3, call 4: ncu --set full (with source counters) of the panel launch after the three changes.
tag=${1:-r02ad}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gptq_layer_kernel' -s 8 -c 1 -f -o $OUT/${tag}_panel python profiles/ncu_targets.py gptq > $OUT/${tag}_ncu_panel.log 2>&1; tail -2 $OUT/${tag}_ncu_panel.log
