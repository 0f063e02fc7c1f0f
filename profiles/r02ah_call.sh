#!/bin/bash
# Round 2, session 3, call 8 (the last one): A/B of the two trailing-update kernels of the exact schedule, then the whole GPU
# suite with the 8 x 8-tile kernel as the default.
tag=${1:-r02ah}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python profiles/micro.py updab > $OUT/${tag}_micro_updab.log 2>&1; tail -13 $OUT/${tag}_micro_updab.log
timeout 400 python -m pytest tests -m gpu -q > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
