#!/bin/bash
# Round 2, session 3, call 2: A/B of the in-super-block update from shared memory (mid_update_smem; libgq_v2.so = the commit before),
# the column-loop GPU tests on the new library.
tag=${1:-r02ac}
OUT=gpurun_out
mkdir -p $OUT
GQ_LIB_PATH=$PWD/gptq_gguf_toolkit_b200/libgq_v2.so timeout 300 python profiles/micro.py gptq nofast > $OUT/${tag}_micro_v2.log 2>&1; tail -9 $OUT/${tag}_micro_v2.log
timeout 300 python profiles/micro.py gptq nofast > $OUT/${tag}_micro_new.log 2>&1; tail -9 $OUT/${tag}_micro_new.log
timeout 1500 python -m pytest tests -m gpu -q -x -k "parity or schedules or variants or llama_widths or driver" > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
