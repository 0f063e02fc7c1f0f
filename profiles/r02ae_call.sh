#!/bin/bash
# Round 2, session 3, call 5: A/B of the serial loop's instruction diet (libgq_v3.so = the commit before), column-loop GPU tests.
tag=${1:-r02ae}
OUT=gpurun_out
mkdir -p $OUT
GQ_LIB_PATH=$PWD/gptq_gguf_toolkit_b200/libgq_v3.so timeout 300 python profiles/micro.py gptq nofast > $OUT/${tag}_micro_v3.log 2>&1; tail -9 $OUT/${tag}_micro_v3.log
timeout 300 python profiles/micro.py gptq nofast > $OUT/${tag}_micro_new.log 2>&1; tail -9 $OUT/${tag}_micro_new.log
timeout 1500 python -m pytest tests -m gpu -q -x -k "parity or schedules or variants or llama_widths or driver or f32x2" > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
