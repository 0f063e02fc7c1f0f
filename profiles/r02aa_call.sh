#!/bin/bash
# Round 2, session 3, call 1: A/B of the conversion-free K-quant search / column steps (libgq_base.so = the commit before),
# the GPU suite on the new library, one ncu --set full capture of the new panel kernel.
tag=${1:-r02aa}
OUT=gpurun_out
mkdir -p $OUT
GQ_LIB_PATH=$PWD/gptq_gguf_toolkit_b200/libgq_base.so timeout 300 python profiles/micro.py gptq nofast rtn > $OUT/${tag}_micro_base.log 2>&1; tail -12 $OUT/${tag}_micro_base.log
timeout 300 python profiles/micro.py gptq nofast rtn > $OUT/${tag}_micro_new.log 2>&1; tail -12 $OUT/${tag}_micro_new.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gptq_layer_kernel' -s 8 -c 1 -f -o $OUT/${tag}_panel python profiles/ncu_targets.py gptq > $OUT/${tag}_ncu_panel.log 2>&1; tail -2 $OUT/${tag}_ncu_panel.log
ls -la $OUT | grep ${tag}
