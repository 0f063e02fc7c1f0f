#!/bin/bash
# Round 2, GPU call 2: first run of the split-fp16 tcgen05 GEMM (csrc/gemm_f16x3.cu) -- unit tests against fp64, GQ_MODE_FAST on
# it, gq_prepare accuracy at n = 4096 / 14336 for both GEMM back-ends, micro-timings, one ncu --set full capture.
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tf32.py tests/test_gpu_prepare_large.py "tests/test_gpu_parity.py" -m gpu -q -s -k "gemm or prepare or fast" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error|gq_prepare n=|fast \(group|Error|assert" gpurun_out/${tag}_pytest.log | tail -60
timeout 600 python profiles/micro.py gptq prepare > gpurun_out/${tag}_micro.log 2>&1
grep -v "phase cycles" gpurun_out/${tag}_micro.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_f16x3_kernel' -s 20 -c 3 -f -o gpurun_out/${tag}_f16x3 \
    python profiles/ncu_targets.py fast > gpurun_out/${tag}_ncu_f16x3.log 2>&1
tail -3 gpurun_out/${tag}_ncu_f16x3.log
ls -la gpurun_out | tail -8
