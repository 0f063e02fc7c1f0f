#!/bin/bash
# First GPU call of round 2 (1 GPU, ~6 min): everything that was written after round 1's GPU budget was spent.
#   bash profiles/r02_first_call.sh r02a        (under gpurun; results in gpurun_out/<tag>_*)
# 1. the GPU suite incl. the end-of-round-1 additions (test_zz_gpu_driver_golden.py) and the opt-in experimental test
# 2. micro-benchmarks of the experimental variants: GQ_UPDATE_V2 (exact_update_v2_kernel), GQ_DIAG_V2=3, GQ_PREPARE_LOOKAHEAD
# 3. the stand-alone probe of the register-resident diagonal-block kernel
# Decide from the numbers which of them become defaults; nothing here changes the product path by itself.
tag=${1:-r02a}
mkdir -p gpurun_out
GQ_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
timeout 600 python profiles/micro.py prepare schedules experimental > gpurun_out/${tag}_micro_experimental.log 2>&1
cat gpurun_out/${tag}_micro_experimental.log
( cd profiles/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o chol_diag_v3 chol_diag_v3.cu \
  && timeout 120 ./chol_diag_v3 ) > gpurun_out/${tag}_chol_diag_v3.log 2>&1
cat gpurun_out/${tag}_chol_diag_v3.log
# end-to-end effect of the two library flags on the bench step (same box, back to back)
B="--gpus 1 --steps 1 --warmup 2 --no-e2e --mode exact --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
GQ_UPDATE_V2=1 timeout 600 python bench.py $B > gpurun_out/${tag}_bench_update_v2.json 2> gpurun_out/${tag}_bench_update_v2.err
GQ_PREPARE_LOOKAHEAD=1 timeout 600 python bench.py $B > gpurun_out/${tag}_bench_lookahead.json 2> gpurun_out/${tag}_bench_lookahead.err
GQ_DIAG_V2=3 timeout 600 python bench.py $B > gpurun_out/${tag}_bench_diag_v3.json 2> gpurun_out/${tag}_bench_diag_v3.err
for f in default update_v2 lookahead diag_v3; do echo "== $f"; cut -c1-140 gpurun_out/${tag}_bench_$f.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_$f.json; done
