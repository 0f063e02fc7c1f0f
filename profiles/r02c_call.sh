#!/bin/bash
# Round 2, GPU call 3: chol_diag_v4 (two-level diagonal kernel), grouped trailing updates of the Cholesky, f16 GEMM as the chain's
# default; the whole GPU suite on the new defaults.
tag=${1:-r02c}
mkdir -p gpurun_out
( cd profiles/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o chol_diag_v4 chol_diag_v4.cu \
  && timeout 120 ./chol_diag_v4 ) > gpurun_out/${tag}_chol_diag_v4.log 2>&1
cat gpurun_out/${tag}_chol_diag_v4.log
timeout 600 python profiles/micro.py prepare > gpurun_out/${tag}_micro.log 2>&1
cat gpurun_out/${tag}_micro.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "gq_prepare n=" gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
