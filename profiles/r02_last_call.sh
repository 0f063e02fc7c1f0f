#!/bin/bash
# Round 2, last 1-GPU call: the GPU suite and the bench line on the final code (cta_group::2 Hessian, adaptive cta_group::2 rank-k GEMM)
tag=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${tag}_smoke.log 2>&1; tail -1 $OUT/${tag}_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 > $OUT/${tag}_bench_n1.json 2> $OUT/${tag}_bench_n1.err; cut -c1-160 $OUT/${tag}_bench_n1.json; tail -2 $OUT/${tag}_bench_n1.err
