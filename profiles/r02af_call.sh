#!/bin/bash
# Round 2, session 3, call 6: the whole GPU suite, smoke(), the bench line of both arms on the code with the reworked panel kernel.
tag=${1:-r02af}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${tag}_smoke.log 2>&1; tail -1 $OUT/${tag}_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 > $OUT/${tag}_bench_n1.json 2> $OUT/${tag}_bench_n1.err; cut -c1-160 $OUT/${tag}_bench_n1.json; tail -2 $OUT/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 1 --warmup 1 > $OUT/${tag}_bench_reference.json 2> $OUT/${tag}_bench_reference.err; cut -c1-200 $OUT/${tag}_bench_reference.json
