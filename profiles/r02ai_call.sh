#!/bin/bash
# Round 2, session 3, call 9: the whole GPU suite and smoke() on the final tree.
tag=${1:-r02ai}
OUT=gpurun_out
mkdir -p $OUT
timeout 110 python -m pytest tests -m gpu -q > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${tag}_smoke.log 2>&1; tail -1 $OUT/${tag}_smoke.log
