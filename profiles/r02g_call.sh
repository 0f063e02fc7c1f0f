#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 600 python profiles/micro.py prepare 2>&1 | grep -E "defaults|GROUP=1" > gpurun_out/${tag}_micro.log; cat gpurun_out/${tag}_micro.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
