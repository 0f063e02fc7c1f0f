#!/bin/bash
# Final 1-GPU call of the round: full GPU suite, the default bench line (+ left-looking ablation), the reference arm,
# the ncu launch list and the --set full captures.
tag=${1:-r01z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 1200 python bench.py --compare-left 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cut -c1-200 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
echo "reference exit $?"; cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 600 python bench.py --batch 16 --steps 1 --warmup 2 --no-e2e --mode exact --no-cpu-baseline > gpurun_out/${tag}_bench_batch16.json 2> gpurun_out/${tag}_bench_batch16.err
echo "batch16 exit $?"; cut -c1-160 gpurun_out/${tag}_bench_batch16.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_batch16.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
bash profiles/profile.sh ${tag} > gpurun_out/${tag}_profile.log 2>&1
tail -12 gpurun_out/${tag}_profile.log
