#!/bin/bash
# Round 2, final 1-GPU call: the GPU suite, smoke(), the bench line (both arms), ncu launch lists and --set full captures.
tag=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${tag}_pytest.log; tail -3 $OUT/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${tag}_smoke.log 2>&1; tail -1 $OUT/${tag}_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 > $OUT/${tag}_bench_n1.json 2> $OUT/${tag}_bench_n1.err; cut -c1-160 $OUT/${tag}_bench_n1.json; tail -2 $OUT/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 1 --warmup 1 > $OUT/${tag}_bench_reference.json 2> $OUT/${tag}_bench_reference.err; cut -c1-200 $OUT/${tag}_bench_reference.json
BENCH="python bench.py --workload llama3-8b-dev1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${tag}_launches_exact.csv $BENCH --mode exact > $OUT/${tag}_launches_exact.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${tag}_launches_fast.csv $BENCH --mode fast > $OUT/${tag}_launches_fast.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_f16x3_kernel' -s 40 -c 4 -f -o $OUT/${tag}_f16x3 python profiles/ncu_targets.py fast > $OUT/${tag}_ncu_f16x3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'exact_update_kernel|gptq_layer_kernel' -s 8 -c 4 -f -o $OUT/${tag}_colloop python profiles/ncu_targets.py gptq > $OUT/${tag}_ncu_colloop.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'chol_diag_v4_kernel|hessian_tc_kernel' -s 3 -c 3 -f -o $OUT/${tag}_linalg python profiles/ncu_targets.py prepare hessian > $OUT/${tag}_ncu_linalg.log 2>&1
ls -la $OUT | grep ${tag}
