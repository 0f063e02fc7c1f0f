#!/bin/bash
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${tag}_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_multi.log; tail -2 gpurun_out/${tag}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 2 --steps 2 --warmup 3 --mode exact > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "n2 exit $?"; cut -c1-120 gpurun_out/${tag}_bench_n2.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_n2.json; grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_bench_n2.json; tail -2 gpurun_out/${tag}_bench_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none -k regex:'hessian_tc_kernel' -s 1 -c 1 -f -o gpurun_out/${tag}_hessian python profiles/ncu_targets.py hessian > gpurun_out/${tag}_ncu_hessian.log 2>&1; tail -1 gpurun_out/${tag}_ncu_hessian.log
