#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_driver.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
( export CUDA_VISIBLE_DEVICES=0; timeout 900 python bench.py --gpus 1 --steps 1 --warmup 2 --no-cpu-baseline --mode exact > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err )
cut -c1-120 gpurun_out/${tag}_bench_n1.json; grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_bench_n1.json; tail -2 gpurun_out/${tag}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 2 --steps 1 --warmup 2 --mode exact > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "n2 exit $?"; cut -c1-120 gpurun_out/${tag}_bench_n2.json; grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_bench_n2.json; grep -o '"checksums": {[^}]*}' gpurun_out/${tag}_bench_n2.json | cut -c1-200; tail -2 gpurun_out/${tag}_bench_n2.err
