#!/bin/bash
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 \
    bench.py --gpus 8 --steps 2 --warmup 2 --mode exact > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
echo "n8 exit $?"
grep -o '"value": [0-9.]*' gpurun_out/${tag}_bench_n8.json | head -1; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_n8.json; grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_bench_n8.json
tail -2 gpurun_out/${tag}_bench_n8.err
