#!/bin/bash
# 8-GPU box: scaling check N = 8 and N = 4 (exact mode, no e2e at 4; e2e at 8)
tag=${1:-r02h}
mkdir -p gpurun_out
for n in 8 4; do
  extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n \
      bench.py --gpus $n --steps 1 --warmup 2 --mode exact $extra > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  echo "n$n exit $?"
  grep -o '"value": [0-9.]*' gpurun_out/${tag}_bench_n$n.json | head -1; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_n$n.json; grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_bench_n$n.json
  tail -2 gpurun_out/${tag}_bench_n$n.err
done
