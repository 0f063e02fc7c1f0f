#!/bin/bash
# Round 2, GPU call 4 (2-GPU box): NCCL world-2 tests (chain owner + broadcast of U), full N=1 bench line (exact + fast + e2e +
# cpu_baseline) on GPU 0 while GPU 1 runs the mixed-config bench (BASELINE configs[2]), then the N=2 bench.
tag=${1:-r02d}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${tag}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${tag}_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_multi.log
tail -5 gpurun_out/${tag}_pytest_multi.log
( export CUDA_VISIBLE_DEVICES=0
  timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err ) &
( export CUDA_VISIBLE_DEVICES=1
  timeout 1200 python bench.py --gpus 1 --steps 1 --warmup 2 --mode exact --no-cpu-baseline --bit-width-configuration profiles/configs/mixed_q2k_q6k.json \
      > gpurun_out/${tag}_bench_mixed.json 2> gpurun_out/${tag}_bench_mixed.err ) &
wait
for f in bench_n1 bench_mixed; do cut -c1-200 gpurun_out/${tag}_$f.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_$f.json | head -2; tail -2 gpurun_out/${tag}_$f.err; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 2 --steps 1 --warmup 2 --mode exact > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "n2 exit $?"
cut -c1-200 gpurun_out/${tag}_bench_n2.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_n2.json
tail -3 gpurun_out/${tag}_bench_n2.err
