#!/bin/bash
# One gpurun call on a 2-GPU box: GPU parity suite (incl. the NCCL world-2 test), schedule / diagonal-kernel
# micro-benchmarks, 1-GPU bench steps (two-CTA panel build on / off) and a 2-GPU bench step.
# Everything lands in gpurun_out/<tag>_*.
tag=${1:-r01m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${tag}_gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
B="--gpus 1 --steps 1 --warmup 2 --no-e2e --mode exact --no-cpu-baseline"
( export CUDA_VISIBLE_DEVICES=0
  timeout 600 python profiles/micro.py prepare schedules > gpurun_out/${tag}_micro.log 2>&1
  timeout 900 python bench.py $B > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err ) &
( export CUDA_VISIBLE_DEVICES=1 GQ_PANEL_2CTA=0
  timeout 600 python profiles/micro.py schedules > gpurun_out/${tag}_micro_1cta.log 2>&1
  timeout 900 python bench.py $B > gpurun_out/${tag}_bench_n1_1cta.json 2> gpurun_out/${tag}_bench_n1_1cta.err ) &
wait
cat gpurun_out/${tag}_micro.log
echo "--- GQ_PANEL_2CTA=0"; cat gpurun_out/${tag}_micro_1cta.log
for f in bench_n1 bench_n1_1cta; do cut -c1-160 gpurun_out/${tag}_$f.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_$f.json; tail -2 gpurun_out/${tag}_$f.err; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 2 --steps 1 --warmup 2 --no-e2e --mode exact > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "n2 exit $?"
cut -c1-160 gpurun_out/${tag}_bench_n2.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_n2.json
tail -3 gpurun_out/${tag}_bench_n2.err
