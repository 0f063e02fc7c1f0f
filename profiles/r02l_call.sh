#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_multi.py tests/test_zz_gpu_driver_golden.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
B="--gpus 1 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline"
( export CUDA_VISIBLE_DEVICES=0; timeout 900 python bench.py $B --mode both > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err ) &
( export CUDA_VISIBLE_DEVICES=1; GQ_FAST_GROUP=4 timeout 900 python bench.py $B --mode fast > gpurun_out/${tag}_bench_n1_fast_g4.json 2> gpurun_out/${tag}_bench_n1_fast_g4.err ) &
wait
for f in bench_n1 bench_n1_fast_g4; do cut -c1-120 gpurun_out/${tag}_$f.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_$f.json | head -2; tail -2 gpurun_out/${tag}_$f.err; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/'+"r02l"+'_bench_n1.json') if l.startswith('{')][-1])
f=d.get('fast_mode',{})
print('fast value', f.get('value'), 'roofline', {k:v for k,v in f.get('roofline',{}).items() if k in ('achieved','frac','total_ms','launches','uncontended')})
print('exact roofline', {k:d['roofline'][k] for k in ('achieved','frac','total_ms_per_step','frac_of_simt_fp32_peak')}, d['roofline']['panel_kernel']['total_ms_per_step'])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 2 --steps 1 --warmup 2 --mode exact --no-e2e > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "n2 exit $?"; cut -c1-120 gpurun_out/${tag}_bench_n2.json; grep -o '"phases_s": {[^}]*}' gpurun_out/${tag}_bench_n2.json; tail -2 gpurun_out/${tag}_bench_n2.err
