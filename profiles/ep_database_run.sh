#!/bin/bash
# BASELINE.json configs[4]: all five K-quant levels of a random-init Llama back to back (one process per level and GPU, replicas,
# no collective) + the EvoPress layer database, timed.  Offline: random-init weights from a config, synthetic calibration tokens.
#   bash profiles/ep_database_run.sh 8b  <tag>     Llama-3-8B, 128 x 2048 tokens, levels Q2_K..Q6_K     (8-GPU box)
#   bash profiles/ep_database_run.sh tiny <tag>    plumbing check on one GPU (2 levels, d_model 256)
size=${1:-tiny}; tag=${2:-ep}
mkdir -p gpurun_out
work=/dev/shm/ep_run_$tag; rm -rf "$work"; mkdir -p "$work/model"
if [ "$size" = 8b ]; then
  cfg='{"architectures":["LlamaForCausalLM"],"model_type":"llama","hidden_size":4096,"intermediate_size":14336,"num_hidden_layers":32,"num_attention_heads":32,"num_key_value_heads":8,"vocab_size":128256,"max_position_embeddings":8192,"rms_norm_eps":1e-05,"rope_theta":500000.0,"tie_word_embeddings":false,"torch_dtype":"bfloat16"}'
  levels="Q2_K Q3_K Q4_K Q5_K Q6_K"; tokens=262144; seq=2048; vocab=128256; hf="--no_hf_layers"
else
  cfg='{"architectures":["LlamaForCausalLM"],"model_type":"llama","hidden_size":256,"intermediate_size":768,"num_hidden_layers":2,"num_attention_heads":4,"num_key_value_heads":2,"vocab_size":512,"max_position_embeddings":256,"rms_norm_eps":1e-05,"tie_word_embeddings":false,"torch_dtype":"bfloat16"}'
  levels="Q4_K Q6_K"; tokens=2048; seq=128; vocab=512; hf=""
fi
echo "$cfg" > "$work/model/config.json"
df -h /dev/shm | tail -1
t0=$(date +%s.%N)
MODEL="$work/model" RANDOM_INIT_CONFIG="$work/model/config.json" CALIB_DATA="random:$vocab" CALIB_TOKENS=$tokens CALIB_SEQ_LEN=$seq \
  DTYPE=bfloat16 OUT="$work/out" LEVELS="$levels" EP_DATABASE_ARGS="$hf" \
  bash gptq_gguf_toolkit_b200/build_ep_database.sh > gpurun_out/${tag}_ep_database.log 2>&1
rc=$?
t1=$(date +%s.%N)
{
  echo "exit $rc  wall $(awk "BEGIN{print $t1 - $t0}") s  size=$size levels=[$levels] gpus=${CUDA_VISIBLE_DEVICES:-all}"
  grep -h "Quantization took" "$work"/out/logs/*.log
  grep -h "quantise_wall_s\|database_wall_s" gpurun_out/${tag}_ep_database.log
  du -sh "$work"/out/* 2>/dev/null | tail -8
  ls "$work/out/ep_database/layers-gguf" 2>/dev/null | wc -l
} | tee gpurun_out/${tag}_ep_database_summary.txt
tail -5 gpurun_out/${tag}_ep_database.log
rm -rf "$work"
