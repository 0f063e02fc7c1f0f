#!/usr/bin/env bash
# run_quant.sh -- launcher with the same environment-variable surface as the reference's
# quant/gptq/run_quant.sh (BITS as $1; MODEL, CALIB_DATA, CALIB_TOKENS, CALIB_SEQ_LEN, QUANT_SCALE,
# REL_DAMP, BLOCK_SIZE, BIT_WIDTH_CONFIGURATION, RMIN, RDELTA, NSTEP, DTYPE, SEED, ATTN_IMPL, SAVE_DIR).
# One process per GPU; rendezvous on 127.0.0.1.
set -euo pipefail

BITS=${1:-Q4_K}
export OMP_NUM_THREADS=8
NUM_GPUS=$(echo "${CUDA_VISIBLE_DEVICES:-0}" | tr ',' '\n' | wc -l)
MASTER_PORT="${MASTER_PORT:-29500}"
REPO_ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
cd "$REPO_ROOT"

# the reference always passes ./config.json and fails if it is missing (run_quant.sh:30, quant.py:204-205);
# here the JSON is optional: without it --default_bit_width applies to every projection.
CFG_ARGS=()
if [[ -n "${BIT_WIDTH_CONFIGURATION:-}" ]]; then CFG_ARGS=(--bit_width_configuration "$BIT_WIDTH_CONFIGURATION"); fi

python -m torch.distributed.run --nnodes=1 --nproc-per-node="$NUM_GPUS" --master-addr 127.0.0.1 --master-port "$MASTER_PORT" \
    -m gptq_gguf_toolkit_b200.quant \
    --model_name_or_path "${MODEL:-meta-llama/Llama-3.2-1B-Instruct}" \
    --tokenizer_name "${TOKENIZER_NAME:-}" \
    --quantizable_modules '.*layers.*((q|k|v|o|gate|up|down)_proj)$' \
    --pre_block_modules model.embed_tokens \
    --block_modules model.layers \
    --post_block_modules lm_head \
    --quant_non_block_modules \
    --calibration_data "${CALIB_DATA:-fineweb_edu}" \
    --calibration_tokens "${CALIB_TOKENS:-4194304}" \
    --calibration_sequence_length "${CALIB_SEQ_LEN:-4096}" \
    --quant_scale "${QUANT_SCALE:-absmax}" \
    --rel_damp "${REL_DAMP:-0.01}" \
    --block_size "${BLOCK_SIZE:-128}" \
    --default_bit_width "${BITS:-Q4_K}" \
    "${CFG_ARGS[@]}" \
    --rmin "${RMIN:--1.0}" \
    --rdelta "${RDELTA:-0.1}" \
    --nstep "${NSTEP:-20}" \
    --dtype "${DTYPE:-auto}" \
    --seed "${SEED:-0}" \
    --attn_implementation "${ATTN_IMPL:-}" \
    --eval_sequence_length "${EVAL_SEQ_LEN:-2048}" \
    --verbose \
    --save_dir "${SAVE_DIR:-./quantized_model}" \
    ${RANDOM_INIT_CONFIG:+--random_init_config "$RANDOM_INIT_CONFIG"}
