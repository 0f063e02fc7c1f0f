#!/usr/bin/env bash
# Launcher for `python -m gptq_gguf_toolkit_b200.quant`, one process per visible GPU (rendezvous on 127.0.0.1).
#
# It accepts what a user of the reference's quant/gptq/run_quant.sh already has in their shell: the quantisation level as
# the first argument and the same environment variables with the same defaults.  The table below maps every variable to the
# quant.py flag it feeds; empty-by-default variables are passed as empty strings exactly like the reference does.
# Differences: BIT_WIDTH_CONFIGURATION is optional here (the reference insists on ./config.json, quant.py:204-205),
# --eval_perplexity is not part of this package (eval/ is out of scope), RANDOM_INIT_CONFIG=<json> builds a random-init model
# from a config instead of loading a checkpoint (offline benchmarking).
set -euo pipefail

level="${1:-Q4_K}"
export OMP_NUM_THREADS="${OMP_NUM_THREADS:-8}"

IFS=',' read -r -a visible <<< "${CUDA_VISIBLE_DEVICES:-0}"
ranks="${#visible[@]}"
port="${MASTER_PORT:-29500}"

cd "$(dirname "${BASH_SOURCE[0]}")/.."

#        flag                            variable        default
table=( "model_name_or_path              MODEL           meta-llama/Llama-3.2-1B-Instruct"
        "tokenizer_name                  TOKENIZER_NAME  "
        "calibration_data                CALIB_DATA      fineweb_edu"
        "calibration_tokens              CALIB_TOKENS    4194304"
        "calibration_sequence_length     CALIB_SEQ_LEN   4096"
        "quant_scale                     QUANT_SCALE     absmax"
        "rel_damp                        REL_DAMP        0.01"
        "block_size                      BLOCK_SIZE      128"
        "rmin                            RMIN            -1.0"
        "rdelta                          RDELTA          0.1"
        "nstep                           NSTEP           20"
        "dtype                           DTYPE           auto"
        "seed                            SEED            0"
        "attn_implementation             ATTN_IMPL       "
        "eval_sequence_length            EVAL_SEQ_LEN    2048"
        "save_dir                        SAVE_DIR        ./quantized_model" )

args=( --default_bit_width "$level"
       --quantizable_modules '.*layers.*((q|k|v|o|gate|up|down)_proj)$'
       --pre_block_modules model.embed_tokens --block_modules model.layers --post_block_modules lm_head
       --quant_non_block_modules --verbose )
for row in "${table[@]}"; do
    read -r flag var default <<< "$row"
    args+=( "--$flag" "${!var:-${default:-}}" )
done
[[ -n "${BIT_WIDTH_CONFIGURATION:-}" ]] && args+=( --bit_width_configuration "$BIT_WIDTH_CONFIGURATION" )
[[ -n "${RANDOM_INIT_CONFIG:-}" ]] && args+=( --random_init_config "$RANDOM_INIT_CONFIG" )

exec python -m torch.distributed.run --nnodes=1 --nproc-per-node="$ranks" --master-addr 127.0.0.1 --master-port "$port" \
    -m gptq_gguf_toolkit_b200.quant "${args[@]}"
