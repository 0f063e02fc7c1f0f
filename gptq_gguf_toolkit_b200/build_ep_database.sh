#!/usr/bin/env bash
# EvoPress layer database for several quantisation levels in one go (BASELINE.json configs[4]: all five levels Q2_K..Q6_K on
# one 8-GPU box).  What the reference does with mapper/build_ep_database.sh -- write a .gguf per level, split it again -- is
# done here straight from the GPTQ results (gptq_gguf_toolkit_b200/ep_database.py, byte-identical to the reference splitter's
# output).  The levels are independent pipelines (only block 0's Hessians are level-independent), so they run as REPLICAS:
# one process per level, each on its own GPU of CUDA_VISIBLE_DEVICES (round-robin if there are fewer GPUs than levels, in
# waves); no collective is involved.
#
#   MODEL=<hf dir>  OUT=./ep_run  LEVELS="Q2_K Q3_K Q4_K Q5_K Q6_K"  bash gptq_gguf_toolkit_b200/build_ep_database.sh
#
# Every other variable of run_quant.sh (CALIB_DATA, CALIB_TOKENS, REL_DAMP, ...) is passed through unchanged.
set -euo pipefail

: "${MODEL:?set MODEL to the HF model directory}"
out="${OUT:-./ep_run}"
read -r -a levels <<< "${LEVELS:-Q2_K Q3_K Q4_K Q5_K Q6_K}"
IFS=',' read -r -a gpus <<< "${CUDA_VISIBLE_DEVICES:-0}"
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$out/logs"

wave=()
finish_wave() {
    local failed=0
    for pid in "${wave[@]}"; do wait "$pid" || failed=1; done
    wave=()
    [[ $failed -eq 0 ]] || { echo "a quantisation run failed, see $out/logs" >&2; exit 1; }
}

t_start=$(date +%s.%N)
i=0
for level in "${levels[@]}"; do
    gpu="${gpus[$((i % ${#gpus[@]}))]}"
    echo "[$level] on GPU $gpu -> $out/$level"
    CUDA_VISIBLE_DEVICES="$gpu" MASTER_PORT=$((29700 + i)) SAVE_DIR="$out/$level" MODEL="$MODEL" \
        bash "$here/run_quant.sh" "$level" > "$out/logs/$level.log" 2>&1 &
    wave+=("$!")
    i=$((i + 1))
    (( i % ${#gpus[@]} == 0 )) && finish_wave
done
finish_wave

t_quant=$(date +%s.%N)
args=()
for level in "${levels[@]}"; do args+=( --dir_model_quant "$out/$level" ); done
# EP_DATABASE_ARGS: extra flags for the emitter, e.g. --no_hf_layers
# shellcheck disable=SC2086
python -m gptq_gguf_toolkit_b200.ep_database "$MODEL" "${args[@]}" --output_dir "$out/ep_database" ${EP_DATABASE_ARGS:-}
t_end=$(date +%s.%N)
echo "quantise_wall_s $(awk "BEGIN{print $t_quant - $t_start}")  database_wall_s $(awk "BEGIN{print $t_end - $t_quant}")"
echo "database in $out/ep_database"
