"""CLI of the GPTQ -> K-quant stage -- same flag surface as the reference's quant/gptq/quant.py:18-142
(so run_quant.sh and downstream scripts keep working), driving the libgq-backed Quantizer.

    torchrun --nnodes=1 --nproc-per-node=$NUM_GPUS -m gptq_gguf_toolkit_b200.quant --model_name_or_path ... (see run_quant.sh)

Extra flags: --mode exact|fast, --calibration_batch_size, --random_init_config (build a random-init model from a config.json /
LlamaConfig kwargs instead of loading weights; used by the offline benchmark), --no_share_hessians.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from .data_utils import get_data
from .model_utils import fix_seed
from .quant_utils import GGMLQuantizationType
from .quantizer import PhaseTimer, Quantizer

BIT_WIDTHS = ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"]


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    # Model params
    p.add_argument("--model_name_or_path", type=str, required=True, help="The name or path to quantized model.")
    p.add_argument("--tokenizer_name", type=str, default=None)
    p.add_argument("--quantizable_modules", type=str, required=True, help="Regex for modules to quantize")
    p.add_argument("--pre_block_modules", nargs="+", type=str, required=True)
    p.add_argument("--block_modules", type=str, required=True)
    p.add_argument("--post_block_modules", nargs="+", type=str, default=[])
    p.add_argument("--quant_non_block_modules", action="store_true")
    # Data params
    p.add_argument("--calibration_data", type=str, required=True)
    p.add_argument("--calibration_tokens", default=int(2**20), type=int)
    p.add_argument("--calibration_sequence_length", default=None, type=int)
    # Quantization params
    p.add_argument("--quant_scale", type=str, default="absmax", choices=["absmax", "mse"])
    p.add_argument("--act_order", action="store_true")
    p.add_argument("--static_groups", action="store_true")
    p.add_argument("--rel_damp", type=float, default=1e-2)
    p.add_argument("--block_size", type=int, default=128)
    p.add_argument("--default_bit_width", type=str, default="Q4_K")
    p.add_argument("--bit_width_configuration", type=str, default=None)
    # K-Scales params
    p.add_argument("--rmin", type=float, default=-1.0)
    p.add_argument("--rdelta", type=float, default=0.1)
    p.add_argument("--nstep", type=int, default=20)
    # Logging params
    p.add_argument("--log_wandb", default=False, action="store_true")
    # Misc params
    p.add_argument("--dtype", type=str, default="auto", choices=["auto", "float16", "float32", "bfloat16"])
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--low_cpu_mem_usage", action="store_true")
    p.add_argument("--attn_implementation", type=str, default=None, choices=["eager", "sdpa", "flash_attention_2", ""])
    p.add_argument("--cpu_offload_modules", action="store_true")
    p.add_argument("--cpu_offload_activations", action="store_true")
    p.add_argument("--eval_perplexity", action="store_true")
    p.add_argument("--eval_sequence_length", type=int, default=4096)
    p.add_argument("--verbose", action="store_true")
    # Save params
    p.add_argument("--save_dir", type=str, required=True)
    # ---- additions ----
    p.add_argument("--calibration_batch_size", type=int, default=8)
    p.add_argument("--random_init_config", type=str, default=None,
                   help="JSON file with LlamaConfig kwargs: build a random-init model instead of loading weights")
    p.add_argument("--no_share_hessians", action="store_true")
    p.add_argument("--mode", type=str, default="exact", choices=["exact", "fast"],
                   help="exact: the reference's fp32 rank-k arithmetic, bit-identical bytes given (W, U); fast: rank-k updates on "
                        "tcgen05 tensor cores (fp32-class accuracy, ~2x faster column loops)")
    p.add_argument("--rtn_fp32_arith", action="store_true",
                   help="embed_tokens / lm_head of a 16-bit model: widen the weights to fp32 for the scale search instead of "
                        "searching in the weight's own arithmetic like the reference does")
    return p.parse_args(argv)


def build_quant_config(default_bit_width, bit_width_configuration):
    """quant.py:183-217: uniform config from --default_bit_width, REPLACED wholesale by the JSON file if given."""
    if default_bit_width is None and bit_width_configuration is None:
        raise ValueError("Either default_bit_width or bit_width_configuration must be provided.")
    quant_config = None
    if default_bit_width is not None:
        if default_bit_width not in BIT_WIDTHS:
            raise ValueError("default_bit_width must be one of [Q2_K, Q3_K, Q4_K, Q5_K, Q6_K]")
        bw = GGMLQuantizationType[default_bit_width]
        quant_config = {k: bw for k in ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "down_proj", "up_proj",
                                        "embed_tokens", "lm_head"]}
    if bit_width_configuration is not None:
        if not os.path.isfile(bit_width_configuration):
            raise ValueError("bit_width_configuration must be a valid file path.")
        with open(bit_width_configuration, "r") as f:
            cfg = json.load(f)
        quant_config = {}
        for key, value in cfg.items():
            if value not in BIT_WIDTHS:
                raise ValueError("All bit widths in bit_width_configuration must be one of [Q2_K, Q3_K, Q4_K, Q5_K, Q6_K]")
            quant_config[key] = GGMLQuantizationType[value]
    return quant_config


def load_model(args, device):
    from transformers import AutoModelForCausalLM, LlamaConfig, LlamaForCausalLM
    if args.random_init_config:
        with open(args.random_init_config) as f:
            cfg = LlamaConfig(**json.load(f))
        dtype = {"auto": torch.bfloat16, "float16": torch.float16, "float32": torch.float32, "bfloat16": torch.bfloat16}[args.dtype]
        torch.manual_seed(args.seed)
        with torch.device(device):
            model = LlamaForCausalLM(cfg).to(dtype)
        return model.eval()
    model = AutoModelForCausalLM.from_pretrained(
        args.model_name_or_path, trust_remote_code=True, torch_dtype=args.dtype,
        low_cpu_mem_usage=args.low_cpu_mem_usage, attn_implementation=args.attn_implementation or None)
    if not args.cpu_offload_modules:
        model = model.to(device)
    return model.eval()


def main(argv=None):
    args = parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("gptq_gguf_toolkit_b200 needs a CUDA device (no CPU fallback)")
    distributed = "RANK" in os.environ and "WORLD_SIZE" in os.environ
    if distributed:
        dist.init_process_group(backend="nccl", init_method="env://")
    world_size = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(device)

    model = load_model(args, device)
    tokenizer = None
    if not (os.path.isfile(args.calibration_data) or args.calibration_data.startswith("random:")):
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(args.tokenizer_name or args.model_name_or_path, use_fast=False)
    args.calibration_sequence_length = args.calibration_sequence_length or model.config.max_position_embeddings
    calibration_data = get_data(args.calibration_data, args.calibration_tokens, args.calibration_sequence_length, tokenizer, train=True)
    if distributed:                                                                  # quant.py:177-179
        n = len(calibration_data) // world_size
        calibration_data = calibration_data[rank * n:(rank + 1) * n]
    calibration_data = [([], {"input_ids": ids}) for ids in calibration_data]
    quant_config = build_quant_config(args.default_bit_width, args.bit_width_configuration)

    timer = PhaseTimer(args.verbose)
    quantizer = Quantizer(
        model, data_loader=calibration_data, quantizable_modules=args.quantizable_modules,
        quantizer_kwargs=dict(rel_damp=args.rel_damp, block_size=args.block_size, act_order=args.act_order,
                              quant_scale=args.quant_scale, static_groups=args.static_groups, rmin=args.rmin,
                              rdelta=args.rdelta, nstep=args.nstep, verbose=args.verbose, mode=args.mode),
        pre_block_modules=args.pre_block_modules, block_modules=args.block_modules,
        post_block_modules=args.post_block_modules, quant_non_block_modules=args.quant_non_block_modules,
        cpu_offload_modules=args.cpu_offload_modules, cpu_offload_activations=args.cpu_offload_activations,
        device=device, verbose=args.verbose, save_dir=args.save_dir,
        calibration_batch_size=args.calibration_batch_size, share_hessians=not args.no_share_hessians, timer=timer,
        rtn_native_arith=not args.rtn_fp32_arith)
    if rank == 0:
        os.makedirs(args.save_dir, exist_ok=True)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    quantizer.quantize(quant_config)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        print(f"Quantization took {(t2 - t1)} s.")                                  # quant.py:254
        if args.verbose:
            print("phase seconds:", json.dumps({k: round(v, 3) for k, v in timer.totals().items()}))
        bad = quantizer.non_invertible_modules()
        if bad:
            print(f"WARNING: non-invertible Hessian (identity fallback) for: {bad}")
    if args.eval_perplexity and rank == 0:
        print("--eval_perplexity needs the wikitext-2 dataset (network); skipped in this build.")
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
