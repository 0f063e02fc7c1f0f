"""GPTQ K-quant results -> .gguf  (SURVEY §8f N1).

Host-side mirror of the ONE patch the reference makes to llama.cpp's converter
(quant/gptq/pack_gptq_into_gguf.py:282-349 `ModelBase.prepare_tensors`, CLI flag `--dir_model_quant`
:8816-8820): every HF tensor whose name (minus ".weight") is a directory of `<dir_model_quant>` is replaced by the
K-quant block bytes of that directory's data.pth, after the architecture's row permutation (Llama q/k,
:2177-2183, 2217-2221); everything else is written as in the upstream converter (norms F32, the rest `--outtype`).

It is NOT a copy of the vendored 9 k-line converter: it is a small Llama-family writer on gguf-py, enough for the
Llama configs of BASELINE.json, and it differs from the reference in one B200-first way: the GGUF block bytes are
already in data.pth (`packed`, emitted by the fused kernel), so the host `pack_Q*K` loops disappear -- packing is
row-local, so the q/k permutation is applied to rows of bytes (SURVEY §8a "K-quant block layouts").  data.pth files
written by the reference itself (no `packed` entry) are packed through gq_pack on the GPU.

    python -m gptq_gguf_toolkit_b200.pack_gptq_into_gguf <hf_model_dir> --dir_model_quant <save_dir> \
        --outfile model.gguf --outtype f16
"""
from __future__ import annotations

import argparse
import json
import math
import os
from hashlib import sha256
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch

import gguf

K_QUANTS = {10: gguf.GGMLQuantizationType.Q2_K, 11: gguf.GGMLQuantizationType.Q3_K, 12: gguf.GGMLQuantizationType.Q4_K,
            13: gguf.GGMLQuantizationType.Q5_K, 14: gguf.GGMLQuantizationType.Q6_K}
OUTTYPES = {"f32": (torch.float32, gguf.GGMLQuantizationType.F32, gguf.LlamaFileType.ALL_F32),
            "f16": (torch.float16, gguf.GGMLQuantizationType.F16, gguf.LlamaFileType.MOSTLY_F16)}


def llama_permute(w: torch.Tensor, n_head: int, n_head_kv: Optional[int]) -> torch.Tensor:
    """Row permutation llama.cpp applies to HF q_proj / k_proj (reference LlamaModel.permute, :2177-2183).
    Works on any tensor whose dim 0 is d_row -- weights, the five K-quant tensors, or rows of packed bytes."""
    if n_head_kv is not None and n_head != n_head_kv:
        n_head = n_head_kv
    return (w.reshape(n_head, 2, w.shape[0] // n_head // 2, *w.shape[1:]).swapaxes(1, 2).reshape(w.shape))


def _packed_bytes(obj: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(d_row, n_superblocks * type_size) uint8 for one data.pth dict (reference schema, quantizer.py:267-275)."""
    if obj.get("packed") is not None:
        return obj["packed"]
    from . import packing_utils          # reference-written data.pth: pack on the GPU (no CPU fallback)
    qt = int(obj["q_type"])
    fn = {10: packing_utils.pack_Q2K, 11: packing_utils.pack_Q3K, 12: packing_utils.pack_Q4K, 13: packing_utils.pack_Q5K,
          14: packing_utils.pack_Q6K}[qt]
    dev = "cuda"
    args = [obj["qweight"].to(dev), obj["super_group_scale"].to(dev), obj["group_scale_quant"].to(dev)]
    if qt in (10, 12, 13):
        args += [obj["super_group_zero"].to(dev), obj["group_zero_quant"].to(dev)]
    out = fn(*args)
    return torch.from_numpy(out) if isinstance(out, np.ndarray) else out.cpu()


def _iter_state(model_or_state) -> Iterable[Tuple[str, torch.Tensor]]:
    sd = model_or_state.state_dict() if hasattr(model_or_state, "state_dict") else model_or_state
    for name, t in sd.items():
        if name.endswith((".attention.masked_bias", ".attention.bias", ".rotary_emb.inv_freq")):
            continue
        yield name, t


# llama.cpp identifies a BPE pre-tokenizer by hashing the token ids of a fixed probe text (the reference's vendored converter,
# get_vocab_base_pre, quant/gptq/pack_gptq_into_gguf.py:721-949, carries ~100 such hashes).  The probe text and the hashes are
# DATA that must match llama.cpp's; this writer knows the Llama-family ones and refuses the rest (use the reference converter).
_PRE_PROBE = ('\n \n\n \n\n\n \t \t\t \t\n  \n   \n    \n     \n🚀 (normal) 😶\u200d🌫\ufe0f (multiple emojis concatenated) '
              '\u2705 🦙🦙 3 33 333 3333 33333 333333 3333333 33333333 3.3 3..3 3...3 '
              '\u1780\u17b6\u1793\u17cb\u178f\u17c2\u1796\u17b7\u179f\u17c1\u179f\u17a2\u17b6\u1785😁 ?\u6211\u60f3\u5728apple\u5de5\u4f5c1314151\u5929\uff5e '
              '------======= \u043d\u0435\u0449\u043e \u043d\u0430 \u0411\u044a\u043b\u0433\u0430\u0440\u0441\u043a\u0438 \'\'\'\'\'\'```````""""......!!!!!!?????? '
              'I\'ve been \'told he\'s there, \'RE you sure? \'M not sure I\'ll make it, \'D you like some tea? We\'Ve a\'lL')
_PRE_HASHES = {
    "0ef9807a4087ebef797fc749390439009c3b9eda9ad1a097abbe738f486c01e5": "llama-bpe",      # Meta-Llama-3 / 3.1 / 3.2
    "b6e8e1518dc4305be2fe39c313ed643381c4da5db34a98f6a04c093f8afbe99b": "qwen2",
}


def bpe_pre_tokenizer(hf_dir: str) -> str:
    """tokenizer.ggml.pre of a BPE model (reference set_vocab -> get_vocab_base_pre -> add_tokenizer_pre, :682, 957)."""
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(hf_dir)
    h = sha256(str(tok.encode(_PRE_PROBE)).encode()).hexdigest()
    if h not in _PRE_HASHES:
        raise NotImplementedError(f"BPE pre-tokenizer with probe hash {h} is not known to this writer; convert with the reference's "
                                  "quant/gptq/pack_gptq_into_gguf.py (it reads the same data.pth files)")
    return _PRE_HASHES[h]


def llama3_rope_factors(cfg: dict, rope_scaling: dict) -> torch.Tensor:
    """rope_freqs.weight of a Llama-3.1-style model (rope_type "llama3"): per frequency 1, `factor`, or the smooth blend
    between them, by wavelength against original_max_position_embeddings / {high, low}_freq_factor (the published Llama 3.1
    scaling rule; reference LlamaModel.generate_extra_tensors, :2259-2287)."""
    base = float(cfg.get("rope_theta") or rope_scaling.get("rope_theta") or 10000.0)
    dim = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
    freqs = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    factor = rope_scaling.get("factor", 8.0)
    lo_f, hi_f = rope_scaling.get("low_freq_factor", 1.0), rope_scaling.get("high_freq_factor", 4.0)
    old_ctx = rope_scaling.get("original_max_position_embeddings") or cfg.get("original_max_position_embeddings", 8192)
    out = []
    for f in freqs:
        wavelen = 2 * math.pi / f
        if wavelen < old_ctx / hi_f:
            out.append(1.0)
        elif wavelen > old_ctx / lo_f:
            out.append(float(factor))
        else:
            smooth = (old_ctx / wavelen - lo_f) / (hi_f - lo_f)
            out.append(float(1 / ((1 - smooth) / factor + smooth)))
    return torch.tensor(out, dtype=torch.float32)


def _rope_scaling(cfg: dict) -> dict:
    rs = cfg.get("rope_scaling") or {}
    if not rs:      # transformers >= 5 keeps it under rope_parameters
        rp = cfg.get("rope_parameters") or {}
        if rp.get("rope_type", "default") not in ("default", None):
            rs = rp
    return rs


def _add_vocab(writer: gguf.GGUFWriter, hf_dir: Optional[str], vocab_size: int) -> str:
    """Tokenizer metadata.  With a real HF directory the upstream vocab loaders are used; without one (random-init
    benchmark models, no network) a placeholder vocabulary of the right size keeps the file loadable."""
    if hf_dir is not None:
        for cls in (gguf.LlamaHfVocab, gguf.SentencePieceVocab, gguf.BpeVocab):
            try:
                from pathlib import Path
                vocab = cls(Path(hf_dir))
                toks, scores, types = [], [], []
                for text, score, ttype in vocab.all_tokens():
                    toks.append(text); scores.append(score); types.append(int(ttype))
                while len(toks) < vocab_size:
                    toks.append(f"[PAD{len(toks)}]".encode()); scores.append(-1000.0); types.append(int(gguf.TokenType.UNUSED))
                writer.add_tokenizer_model("gpt2" if cls is gguf.BpeVocab else "llama")
                if cls is gguf.BpeVocab:
                    writer.add_tokenizer_pre(bpe_pre_tokenizer(hf_dir))
                writer.add_token_list(toks)
                if cls is not gguf.BpeVocab:
                    writer.add_token_scores(scores)
                writer.add_token_types(types)
                gguf.SpecialVocab(Path(hf_dir), load_merges=cls is gguf.BpeVocab, n_vocab=len(toks)).add_to_gguf(writer)
                return cls.__name__
            except NotImplementedError:
                raise
            except Exception:      # noqa: BLE001 -- try the next loader, as the reference's set_vocab does (:2120-2136)
                continue
    writer.add_tokenizer_model("llama")
    writer.add_token_list([f"<tok_{i}>".encode() for i in range(vocab_size)])
    writer.add_token_scores([0.0] * vocab_size)
    writer.add_token_types([int(gguf.TokenType.NORMAL)] * vocab_size)
    return "placeholder"


def write_gguf(model_or_state, config, dir_model_quant: str, outfile: str, outtype: str = "f16",
               hf_dir: Optional[str] = None, name: str = "gptq-gguf-toolkit-b200") -> Dict[str, str]:
    """Write `outfile`.  `config`: a transformers LlamaConfig (or a dict with the same keys).
    Returns {gguf tensor name: ggml type name} for the tensors written."""
    cfg = config if isinstance(config, dict) else config.to_dict()
    arch_ok = cfg.get("model_type", "llama") in ("llama", "mistral")
    if not arch_ok:
        raise NotImplementedError("this writer covers the Llama family; use the reference's vendored converter for others")
    n_layer, n_head = cfg["num_hidden_layers"], cfg["num_attention_heads"]
    n_kv = cfg.get("num_key_value_heads") or n_head
    tdtype, ttype, ftype = OUTTYPES[outtype]
    quant_dirs = {d for d in os.listdir(dir_model_quant) if os.path.isdir(os.path.join(dir_model_quant, d))}

    w = gguf.GGUFWriter(outfile, "llama")
    w.add_name(name)
    w.add_vocab_size(cfg["vocab_size"])
    w.add_context_length(cfg.get("max_position_embeddings", 2048))
    w.add_embedding_length(cfg["hidden_size"])
    w.add_block_count(n_layer)
    w.add_feed_forward_length(cfg["intermediate_size"])
    w.add_head_count(n_head)
    w.add_head_count_kv(n_kv)
    w.add_rope_dimension_count(cfg.get("head_dim") or cfg["hidden_size"] // n_head)
    rope_theta = cfg.get("rope_theta") or (cfg.get("rope_parameters") or {}).get("rope_theta")
    if rope_theta is not None:
        w.add_rope_freq_base(float(rope_theta))
    w.add_layer_norm_rms_eps(float(cfg.get("rms_norm_eps", 1e-5)))
    # RoPE scaling (reference LlamaModel.set_gguf_parameters :2172-2175, 2306-2310; generate_extra_tensors :2259-2287)
    rs = _rope_scaling(cfg)
    rs_type = str(rs.get("rope_type", rs.get("type", ""))).lower() if rs else ""
    rope_freqs = None
    if rs_type == "linear" and "factor" in rs:
        w.add_rope_scaling_type(gguf.RopeScalingType.LINEAR)
        w.add_rope_scaling_factor(rs["factor"])
    elif rs_type == "yarn" and "factor" in rs:
        w.add_rope_scaling_type(gguf.RopeScalingType.YARN)
        w.add_rope_scaling_factor(rs["factor"])
        w.add_rope_scaling_orig_ctx_len(rs["original_max_position_embeddings"])
    elif rs_type == "llama3":
        rope_freqs = llama3_rope_factors(cfg, rs)
    elif rs_type not in ("", "default"):
        raise NotImplementedError(f"rope_scaling type {rs_type!r}: use the reference's vendored converter")
    vocab_kind = _add_vocab(w, hf_dir, cfg["vocab_size"])
    w.add_description(f"GPTQ K-quant tensors from {os.path.abspath(dir_model_quant)}; vocab: {vocab_kind}")

    tmap = gguf.get_tensor_name_map(gguf.MODEL_ARCH.LLAMA, n_layer)
    written: Dict[str, str] = {}
    n_k = 0
    if rope_freqs is not None:
        rname = gguf.TENSOR_NAMES[gguf.MODEL_TENSOR.ROPE_FREQS] + ".weight"
        w.add_tensor(rname, rope_freqs.numpy(), raw_dtype=gguf.GGMLQuantizationType.F32)
        written[rname] = "F32"
    tied = bool(cfg.get("tie_word_embeddings", False))
    for hf_name, t in _iter_state(model_or_state):
        if tied and hf_name == "lm_head.weight":
            continue
        new_name = tmap.get_name(hf_name, try_suffixes=(".weight", ".bias"))
        if new_name is None:
            raise ValueError(f"cannot map tensor {hf_name!r}")
        permute = hf_name.endswith(("q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias"))
        heads_kv = n_head if "q_proj" in hf_name else n_kv
        base = hf_name.removesuffix(".weight")
        if base in quant_dirs:                                   # pack_gptq_into_gguf.py:305-336
            obj = torch.load(os.path.join(dir_model_quant, base, "data.pth"), map_location="cpu", weights_only=True)
            qtype = K_QUANTS[int(obj["q_type"])]
            data = _packed_bytes(obj)
            if tuple(obj["qweight"].shape) != tuple(t.shape):
                raise ValueError(f"{base}: quantised shape {tuple(obj['qweight'].shape)} != weight shape {tuple(t.shape)}")
            if permute:
                data = llama_permute(data, n_head, heads_kv)
            w.add_tensor(new_name, np.ascontiguousarray(data.numpy()), raw_dtype=qtype)
            written[new_name] = qtype.name
            n_k += 1
            continue
        data_t = t.detach().to("cpu")
        if permute:
            data_t = llama_permute(data_t, n_head, heads_kv)
        if data_t.dim() == 1 or new_name.endswith("_norm.weight"):   # upstream: 1-D tensors and norms stay F32
            w.add_tensor(new_name, data_t.to(torch.float32).numpy(), raw_dtype=gguf.GGMLQuantizationType.F32)
            written[new_name] = "F32"
        else:
            w.add_tensor(new_name, np.ascontiguousarray(data_t.to(tdtype).numpy()), raw_dtype=ttype)
            written[new_name] = ttype.name
    w.add_file_type(int(ftype))
    w.add_quantization_version(gguf.GGML_QUANT_VERSION)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    if n_k == 0:
        raise ValueError(f"no tensor of the model has a data.pth under {dir_model_quant}")
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description="Pack GPTQ K-quant results (data.pth per module) into a .gguf")
    ap.add_argument("model", type=str, help="HF model directory (config.json + safetensors [+ tokenizer])")
    ap.add_argument("--dir_model_quant", type=str, required=True, help="save_dir of quant.py (one directory per module)")
    ap.add_argument("--outfile", type=str, required=True)
    ap.add_argument("--outtype", type=str, default="f16", choices=list(OUTTYPES))
    args = ap.parse_args(argv)
    from transformers import AutoConfig, AutoModelForCausalLM
    cfg = AutoConfig.from_pretrained(args.model)
    model = AutoModelForCausalLM.from_pretrained(args.model, torch_dtype="auto")
    written = write_gguf(model, cfg, args.dir_model_quant, args.outfile, args.outtype, hf_dir=args.model,
                         name=os.path.basename(os.path.normpath(args.model)))
    print(json.dumps({"outfile": args.outfile, "tensors": len(written),
                      "k_quant_tensors": sum(v.startswith("Q") for v in written.values())}))


if __name__ == "__main__":
    main()
