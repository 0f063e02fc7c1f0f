#!/usr/bin/env python
"""Golden for BASELINE.json configs[0]: the UNMODIFIED reference driver (quant/gptq/src/quantizer.py::Quantizer, CPU)
on a random-init 2-layer Llama (d_model 256), 8 calibration sequences of 128 tokens, uniform Q4_K.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_driver.py

Writes tests/golden/driver_tiny.npz: for every quantised module the reference's data.pth tensors
(quantizer.py:267-275); the module's final weight is their dequantisation (quantizer.py:257-264).  tests/test_driver_golden_cpu.py builds the same model
and tokens and runs this repo's driver on them.  The comparison is boundary B3: embed_tokens / lm_head
(RTN, no Hessian) must be bit-exact, the GPTQ layers agree statistically -- the reference does not reproduce itself bit
for bit across BLAS/LAPACK thread counts at this boundary.
"""
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference/quant/gptq"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from src.quantizer import Quantizer as RefQuantizer  # noqa: E402
from src import quant_utils as qu  # noqa: E402

REGEX = r".*layers.*((q|k|v|o|gate|up|down)_proj)$"
CFG = dict(vocab_size=1024, hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=4,
           num_key_value_heads=2, max_position_embeddings=256, tie_word_embeddings=False)
N_SEQ, SEQ_LEN = 8, 128


def tiny_model():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    return LlamaForCausalLM(LlamaConfig(**CFG)).float().eval()


def tokens():
    g = torch.Generator().manual_seed(1)
    return [torch.randint(0, CFG["vocab_size"], (1, SEQ_LEN), generator=g) for _ in range(N_SEQ)]


def main():
    torch.set_num_threads(8)
    # torch's CPU sqrt is not correctly rounded (make_golden.py); the goldens of this repo are taken with an IEEE sqrt
    torch.sqrt = lambda t: t.double().sqrt().float()
    model = tiny_model()
    loader = [([], {"input_ids": t}) for t in tokens()]
    T = qu.GGMLQuantizationType
    quant_config = {k: T.Q4_K for k in ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj",
                                        "embed_tokens", "lm_head")}
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        q = RefQuantizer(model, data_loader=loader, quantizable_modules=REGEX,
                         quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                               static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                         pre_block_modules=["model.embed_tokens"], post_block_modules=["lm_head"], block_modules="model.layers",
                         save_dir=tmp, quant_non_block_modules=True, device=torch.device("cpu"))
        q.quantize(quant_config)
        names = sorted(os.listdir(tmp))
        for n in names:
            d = torch.load(os.path.join(tmp, n, "data.pth"))
            for k, v in d.items():
                if isinstance(v, torch.Tensor):
                    a = v.numpy()
                    out[f"{n}|{k}"] = a.view(np.uint16) if a.dtype == np.float16 else a
                else:
                    out[f"{n}|{k}"] = np.asarray(int(v))
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "driver_tiny.npz"), **out)
    print(f"wrote driver_tiny.npz: {len(names)} modules")


if __name__ == "__main__":
    main()
