#!/usr/bin/env python
"""Golden for the non-block RTN path on a BF16 weight (what the reference does to embed_tokens / lm_head of a bf16 model):
Quantizer._quant_non_block_module (quant/gptq/src/quantizer.py:278-330) runs get_scale_and_zero on the weight in its original
dtype, i.e. the whole scale search in bf16 arithmetic.  Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_rtn_bf16.py
Writes tests/golden/rtn_bf16.npz: W as bf16 bit patterns + the reference's five tensors for Q2_K..Q6_K."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/quant/gptq")
HERE = os.path.dirname(os.path.abspath(__file__))
from src import quant_utils as qu  # noqa: E402
from src.quantizer import Quantizer as RefQuantizer  # noqa: E402

T = qu.GGMLQuantizationType


class _Self:
    quantizer_kwargs = {}


def main():
    torch.set_num_threads(8)
    torch.manual_seed(7)
    W = (torch.randn(96, 1024) * 0.02 * torch.exp(0.7 * torch.randn(96, 1))).to(torch.bfloat16)
    W[3, 256:512] = 0.0          # an all-zero super-block
    W[5, :] = 0.0117             # constant rows
    W[7, 512:544] = -0.25        # a constant negative group
    out = {"W_bf16_bits": W.view(torch.int16).numpy().view(np.uint16)}
    for t in (T.Q2_K, T.Q3_K, T.Q4_K, T.Q5_K, T.Q6_K):
        five = RefQuantizer._quant_non_block_module(_Self(), W.clone(), t)
        for k, v in zip(("qweight", "d", "sq", "dmin", "zq"), five):
            a = v.numpy()
            out[f"{t.name}_{k}"] = a.view(np.uint16) if a.dtype == np.float16 else a
    np.savez_compressed(os.path.join(HERE, "rtn_bf16.npz"), **out)
    print("wrote rtn_bf16.npz")
    # the same weight values rounded to fp16: the reference then searches in fp16 arithmetic (clamp_min(1e-9) is a no-op there)
    W16 = W.float().to(torch.float16)
    out = {"W_f16_bits": W16.view(torch.int16).numpy().view(np.uint16)}
    for t in (T.Q2_K, T.Q3_K, T.Q4_K, T.Q5_K, T.Q6_K):
        five = RefQuantizer._quant_non_block_module(_Self(), W16.clone(), t)
        for k, v in zip(("qweight", "d", "sq", "dmin", "zq"), five):
            a = v.numpy()
            out[f"{t.name}_{k}"] = a.view(np.uint16) if a.dtype == np.float16 else a
    np.savez_compressed(os.path.join(HERE, "rtn_f16.npz"), **out)
    print("wrote rtn_f16.npz")


if __name__ == "__main__":
    main()
