"""GGUF block byte packers -- host-side mirror of the reference's quant/gptq/src/packing_utils.py.

Same names and signatures as the reference (pack_gptq_into_gguf.py:326-336 calls them positionally) and
the same return type (np.uint8 array of shape (d_row, d_col/256*type_size)); the packing itself runs in
libgq's CUDA kernel.  Unlike the reference, pack_Q3K / pack_Q6K do NOT modify their inputs in place.
CPU tensors (e.g. loaded from data.pth) are moved to the current CUDA device first; there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .quant_utils import GGMLQuantizationType as T


def _dev(t):
    if t is None or t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise RuntimeError("packing_utils: no CUDA device (libgq has no CPU fallback)")
    return t.cuda()


def _pack(q_type, qweights, d, sq, dmin=None, zq=None) -> np.ndarray:
    out = ops.pack(int(q_type), _dev(qweights), _dev(d.to(torch.float16)), _dev(sq),
                   _dev(dmin.to(torch.float16)) if dmin is not None else None, _dev(zq))
    return out.cpu().numpy()


def pack_Q2K(qweights, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant) -> np.ndarray:
    return _pack(T.Q2_K, qweights, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)


def pack_Q3K(qweights, super_group_scale, group_scale_quant) -> np.ndarray:
    return _pack(T.Q3_K, qweights, super_group_scale, group_scale_quant)


def pack_Q4K(qweights, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant) -> np.ndarray:
    return _pack(T.Q4_K, qweights, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)


def pack_Q5K(qweights, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant) -> np.ndarray:
    return _pack(T.Q5_K, qweights, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)


def pack_Q6K(qweights, super_group_scales, group_scale_quant) -> np.ndarray:
    return _pack(T.Q6_K, qweights, super_group_scales, group_scale_quant)


PACKERS = {T.Q2_K: pack_Q2K, T.Q3_K: pack_Q3K, T.Q4_K: pack_Q4K, T.Q5_K: pack_Q5K, T.Q6_K: pack_Q6K}
