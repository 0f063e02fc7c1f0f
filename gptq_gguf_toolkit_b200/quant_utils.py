"""K-quant numerics -- host-side mirror of the reference's quant/gptq/src/quant_utils.py.

Same names, argument meaning and return order as the reference; the arithmetic runs in libgq's CUDA
kernels (bit-identical to the reference's CPU results, see tests/).  No CPU fallback.
"""
from __future__ import annotations

from enum import Enum, IntEnum

import torch

from . import ops

QK_K = 256


class GGMLQuantizationType(IntEnum):          # quant_utils.py:11-16
    Q2_K = 10
    Q3_K = 11
    Q4_K = 12
    Q5_K = 13
    Q6_K = 14


# bits, q_clamp_values, max_scale_q, group_size, super_group_size, dtype_scale_zero, dtype_qweight
GGML_QUANT_SIZES = {                          # quant_utils.py:19-26
    GGMLQuantizationType.Q2_K: (2, (0, 2**2 - 1), 2**4 - 1, 16, QK_K, torch.uint8, torch.uint8),
    GGMLQuantizationType.Q3_K: (3, (-4, 3), 2**5 - 1, 16, QK_K, torch.int8, torch.int8),
    GGMLQuantizationType.Q4_K: (4, (0, 2**4 - 1), 2**6 - 1, 32, QK_K, torch.uint8, torch.uint8),
    GGMLQuantizationType.Q5_K: (5, (0, 2**5 - 1), 2**6 - 1, 32, QK_K, torch.uint8, torch.uint8),
    GGMLQuantizationType.Q6_K: (6, (-32, 31), 2**6 - 1, 16, QK_K, torch.int8, torch.int8),
}

# bytes per 256-weight super-block in a .gguf file
GGUF_TYPE_SIZE = {GGMLQuantizationType.Q2_K: 84, GGMLQuantizationType.Q3_K: 110, GGMLQuantizationType.Q4_K: 144,
                  GGMLQuantizationType.Q5_K: 176, GGMLQuantizationType.Q6_K: 210}


class QuantizationScale(str, Enum):           # quant_utils.py:29-31
    ABSMAX = "absmax"
    MSE = "mse"


class Quantizer:
    """Mirror of quant_utils.Quantizer (quant_utils.py:49-145): configure(...) then get_scale_and_zero(x, q_type)."""

    def configure(self, bits, scale_maxq, group_size, group_type, super_group_size,
                  quant_scale=QuantizationScale.ABSMAX, grid=100, maxshrink=0.80, norm=2.0,
                  rmin=-1.0, rdelta=0.1, nstep=20, eps=1e-9):
        self.bits = bits
        self.maxq = 2**bits - 1
        self.scale_maxq = scale_maxq
        self.group_size = group_size
        self.supergroup_size = super_group_size
        self.group_type = group_type
        self.rmin, self.rdelta, self.nstep, self.eps = rmin, rdelta, nstep, eps
        self.quant_scale = QuantizationScale(quant_scale)
        self.grid, self.maxshrink, self.norm = grid, maxshrink, norm
        if eps != 1e-9:
            raise NotImplementedError("libgq fixes eps = 1e-9 (the reference's default)")

    def get_scale_and_zero(self, x: torch.Tensor, q_type: GGMLQuantizationType):
        """x: (rows, 256) fp32 on a CUDA device -> (super_group_scale fp16, group_scale_quant,
        super_group_zero fp16, group_zero_quant), quant_utils.py:90-145."""
        assert x.ndim == 2 and x.shape[1] == QK_K, f"expected (rows, {QK_K})"
        if x.dtype != torch.float32:
            x = x.float()
        if x.stride(1) != 1 or x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0:
            x = x.contiguous()
        return ops.get_scale_and_zero(x, int(q_type), getattr(self, "rmin", -1.0), getattr(self, "rdelta", 0.1),
                                      getattr(self, "nstep", 20))


def dequantize_linear_weight(q_type, qweight, super_group_scale, group_scale_quant, super_group_zero,
                             group_zero_quant) -> torch.Tensor:
    """quant_utils.py:277-310 -> fp32 (d_row, d_col)."""
    return ops.dequantize(int(q_type), qweight, super_group_scale, group_scale_quant, super_group_zero,
                          group_zero_quant, torch.float32)
