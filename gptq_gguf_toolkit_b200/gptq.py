"""Per-layer GPTQ handle -- host-side mirror of the reference's quant/gptq/src/gptq.py::GPTQ.

Same protocol as the reference (created in quantizer.py:239-240, fed from forward hooks :227-236, consumed
in _quant_group :256-265):

    h = GPTQ(layer, rel_damp=..., block_size=128, ...)
    h.update(x)                      # once per calibration batch, any float dtype, on the layer's device
    qweight, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant = h.quantize(q_type)
    h.reset()

All arithmetic is libgq's CUDA (include/gq.h); this class only owns buffers and ordering.  Additions over
the reference: `h.packed` (GGUF block bytes, fused bit-pack) and `h.wdeq` (dequantised weight in the layer's
dtype) are available after quantize(); several handles whose layers see the same input (q/k/v, gate/up) may
share one HessianAccumulator so that H and U are computed once.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.nn.modules.conv import _ConvNd

from . import _lib as L
from . import ops
from .quant_utils import GGML_QUANT_SIZES, GGMLQuantizationType, QuantizationScale


class HessianAccumulator:
    """Running H = 2/N * sum X^T X over calibration batches (gptq.py:80-114), shareable between handles."""

    def __init__(self, d_col: int):
        self.d_col = d_col
        self.H: Optional[torch.Tensor] = None
        self.num_samples = 0
        self.U: Optional[torch.Tensor] = None          # filled by prepare(); shared by the handles
        self.perm: Optional[torch.Tensor] = None       # act_order: argsort(diag H, descending) (gptq.py:210)
        self.U_perm: Optional[torch.Tensor] = None     # ... and the factor of H[perm][:, perm]
        self.not_pd: Optional[torch.Tensor] = None
        self.dead: Optional[torch.Tensor] = None       # diag(H) == 0 before the dead-channel fix
        self.zero_cols: Optional[torch.Tensor] = None  # all-zero weight columns U was built for
        self.synced = False
        self.users = 0                                 # handles attached (GPTQ.reset frees on the last one)

    @torch.no_grad()
    def update(self, input: torch.Tensor) -> None:
        batch_size = input.shape[0]                                         # gptq.py:88 (counts sequences)
        if self.H is None:
            self.H = torch.zeros((self.d_col, self.d_col), device=input.device, dtype=torch.float32)
        x = input.reshape(-1, input.shape[-1])                              # gptq.py:96
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        beta = self.num_samples / (self.num_samples + batch_size)           # gptq.py:110
        alpha = 2.0 / (self.num_samples + batch_size)                       # gptq.py:111
        ops.hessian_update(self.H, x, beta, alpha)                          # gptq.py:112
        self.num_samples += batch_size

    def all_reduce(self) -> None:
        """gptq.py:131-132 (once per accumulator)."""
        if not self.synced and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.H, op=dist.ReduceOp.AVG)
        self.synced = True

    def reset(self) -> None:
        self.H, self.U, self.not_pd, self.dead, self.zero_cols = None, None, None, None, None
        self.perm, self.U_perm = None, None
        self.num_samples = 0
        self.synced = False


class GPTQ:
    def __init__(self, layer: nn.Module, rel_damp: float = 1e-2, block_size: Optional[int] = None,
                 act_order: bool = False, quant_scale: str = "absmax", rmin: float = -1.0, rdelta: float = 0.1,
                 nstep: int = 20, grid: int = 100, static_groups: bool = False, verbose: bool = False,
                 mode: str = "exact", hessian: Optional[HessianAccumulator] = None):
        if act_order:
            assert static_groups                                            # gptq.py:45-46
        assert isinstance(layer, (nn.Linear, _ConvNd)), "OBC supports only linear and convolutional layers."   # gptq.py:77
        self.layer = layer
        self.W = self.layer.weight
        self.d_row, self.d_col = layer.weight.shape[0], int(layer.weight[0].numel())      # model_utils.py:56-57
        self.rel_damp = rel_damp
        self.block_size = block_size or self.d_col                          # gptq.py:55
        self.act_order = act_order
        self.quant_scale = QuantizationScale(quant_scale)   # "mse" only differs for Q3_K/Q6_K and is a no-op there
        self.static_groups = static_groups
        self.grid = grid
        self.rmin, self.rdelta, self.nstep = rmin, rdelta, nstep
        self.mode = {"exact": L.GQ_MODE_EXACT, "fast": L.GQ_MODE_FAST, "exact_left": L.GQ_MODE_EXACT_LEFT,
                     "exact_right": L.GQ_MODE_EXACT_RIGHT}[mode]
        self.W_device, self.W_dtype, self.W_shape = self.W.device, self.W.dtype, self.W.shape
        self.hessian = hessian if hessian is not None else HessianAccumulator(self.d_col)
        self.hessian.users += 1
        self.verbose = verbose
        self.issue_non_invertible = False
        self.packed = None
        self.wdeq = None
        self.not_pd = None

    # -- reference attribute surface ---------------------------------------------------------
    @property
    def H(self):
        return self.hessian.H

    @property
    def num_samples(self):
        return self.hessian.num_samples

    @torch.no_grad()
    def update(self, input: torch.Tensor) -> None:
        """gptq.py:80-114.  Convolutions: the input is unfolded into patches (rows = patches, columns = in_channels x kernel),
        the batch still counts images (gptq.py:88, 97-107)."""
        if isinstance(self.layer, _ConvNd):
            unfold = nn.Unfold(self.layer.kernel_size, dilation=self.layer.dilation, padding=self.layer.padding,
                               stride=self.layer.stride)
            patches = unfold(input).transpose(1, 2)            # (batch, num_patches, channels * prod(kernel_size))
            self.hessian.update(patches)
            return
        self.hessian.update(input)

    def reset(self) -> None:
        """gptq.py:116-120."""
        self.W = self.layer.weight
        self.hessian.users -= 1
        if self.hessian.users <= 0:     # the last handle sharing the accumulator frees H and U
            self.hessian.reset()
            self.hessian.users = 0
        self.packed, self.wdeq = None, None

    @torch.no_grad()
    def quantization_pre_step(self) -> None:
        """gptq.py:123-143."""
        assert self.H is not None, "One has to process at least one sample of calibration data to run pruning"
        self.hessian.all_reduce()
        self.W = self.W.clone().float().flatten(1, -1).contiguous()      # gptq.py:138-140 (convolutions: (out, in x kernel))
        if self.hessian.dead is None:
            self.hessian.dead = torch.diagonal(self.H) == 0     # bookkeeping for handles sharing this H
            ops.pre_step(self.H, self.W)
        else:   # a sharer already replaced the zero diagonal entries by 1 (gptq.py:135); only zero our columns
            self.W.masked_fill_(self.hessian.dead[None, :], 0.0)
        self.pre_step_completed = True

    @torch.no_grad()
    def _prepare(self) -> torch.Tensor:
        """gptq.py:305-324; returns U = H_inv_cho, ROW-major upper triangular (the reference's is column-major)."""
        if self.hessian.U is None:
            self.hessian.zero_cols = (self.W == 0).all(dim=0)
            self.hessian.U, self.hessian.not_pd = ops.prepare(self.H, self.W, self.rel_damp)
        elif not torch.equal((self.W == 0).all(dim=0), self.hessian.zero_cols):
            # U depends on which weight columns are entirely zero (gptq.py:308-313)
            raise RuntimeError("layers sharing a HessianAccumulator have different all-zero weight columns; "
                               "give them separate accumulators")
        self.not_pd = self.hessian.not_pd
        return self.hessian.U

    @torch.no_grad()
    def _prepare_act_order(self):
        """gptq.py:209-216: perm = argsort(diag H, descending); the loop runs on W[:, perm] with the factor of
        H[perm][:, perm].  Shared by the handles attached to the same accumulator (same H => same perm)."""
        hs = self.hessian
        if hs.perm is None:
            hs.perm = torch.argsort(torch.diag(hs.H), descending=True)
            Hp = hs.H.index_select(0, hs.perm).index_select(1, hs.perm).contiguous()
            hs.U_perm, hs.not_pd = ops.prepare(Hp, self.W.index_select(1, hs.perm).contiguous(), self.rel_damp)
        self.not_pd = hs.not_pd
        return hs.perm, hs.U_perm

    @torch.no_grad()
    def step(self, q_type: GGMLQuantizationType):
        """gptq.py:146-295."""
        q_type = GGMLQuantizationType(int(q_type))
        if q_type == GGMLQuantizationType.Q3_K:             # gptq.py:204-206 (the reference switches them off for good)
            self.act_order = False
            self.static_groups = False
        perm = None
        if self.act_order:
            perm, U = self._prepare_act_order()
        else:
            U = self._prepare()
        out = ops.gptq_quantize(self.W, U, int(q_type), self.block_size, self.rmin, self.rdelta, self.nstep,
                                self.mode, packed=True, wdeq_dtype=self.W_dtype, static_groups=self.static_groups, perm=perm)
        qweight, d, sq, dmin, zq, self.packed, self.wdeq, _ = out
        self.W = None   # the working copy was consumed by the kernel
        return qweight, d, sq, dmin, zq

    def quantize(self, q_type: GGMLQuantizationType):
        """gptq.py:297-302 -> (qweight, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)."""
        self.quantization_pre_step()
        return self.step(q_type)

    def non_invertible(self) -> bool:
        """True if the Hessian was not positive definite and U fell back to the identity (gptq.py:321-323).
        Reads a device flag: synchronises."""
        self.issue_non_invertible = bool(self.not_pd.item()) if self.not_pd is not None else False
        return self.issue_non_invertible


__all__ = ["GPTQ", "HessianAccumulator", "GGML_QUANT_SIZES", "GGMLQuantizationType"]
