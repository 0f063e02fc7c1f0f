"""Fused element-wise pieces of the calibration block forwards (csrc/fwd_ops.cu, include/gq_fwd.h).

The block forwards are the caller's model code (HF transformers), not the reference's hot path -- but they are 60 % of
the wall-clock once the hot path is fast, and in HF's eager Llama a third of that is memory-bound element-wise kernels.
`fused_forward(model)` is a context manager that, for the duration of a quantisation run, swaps three of them for
single-pass CUDA kernels with HF's own rounding points:

    LlamaRMSNorm.forward                         -> gq_fwd_rmsnorm   (6 launches -> 1)
    LlamaMLP.forward: act_fn(gate) * up          -> gq_fwd_silu_mul  (2 -> 1, SiLU only)
    modeling_*.apply_rotary_pos_emb(q, k, ...)   -> gq_fwd_rope      (10 -> 2)

Every replacement is CHECKED against the original on random inputs when it is installed (SiLU*up and the rotary
embedding must be bit-identical, RMSNorm within one 16-bit ulp -- its fp32 mean is summed in a different order); a
replacement that fails its check, a non-16-bit dtype, or an unknown module class simply keeps HF's implementation.
Everything is restored on exit.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import sys

import torch
import torch.nn as nn

from . import _lib as L

_SIGS = {
    "gq_fwd_rmsnorm": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_int, C.c_void_p],
    "gq_fwd_silu_mul": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p],
    "gq_fwd_rope": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long,
                    C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, C.c_void_p],
}
_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = L.load()
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = lib
    return _lib


def _ok16(*ts) -> bool:
    return all(t.is_cuda and t.dtype in (torch.bfloat16, torch.float16) for t in ts) and len({t.dtype for t in ts}) == 1


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xc = x.contiguous()
    out = torch.empty_like(xc)
    dim = xc.shape[-1]
    L.check(_load().gq_fwd_rmsnorm(L.ptr(xc), L.ptr(weight), L.ptr(out), xc.numel() // dim, dim, float(eps), L.dtype_code(xc.dtype),
                                   L.stream_of(xc.device)))
    return out


def silu_mul(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(gate)
    L.check(_load().gq_fwd_silu_mul(L.ptr(gate), L.ptr(up), L.ptr(out), gate.numel(), L.dtype_code(gate.dtype), L.stream_of(gate.device)))
    return out


def rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: (B, H, L, hd) with stride(-1) == 1 (any other strides); cos / sin: (B or 1, L, hd) contiguous."""
    B, H, Lq, hd = x.shape
    out = torch.empty_like(x)           # preserves the (dense) strides of x, like the element-wise ops it replaces
    L.check(_load().gq_fwd_rope(L.ptr(x), L.ptr(out), L.ptr(cos), L.ptr(sin), B, H, Lq, hd, x.stride(0), x.stride(1), x.stride(2),
                                out.stride(0), out.stride(1), out.stride(2), 1 if cos.shape[0] == B and B > 1 else 0,
                                L.dtype_code(x.dtype), L.stream_of(x.device)))
    return out


# --------------------------------------------------------------------------------------------------------------------
def _norm_forward(orig):
    def forward(self, hidden_states):
        w = self.weight
        if (_ok16(hidden_states, w) and hidden_states.shape[-1] % 8 == 0 and hidden_states.shape[-1] <= 8192
                and w.is_contiguous() and hidden_states.numel() > 0):
            return rmsnorm(hidden_states, w, self.variance_epsilon)
        return orig(hidden_states)
    return forward


def _mlp_forward(orig):
    def forward(self, x):
        g, u = self.gate_proj(x), self.up_proj(x)
        if _ok16(g, u) and g.is_contiguous() and u.is_contiguous() and g.numel() % 8 == 0:
            return self.down_proj(silu_mul(g, u))
        return self.down_proj(self.act_fn(g) * u)
    return forward


def _rope_fn(orig):
    def apply_rotary_pos_emb(q, k, cos, sin, unsqueeze_dim=1):
        ok = (unsqueeze_dim == 1 and q.dim() == 4 and k.dim() == 4 and cos.dim() == 3 and _ok16(q, k, cos, sin)
              and q.stride(-1) == 1 and k.stride(-1) == 1 and q.shape[-1] % 16 == 0 and cos.is_contiguous() and sin.is_contiguous()
              and cos.shape[-1] == q.shape[-1] and cos.shape[0] in (1, q.shape[0])
              and all(s % 8 == 0 for t in (q, k) for s in t.stride()[:3]))
        if ok:
            return rope(q, cos, sin), rope(k, cos, sin)
        return orig(q, k, cos, sin, unsqueeze_dim)
    return apply_rotary_pos_emb


def _check_norm(mod) -> bool:
    dim = mod.weight.shape[0]
    x = torch.randn(37, dim, device=mod.weight.device, dtype=mod.weight.dtype) * 1.7
    want = type(mod).forward(mod, x)
    got = rmsnorm(x, mod.weight, mod.variance_epsilon)
    ulp = 2.0 ** (-7 if x.dtype == torch.bfloat16 else -10)
    return bool(((got.float() - want.float()).abs() <= ulp * want.float().abs() + 1e-30).all())


def _check_mlp(mod) -> bool:
    dev, dt = mod.gate_proj.weight.device, mod.gate_proj.weight.dtype
    g = torch.randn(16, 1024, device=dev, dtype=dt) * 3
    u = torch.randn(16, 1024, device=dev, dtype=dt)
    return bool(torch.equal(silu_mul(g, u), mod.act_fn(g) * u))


def _check_rope(orig, dev, dt) -> bool:
    B, H, Hk, Lq, hd = 2, 4, 2, 24, 64
    q = torch.randn(B, Lq, H, hd, device=dev, dtype=dt).transpose(1, 2)
    k = torch.randn(B, Lq, Hk, hd, device=dev, dtype=dt).transpose(1, 2)
    ang = torch.rand(1, Lq, hd, device=dev) * 6.28
    cos, sin = ang.cos().to(dt), ang.sin().to(dt)
    wq, wk = orig(q, k, cos, sin)
    gq, gk = rope(q, cos, sin), rope(k, cos, sin)
    return bool(torch.equal(gq, wq) and torch.equal(gk, wk) and gq.stride() == wq.stride())


@contextlib.contextmanager
def fused_forward(model: nn.Module, verbose: bool = False):
    """Install the fused pieces on `model` (in place, reversibly).  Yields the list of what was installed."""
    installed, undo = [], []
    try:
        p = next(model.parameters(), None)
        if p is None or not p.is_cuda or p.dtype not in (torch.bfloat16, torch.float16):
            yield installed
            return
        norm_ok, mlp_ok = {}, {}
        rope_modules = set()
        for mod in model.modules():
            cls = type(mod)
            name = cls.__name__
            if name.endswith("RMSNorm") and hasattr(mod, "variance_epsilon") and isinstance(getattr(mod, "weight", None), torch.Tensor) \
                    and "forward" not in mod.__dict__:
                if cls not in norm_ok:
                    src_ok = "rsqrt" in (getattr(cls.forward, "__code__", None).co_names if hasattr(cls.forward, "__code__") else ())
                    norm_ok[cls] = bool(src_ok) and mod.weight.dtype == p.dtype and _check_norm(mod)
                if norm_ok[cls] and mod.weight.dtype == p.dtype:
                    orig = mod.forward
                    mod.forward = _norm_forward(orig).__get__(mod, cls)
                    undo.append(lambda m=mod: m.__dict__.pop("forward", None))
                    installed.append(f"{name}.forward")
            elif name.endswith("MLP") and all(hasattr(mod, a) for a in ("gate_proj", "up_proj", "down_proj", "act_fn")) \
                    and "forward" not in mod.__dict__:
                if cls not in mlp_ok:
                    act = type(mod.act_fn).__name__.lower()
                    mlp_ok[cls] = ("silu" in act) and isinstance(mod.gate_proj, nn.Linear) and _check_mlp(mod)
                if mlp_ok[cls]:
                    mod.forward = _mlp_forward(mod.forward).__get__(mod, cls)
                    undo.append(lambda m=mod: m.__dict__.pop("forward", None))
                    installed.append(f"{name}.forward")
            elif name.endswith("Attention") and hasattr(sys.modules.get(cls.__module__), "apply_rotary_pos_emb"):
                rope_modules.add(cls.__module__)
        for mname in rope_modules:
            pymod = sys.modules[mname]
            orig = pymod.apply_rotary_pos_emb
            if _check_rope(orig, p.device, p.dtype):
                pymod.apply_rotary_pos_emb = _rope_fn(orig)
                undo.append(lambda m=pymod, o=orig: setattr(m, "apply_rotary_pos_emb", o))
                installed.append(f"{mname}.apply_rotary_pos_emb")
        if verbose:
            print(f"fused_forward: {len(installed)} replacements installed", flush=True)
        yield installed
    finally:
        for fn in reversed(undo):
            fn()
