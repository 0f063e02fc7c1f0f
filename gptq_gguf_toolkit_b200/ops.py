"""Torch-tensor front-end of the C ABI (include/gq.h): allocation and argument marshalling only.

Every function enqueues on torch's current stream of the tensors' device and returns immediately; the
five K-quant tensors are always returned in the reference's order
    (qweight, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)
(quant/gptq/src/gptq.py:295 of the reference).
"""
from __future__ import annotations

import contextlib
from typing import Optional, Tuple

import torch

from . import _lib as L

QK_K = 256
_ws_cache: dict = {}
_timer = None   # optional quantizer.PhaseTimer: per-op CUDA-event spans ("hessian", "prepare", "gptq", "rtn")


def set_timer(timer) -> None:
    global _timer
    _timer = timer


class _NullSpan:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _span(name: str):
    return _timer.span(name) if _timer is not None else _NullSpan()


def launch_count() -> int:
    """Kernels launched by libgq in this process so far."""
    return int(L.load().gq_launch_count())


def profile_enable(on: bool) -> None:
    """Record CUDA events around the column-loop kernel / rank-k GEMM launches (see gq_profile_enable)."""
    L.load().gq_profile_enable(1 if on else 0)


def profile_read() -> dict:
    """{'panel_ms', 'panel_launches', 'rankk_gemm_ms', 'rankk_gemm_launches', 'split_ms', 'split_launches'} since the last
    read; synchronises.  rankk_* = the rank-k update launches (tcgen05 GEMM in fast mode, exact_update_kernel otherwise),
    split_* = the operand preparation of fast mode's GEMMs (fp16 hi/lo split of the errors, once per layer of U^T)."""
    import ctypes as C
    ms = (C.c_float * 3)()
    n = (C.c_int * 3)()
    L.check(L.load().gq_profile_read3(ms, n))
    return {"panel_ms": ms[0], "panel_launches": n[0], "rankk_gemm_ms": ms[1], "rankk_gemm_launches": n[1],
            "split_ms": ms[2], "split_launches": n[2]}


def _workspace(device, nbytes: int, slot: int = 0) -> torch.Tensor:
    """One growing scratch buffer per (device, slot); callers that run ops concurrently on several streams
    give every stream its own slot (torch caching allocator owns the memory)."""
    key = (device.type, device.index, slot)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            # the old buffer may still be in use by work enqueued on a side stream: keep it alive until the
            # device is idle instead of handing it back to the allocator of the current stream
            torch.cuda.synchronize(device)
        _ws_cache[key] = ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
    return ws


def release_workspaces():
    _ws_cache.clear()


def _code_dtype(fmt) -> torch.dtype:
    return torch.uint8 if fmt["asym"] else torch.int8


def alloc_outputs(q_type: int, d_row: int, d_col: int, device, packed: bool, wdeq_dtype: Optional[torch.dtype]):
    f = L.format_info(q_type)
    cd = _code_dtype(f)
    nsb, ng = d_col // QK_K, d_col // f["group_size"]
    qweight = torch.empty(d_row, d_col, dtype=cd, device=device)
    d = torch.empty(d_row, nsb, dtype=torch.float16, device=device)
    dmin = torch.empty(d_row, nsb, dtype=torch.float16, device=device)
    sq = torch.empty(d_row, ng, dtype=cd, device=device)
    zq = torch.empty(d_row, ng, dtype=cd, device=device)
    pk = torch.empty(d_row, nsb * f["type_size"], dtype=torch.uint8, device=device) if packed else None
    wd = torch.empty(d_row, d_col, dtype=wdeq_dtype, device=device) if wdeq_dtype is not None else None
    return qweight, d, sq, dmin, zq, pk, wd


def hessian_update(H: torch.Tensor, X: torch.Tensor, beta: float, alpha: float) -> None:
    """H <- beta*H + alpha*X^T X  (gptq.py:110-112).  X: (n_tok, d_col) contiguous, fp32/fp16/bf16."""
    L.require_cuda(H, X)
    assert H.dtype == torch.float32 and H.is_contiguous() and X.is_contiguous() and X.dim() == 2
    n_tok, d_col = X.shape
    lib = L.load()
    code = L.dtype_code(X.dtype)
    nws = lib.gq_hessian_workspace_bytes(n_tok, d_col, code)
    ws = _workspace(H.device, nws) if nws else None
    with _span("hessian"):
        L.check(lib.gq_hessian_update(L.ptr(H), L.ptr(X), n_tok, d_col, code, float(beta), float(alpha),
                                      L.ptr(ws), nws, L.stream_of(H.device)))


def pre_step(H: torch.Tensor, W: torch.Tensor) -> None:
    """Dead-channel fix (gptq.py:134-141), in place on H and the fp32 working copy W."""
    L.require_cuda(H, W)
    assert W.dtype == torch.float32 and W.is_contiguous() and H.is_contiguous()
    L.check(L.load().gq_pre_step(L.ptr(H), L.ptr(W), W.shape[0], W.shape[1], L.stream_of(H.device)))


def prepare(H: torch.Tensor, W: torch.Tensor, rel_damp: float, stream: Optional["torch.cuda.Stream"] = None,
            slot: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """U = chol(inv(H + damp I), upper), row-major (gptq.py:305-324).  H is masked + damped in place.
    Returns (U, not_pd) where not_pd is a device int32 scalar (1 => U is the identity).
    stream: enqueue on this side stream instead of the current one (outputs are allocated on the CURRENT stream,
    the side stream first waits for it; the caller waits on the side stream before consuming U).  slot: workspace slot."""
    L.require_cuda(H, W)
    assert H.dtype == torch.float32 and H.is_contiguous() and W.dtype == torch.float32 and W.is_contiguous()
    d_row, d_col = W.shape
    lib = L.load()
    U = torch.empty(d_col, d_col, dtype=torch.float32, device=H.device)
    flag = torch.zeros(1, dtype=torch.int32, device=H.device)
    nws = lib.gq_prepare_workspace_bytes(d_col)
    ws = _workspace(H.device, nws, slot)
    if stream is not None:
        stream.wait_stream(torch.cuda.current_stream(H.device))
    with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        with _span("prepare"):
            L.check(lib.gq_prepare(L.ptr(H), L.ptr(W), d_row, d_col, float(rel_damp), L.ptr(U), L.ptr(ws), nws,
                                   L.ptr(flag), L.stream_of(H.device)))
    return U, flag


def gptq_quantize(W: torch.Tensor, U: torch.Tensor, q_type: int, block_size: int = 128, rmin: float = -1.0,
                  rdelta: float = 0.1, nstep: int = 20, mode: int = L.GQ_MODE_EXACT, packed: bool = True,
                  wdeq_dtype: Optional[torch.dtype] = None, search_flags: bool = False,
                  stream: Optional["torch.cuda.Stream"] = None, static_groups: bool = False,
                  perm: Optional[torch.Tensor] = None, ws_slot: Optional[int] = None):
    """The column loop of one layer (gptq.py:146-295).  W: fp32 working copy in the ORIGINAL column order, CLOBBERED.
    Returns (qweight, d, sq, dmin, zq, packed|None, wdeq|None, flags|None), all in the original column order.
    stream: enqueue on this side stream (outputs are allocated on the CURRENT stream, which the side stream first
    waits for; the caller waits on the side stream before consuming the results).
    ws_slot: scratch-buffer slot of GQ_MODE_FAST (default: 0 on the current stream, 100 on a side stream).
    static_groups (gptq.py:184-196): scales searched up front on W.  perm (act_order, gptq.py:209-216, needs
    static_groups): permutation of the columns, U must belong to H[perm][:, perm].  Q3_K ignores both (:204-206)."""
    L.require_cuda(W, U)
    assert W.dtype == torch.float32 and W.is_contiguous() and U.dtype == torch.float32 and U.is_contiguous()
    d_row, d_col = W.shape
    if int(q_type) == 11:
        static_groups, perm = False, None
    assert perm is None or static_groups, "act_order requires static_groups (gptq.py:45-46)"
    fused = perm is None
    qweight, d, sq, dmin, zq, pk, wd = alloc_outputs(q_type, d_row, d_col, W.device, packed and fused,
                                                     wdeq_dtype if fused else None)
    flags = torch.zeros(d_col // QK_K, 2, dtype=torch.int32, device=W.device) if search_flags else None
    lib = L.load()
    nws = lib.gq_gptq_workspace_bytes(d_row, d_col, int(mode))
    # callers that run several column loops concurrently on different streams give each its own scratch slot
    ws = _workspace(W.device, nws, slot=ws_slot if ws_slot is not None else (100 if stream is not None else 0)) if nws else None
    if stream is not None:
        stream.wait_stream(torch.cuda.current_stream(W.device))
    with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        with _span("gptq"):
            sg, Wk, pm, qk = (1 if static_groups else 0), W, None, qweight
            if perm is not None:
                # scales on the un-permuted weights (= the RTN search, quantizer.py:285-300 does the same loop), then the
                # loop on W[:, perm]; the codes come back in loop order
                scratch = torch.empty_like(qweight)
                L.check(lib.gq_rtn_quantize(L.ptr(W), L.GQ_F32, d_row, d_col, int(q_type), float(rmin), float(rdelta), int(nstep),
                                            L.ptr(scratch), L.ptr(d), L.ptr(sq), L.ptr(dmin), L.ptr(zq), None, None, 0,
                                            L.stream_of(W.device)))
                pm = perm.to(torch.int32).contiguous()
                Wk = W.index_select(1, perm).contiguous()
                sg, qk = 2, scratch
            L.check(lib.gq_gptq_quantize_ex(
                L.ptr(Wk), L.ptr(U), d_row, d_col, int(q_type), int(block_size), float(rmin), float(rdelta), int(nstep), int(mode),
                sg, L.ptr(pm), L.ptr(qk), L.ptr(d), L.ptr(sq), L.ptr(dmin), L.ptr(zq), L.ptr(pk), L.ptr(wd),
                L.dtype_code(wdeq_dtype) if (wdeq_dtype is not None and fused) else 0, L.ptr(flags), L.ptr(ws), nws,
                L.stream_of(W.device)))
            if stream is not None:
                # inputs / temporaries that were allocated on the CURRENT stream and are read by the kernels just enqueued on
                # `stream`: a caller may drop its last reference (a stacked sub-matrix built for this call, the permuted copies)
                # as soon as this function returns, and the caching allocator would hand the memory to the current stream's next
                # allocation underneath the running kernel
                for t in (W, Wk, qk, pm, U):
                    if t is not None:
                        t.record_stream(stream)
            if perm is not None:
                qweight.index_copy_(1, perm, qk)                    # gptq.py:276-277: qweight[:, invperm]
                pk = pack(q_type, qweight, d, sq, dmin, zq) if packed else None
                wd = dequantize(q_type, qweight, d, sq, dmin, zq, wdeq_dtype) if wdeq_dtype is not None else None
    return qweight, d, sq, dmin, zq, pk, wd, flags


def rtn_quantize(W: torch.Tensor, q_type: int, rmin: float = -1.0, rdelta: float = 0.1, nstep: int = 20,
                 packed: bool = True, wdeq_dtype: Optional[torch.dtype] = None, native_arith: bool = False):
    """RTN K-quant of a weight without Hessian (quantizer.py:278-330).  W is read-only (fp32/fp16/bf16).
    Returns (qweight, d, sq, dmin, zq, packed|None, wdeq|None).
    native_arith (gq_rtn_quantize_native): for a bf16 / fp16 weight run the scale search in that dtype's arithmetic like the
    reference does (quantizer.py:303-305) -- what the driver passes for embed_tokens / lm_head; False = widen to fp32."""
    L.require_cuda(W)
    assert W.is_contiguous() and W.dim() == 2
    d_row, d_col = W.shape
    qweight, d, sq, dmin, zq, pk, wd = alloc_outputs(q_type, d_row, d_col, W.device, packed, wdeq_dtype)
    with _span("rtn"):
        fn = L.load().gq_rtn_quantize_native if native_arith else L.load().gq_rtn_quantize
        L.check(fn(
            L.ptr(W), L.dtype_code(W.dtype), d_row, d_col, int(q_type), float(rmin), float(rdelta), int(nstep),
            L.ptr(qweight), L.ptr(d), L.ptr(sq), L.ptr(dmin), L.ptr(zq), L.ptr(pk), L.ptr(wd),
            L.dtype_code(wdeq_dtype) if wdeq_dtype is not None else 0, L.stream_of(W.device)))
    return qweight, d, sq, dmin, zq, pk, wd


def get_scale_and_zero(x: torch.Tensor, q_type: int, rmin: float = -1.0, rdelta: float = 0.1, nstep: int = 20):
    """quant_utils.py:90-145.  x: (rows, 256) fp32 (row stride may exceed 256).
    Returns (super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)."""
    L.require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == QK_K and x.stride(1) == 1
    rows = x.shape[0]
    f = L.format_info(q_type)
    gpr = QK_K // f["group_size"]
    cd = _code_dtype(f)
    d = torch.empty(rows, dtype=torch.float16, device=x.device)
    dmin = torch.empty(rows, dtype=torch.float16, device=x.device)
    sq = torch.empty(rows, gpr, dtype=cd, device=x.device)
    zq = torch.empty(rows, gpr, dtype=cd, device=x.device)
    L.check(L.load().gq_get_scale_and_zero(
        L.ptr(x), x.stride(0), rows, int(q_type), float(rmin), float(rdelta), int(nstep),
        L.ptr(d), L.ptr(dmin), 1, L.ptr(sq), L.ptr(zq), gpr, None, L.stream_of(x.device)))
    return d, sq, dmin, zq


def dequantize(q_type: int, qweight, d, sq, dmin, zq, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """dequantize_linear_weight (quant_utils.py:277-310)."""
    L.require_cuda(qweight, d, sq, dmin, zq)
    d_row, d_col = qweight.shape
    out = torch.empty(d_row, d_col, dtype=out_dtype, device=qweight.device)
    L.check(L.load().gq_dequantize(int(q_type), L.ptr(qweight.contiguous()), L.ptr(d.contiguous()), L.ptr(sq.contiguous()),
                                   L.ptr(dmin.contiguous()), L.ptr(zq.contiguous()), d_row, d_col, L.ptr(out),
                                   L.dtype_code(out_dtype), L.stream_of(qweight.device)))
    return out


def pack(q_type: int, qweight, d, sq, dmin=None, zq=None) -> torch.Tensor:
    """pack_Q*K (packing_utils.py:33-326) -> (d_row, d_col/256*type_size) uint8 on the device."""
    L.require_cuda(qweight, d, sq, dmin, zq)
    d_row, d_col = qweight.shape
    f = L.format_info(q_type)
    out = torch.empty(d_row, d_col // QK_K * f["type_size"], dtype=torch.uint8, device=qweight.device)
    keep = [qweight.contiguous(), d.contiguous(), sq.contiguous(),
            dmin.contiguous() if dmin is not None else None, zq.contiguous() if zq is not None else None]
    L.check(L.load().gq_pack(int(q_type), L.ptr(keep[0]), L.ptr(keep[1]), L.ptr(keep[2]), L.ptr(keep[3]), L.ptr(keep[4]),
                             d_row, d_col, L.ptr(out), L.stream_of(qweight.device)))
    return out
