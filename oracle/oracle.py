"""ctypes front-end of the CPU oracle (oracle/gq_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package never imports this module.

All functions take / return numpy arrays; the 5-tensor order follows the reference's
GPTQ.quantize return value (quant/gptq/src/gptq.py:295):
    (qweight, super_group_scale, group_scale_quant, super_group_zero, group_zero_quant)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgq_oracle.so")

Q2_K, Q3_K, Q4_K, Q5_K, Q6_K = 10, 11, 12, 13, 14
QK_K = 256

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gq_oracle.c")
    if force or not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(_SO) < os.path.getmtime(src)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libgq_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_abi_version.restype = C.c_int
    return _lib


def fmt(qtype: int) -> dict:
    out = (C.c_int * 7)()
    if lib().orc_format(int(qtype), out):
        raise ValueError(f"unsupported q_type {qtype}")
    k = ["bits", "qmin", "qmax", "scale_maxq", "group_size", "asym", "type_size"]
    return dict(zip(k, list(out)))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _code_dtype(qtype):
    return np.uint8 if fmt(qtype)["asym"] else np.int8


def get_scale_and_zero(x: np.ndarray, qtype: int, rmin=-1.0, rdelta=0.1, nstep=20, return_flags=False):
    """quant_utils.py:90-145.  x: (rows, 256) fp32.  Returns (d fp16, sq, dmin fp16, zq)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, w = x.shape
    assert w == QK_K
    f = fmt(qtype)
    gpr = QK_K // f["group_size"]
    d = np.empty(rows, np.uint16)
    dmin = np.empty(rows, np.uint16)
    sq = np.empty((rows, gpr), np.uint8)
    zq = np.empty((rows, gpr), np.uint8)
    flags = np.zeros(2, np.uint32)
    rc = lib().orc_get_scale_and_zero(
        _p(x, C.c_float), C.c_long(QK_K), C.c_int(rows), C.c_int(qtype),
        C.c_double(rmin), C.c_double(rdelta), C.c_int(nstep),
        _p(d, C.c_uint16), C.c_long(1), _p(dmin, C.c_uint16), C.c_long(1),
        _p(sq, C.c_uint8), C.c_long(gpr), _p(zq, C.c_uint8), C.c_long(gpr), _p(flags, C.c_uint32))
    assert rc == 0
    cd = _code_dtype(qtype)
    out = (d.view(np.float16), sq.view(cd), dmin.view(np.float16), zq.view(cd))
    return out + (flags,) if return_flags else out


def gptq_step(W: np.ndarray, U: np.ndarray, qtype: int, block_size=128, rmin=-1.0, rdelta=0.1, nstep=20,
              return_flags=False, static_groups=False, perm=None):
    """gptq.py:146-295.  W (d_row,d_col) fp32 (not modified), U upper-triangular (any strides).
    static_groups / perm (act_order: perm = argsort(diag H, descending), U from the permuted H): gptq.py:184-216.
    Returns (qweight, d, sq, dmin, zq, w_dequant), all in the original column order."""
    Wc = np.array(W, dtype=np.float32, order="C", copy=True)
    U = np.asarray(U, dtype=np.float32)
    d_row, d_col = Wc.shape
    f = fmt(qtype)
    ng, nsb = d_col // f["group_size"], d_col // QK_K
    qw = np.empty((d_row, d_col), np.uint8)
    d = np.empty((d_row, nsb), np.uint16)
    dmin = np.empty((d_row, nsb), np.uint16)
    sq = np.empty((d_row, ng), np.uint8)
    zq = np.empty((d_row, ng), np.uint8)
    flags = np.zeros((nsb, 2), np.uint32)
    rs, cs = U.strides[0] // 4, U.strides[1] // 4
    if qtype == 11:                      # Q3_K: both options are forced off (gptq.py:204-206)
        static_groups, perm = False, None
    pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    rc = lib().orc_gptq_step_ex(
        _p(Wc, C.c_float), _p(U, C.c_float), C.c_long(rs), C.c_long(cs),
        C.c_int(d_row), C.c_int(d_col), C.c_int(qtype), C.c_int(block_size),
        C.c_double(rmin), C.c_double(rdelta), C.c_int(nstep), C.c_int(1 if static_groups else 0),
        _p(pm, C.c_int) if pm is not None else None,
        _p(qw, C.c_uint8), _p(d, C.c_uint16), _p(dmin, C.c_uint16), _p(sq, C.c_uint8), _p(zq, C.c_uint8),
        _p(flags, C.c_uint32))
    assert rc == 0, rc
    cd = _code_dtype(qtype)
    out = (qw.view(cd), d.view(np.float16), sq.view(cd), dmin.view(np.float16), zq.view(cd), Wc)
    return out + (flags,) if return_flags else out


def rtn_quantize(W: np.ndarray, qtype: int, rmin=-1.0, rdelta=0.1, nstep=20, bf16: bool = False, fp16: bool = False):
    """quantizer.py:278-330.  Returns the 5 tensors.  bf16=True / fp16=True: W holds bf16 / fp16 values (widened to fp32) and the
    scale search runs in that dtype's arithmetic, like the reference does for a 16-bit model's embed_tokens / lm_head."""
    W = np.ascontiguousarray(W, dtype=np.float32)
    d_row, d_col = W.shape
    f = fmt(qtype)
    ng, nsb = d_col // f["group_size"], d_col // QK_K
    qw = np.empty((d_row, d_col), np.uint8)
    d = np.empty((d_row, nsb), np.uint16)
    dmin = np.empty((d_row, nsb), np.uint16)
    sq = np.empty((d_row, ng), np.uint8)
    zq = np.empty((d_row, ng), np.uint8)
    fn = lib().orc_rtn_quantize_bf16 if bf16 else lib().orc_rtn_quantize_fp16 if fp16 else lib().orc_rtn_quantize
    rc = fn(
        _p(W, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_int(qtype),
        C.c_double(rmin), C.c_double(rdelta), C.c_int(nstep),
        _p(qw, C.c_uint8), _p(d, C.c_uint16), _p(dmin, C.c_uint16), _p(sq, C.c_uint8), _p(zq, C.c_uint8))
    assert rc == 0, rc
    cd = _code_dtype(qtype)
    return qw.view(cd), d.view(np.float16), sq.view(cd), dmin.view(np.float16), zq.view(cd)


def _raw5(qweight, d, sq, dmin, zq):
    qw = np.ascontiguousarray(qweight).view(np.uint8)
    dd = np.ascontiguousarray(d, dtype=np.float16).view(np.uint16)
    dm = np.ascontiguousarray(dmin, dtype=np.float16).view(np.uint16)
    s = np.ascontiguousarray(sq).view(np.uint8)
    z = np.ascontiguousarray(zq).view(np.uint8)
    return qw, dd, s, dm, z


def dequantize(qtype, qweight, d, sq, dmin, zq):
    """quant_utils.py:277-310 -> (d_row,d_col) fp32."""
    qw, dd, s, dm, z = _raw5(qweight, d, sq, dmin, zq)
    d_row, d_col = qw.shape
    out = np.empty((d_row, d_col), np.float32)
    rc = lib().orc_dequantize(C.c_int(qtype), _p(qw, C.c_uint8), _p(dd, C.c_uint16), _p(s, C.c_uint8),
                              _p(dm, C.c_uint16), _p(z, C.c_uint8), C.c_int(d_row), C.c_int(d_col),
                              _p(out, C.c_float))
    assert rc == 0
    return out


def pack(qtype, qweight, d, sq, dmin, zq):
    """packing_utils.py:33-326 -> (d_row, d_col/256*type_size) uint8.  Inputs are not modified."""
    qw, dd, s, dm, z = _raw5(qweight, d, sq, dmin, zq)
    d_row, d_col = qw.shape
    ts = fmt(qtype)["type_size"]
    out = np.empty((d_row, d_col // QK_K * ts), np.uint8)
    rc = lib().orc_pack(C.c_int(qtype), _p(qw, C.c_uint8), _p(dd, C.c_uint16), _p(s, C.c_uint8),
                        _p(dm, C.c_uint16), _p(z, C.c_uint8), C.c_int(d_row), C.c_int(d_col),
                        _p(out, C.c_uint8))
    assert rc == 0
    return out


def hessian_update(H: np.ndarray, X: np.ndarray, beta: float, alpha: float):
    """gptq.py:110-112 in place on H (fp32 C-contiguous)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    assert H.dtype == np.float32 and H.flags.c_contiguous
    n_tok, d_col = X.shape
    lib().orc_hessian_update(_p(H, C.c_float), _p(X, C.c_float), C.c_long(n_tok), C.c_int(d_col),
                             C.c_float(beta), C.c_float(alpha))
    return H


def prepare(H: np.ndarray, W: np.ndarray, rel_damp: float = 0.01):
    """gptq.py:123-143,305-324.  Returns (U row-major upper, H masked+damped, W masked, not_pd)."""
    Hc = np.array(H, dtype=np.float32, order="C", copy=True)
    Wc = np.array(W, dtype=np.float32, order="C", copy=True)
    d_row, n = Wc.shape
    U = np.empty((n, n), np.float32)
    fail = lib().orc_prepare(_p(Hc, C.c_float), _p(Wc, C.c_float), C.c_int(d_row), C.c_int(n),
                             C.c_float(rel_damp), _p(U, C.c_float))
    return U, Hc, Wc, bool(fail)
