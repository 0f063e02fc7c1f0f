"""TEST-ONLY stand-in for gptq_gguf_toolkit_b200.ops backed by the CPU oracle, so that the host-side logic
(handles, Hessian sharing, stacking, row sharding, data.pth schema, CLI) can be exercised without a GPU.
Never imported by the product package."""
import numpy as np
import torch

from oracle import oracle as orc

QK_K = 256


def _np(t):
    return t.detach().cpu().numpy()


def _five_t(five):
    qw, d, sq, dmin, zq = five
    return (torch.from_numpy(qw.copy()), torch.from_numpy(d.copy()), torch.from_numpy(sq.copy()),
            torch.from_numpy(dmin.copy()), torch.from_numpy(zq.copy()))


def set_timer(timer):
    pass


def launch_count():
    return 0


def hessian_update(H, X, beta, alpha):
    Hn = _np(H)
    orc.hessian_update(Hn, _np(X.float()), float(beta), float(alpha))
    H.copy_(torch.from_numpy(Hn))


def pre_step(H, W):
    dead = torch.diagonal(H) == 0
    W[:, dead] = 0
    idx = torch.nonzero(dead).flatten()
    H[idx, idx] = 1


def prepare(H, W, rel_damp, stream=None, slot=0):
    U, Hd, _, bad = orc.prepare(_np(H), _np(W), float(rel_damp))
    H.copy_(torch.from_numpy(Hd))
    return torch.from_numpy(U), torch.tensor([int(bad)], dtype=torch.int32)


def gptq_quantize(W, U, q_type, block_size=128, rmin=-1.0, rdelta=0.1, nstep=20, mode=0, packed=True, stream=None,
                  wdeq_dtype=None, search_flags=False, static_groups=False, perm=None, ws_slot=None):
    out = orc.gptq_step(_np(W), _np(U), int(q_type), block_size, rmin, rdelta, nstep, static_groups=static_groups,
                        perm=None if perm is None else _np(perm))
    five = _five_t(out[:5])
    pk = torch.from_numpy(orc.pack(int(q_type), *out[:5])) if packed else None
    wd = torch.from_numpy(out[5]).to(wdeq_dtype) if wdeq_dtype is not None else None
    return five + (pk, wd, None)


def rtn_quantize(W, q_type, rmin=-1.0, rdelta=0.1, nstep=20, packed=True, wdeq_dtype=None, native_arith=False):
    out = orc.rtn_quantize(_np(W.float()), int(q_type), rmin, rdelta, nstep, bf16=bool(native_arith and W.dtype == torch.bfloat16),
                           fp16=bool(native_arith and W.dtype == torch.float16))
    five = _five_t(out)
    pk = torch.from_numpy(orc.pack(int(q_type), *out)) if packed else None
    wd = torch.from_numpy(orc.dequantize(int(q_type), *out)).to(wdeq_dtype) if wdeq_dtype is not None else None
    return five + (pk, wd)


def get_scale_and_zero(x, q_type, rmin=-1.0, rdelta=0.1, nstep=20):
    d, sq, dmin, zq = orc.get_scale_and_zero(_np(x), int(q_type), rmin, rdelta, nstep)
    return (torch.from_numpy(d.copy()), torch.from_numpy(sq.copy()), torch.from_numpy(dmin.copy()),
            torch.from_numpy(zq.copy()))


def dequantize(q_type, qweight, d, sq, dmin, zq, out_dtype=torch.float32):
    return torch.from_numpy(orc.dequantize(int(q_type), _np(qweight), _np(d), _np(sq), _np(dmin), _np(zq))).to(out_dtype)


def pack(q_type, qweight, d, sq, dmin=None, zq=None):
    if dmin is None:
        dmin = torch.zeros_like(d)
    if zq is None:
        zq = torch.zeros_like(sq)
    return torch.from_numpy(orc.pack(int(q_type), _np(qweight), _np(d), _np(sq), _np(dmin), _np(zq)))


def install(monkeypatch):
    """Route the host mirror's `ops` calls to the oracle (tests only)."""
    import sys
    from gptq_gguf_toolkit_b200 import gptq as gq_gptq
    from gptq_gguf_toolkit_b200 import quant_utils as gq_qu
    from gptq_gguf_toolkit_b200 import quantizer as gq_quantizer
    me = sys.modules[__name__]
    for mod in (gq_gptq, gq_quantizer, gq_qu):
        monkeypatch.setattr(mod, "ops", me)
