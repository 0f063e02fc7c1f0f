"""ctypes binding of libgq.so (include/gq.h).  There is no CPU fallback: if the library is missing or
no CUDA device is present every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("GQ_LIB_PATH") or os.path.join(_HERE, "libgq.so")   # GQ_LIB_PATH: development variants

GQ_OK, GQ_ERR_INVALID, GQ_ERR_CUDA, GQ_ERR_UNSUPPORTED, GQ_ERR_WORKSPACE = 0, 1, 2, 3, 4
GQ_F32, GQ_F16, GQ_BF16 = 0, 1, 2
GQ_MODE_EXACT, GQ_MODE_FAST, GQ_MODE_EXACT_LEFT, GQ_MODE_EXACT_RIGHT = 0, 1, 2, 3

_DT = {torch.float32: GQ_F32, torch.float16: GQ_F16, torch.bfloat16: GQ_BF16}


class GQError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libgq status {status}: {msg}")
        self.status = status


_vp, _i, _l, _f, _d, _sz = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/gq.h
SIGNATURES = {
    "gq_abi_version": (_i, []),
    "gq_last_error": (C.c_char_p, []),
    "gq_format_info": (_i, [_i, C.POINTER(_i)]),
    "gq_device_count": (_i, []),
    "gq_launch_count": (_l, []),
    "gq_hessian_workspace_bytes": (_sz, [_l, _i, _i]),
    "gq_hessian_update": (_i, [_vp, _vp, _l, _i, _i, _f, _f, _vp, _sz, _vp]),
    "gq_pre_step": (_i, [_vp, _vp, _i, _i, _vp]),
    "gq_prepare_workspace_bytes": (_sz, [_i]),
    "gq_prepare": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp, _sz, _vp, _vp]),
    "gq_gptq_workspace_bytes": (_sz, [_i, _i, _i]),
    "gq_gptq_quantize": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "gq_gptq_quantize_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "gq_profile_enable": (None, [_i]),
    "gq_profile_read": (_i, [C.POINTER(_f), C.POINTER(_i)]),
    "gq_profile_read3": (_i, [C.POINTER(_f), C.POINTER(_i)]),
    "gq_rtn_quantize": (_i, [_vp, _i, _i, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "gq_rtn_quantize_native": (_i, [_vp, _i, _i, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "gq_get_scale_and_zero": (_i, [_vp, _l, _i, _i, _d, _d, _i, _vp, _vp, _l, _vp, _vp, _l, _vp, _vp]),
    "gq_dequantize": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "gq_pack": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libgq.so (raises if it has not been built: run `python -m gptq_gguf_toolkit_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise GQError(GQ_ERR_CUDA, f"{SO_PATH} not built; run python -m gptq_gguf_toolkit_b200.build (no CPU fallback)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.gq_abi_version() != 1:
            raise GQError(GQ_ERR_INVALID, "libgq ABI version mismatch")
        _lib = lib
    return _lib


def check(status: int):
    if status != GQ_OK:
        raise GQError(status, load().gq_last_error().decode())


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise GQError(GQ_ERR_INVALID, f"unsupported dtype {dt}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GQError(GQ_ERR_CUDA, "libgq has no CPU fallback: tensors must live on a CUDA device")


def format_info(q_type: int) -> dict:
    out = (_i * 7)()
    check(load().gq_format_info(int(q_type), out))
    keys = ["bits", "qmin", "qmax", "scale_maxq", "group_size", "asym", "type_size"]
    return dict(zip(keys, list(out)))
