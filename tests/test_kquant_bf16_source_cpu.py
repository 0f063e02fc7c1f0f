"""The DEVICE source of the experimental bf16-arithmetic scale search (gptq_gguf_toolkit_b200/csrc/kquant_bf16.cuh) compiled
for the HOST through a shim of the CUDA intrinsics (tests/helpers/host_shim) and checked against the reference golden
(tests/golden/rtn_bf16.npz): the arithmetic the GPU kernel will execute is verified bit for bit before it ever runs on a GPU
(it was written after round 1's GPU budget was spent).  What this does NOT cover: the kernel around it (tile loads, indexing,
launch) -- that is tests/test_gpu_parity.py::test_rtn_native_bf16_arithmetic_matches_reference_golden (opt-in)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "helpers", "host_shim")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libkqb_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I", SHIM,
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(SHIM, "kquant_bf16_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    return C.CDLL(so)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_device_source_matches_reference_golden(host_lib, golden_dir, tname, dtype):
    g = np.load(os.path.join(golden_dir, f"rtn_{dtype}.npz"))
    if dtype == "bf16":
        W = np.ascontiguousarray((g["W_bf16_bits"].astype(np.uint32) << 16).view(np.float32))
    else:
        W = np.ascontiguousarray(g["W_f16_bits"].view(np.float16).astype(np.float32))
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    gs = 32 if tname in ("Q4_K", "Q5_K") else 16
    d_row, d_col = W.shape
    d = np.zeros((d_row, d_col // 256), np.uint16)
    dmin = np.zeros_like(d)
    sq = np.zeros((d_row, d_col // gs), np.uint8)
    zq = np.zeros_like(sq)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = host_lib.host_scales_native(C.c_int(2 if dtype == "bf16" else 1), C.c_int(qt), p(W, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1),
                                   C.c_int(20), p(d, C.c_uint16), p(dmin, C.c_uint16), p(sq, C.c_uint8), p(zq, C.c_uint8))
    assert rc == 0
    assert np.array_equal(d, g[f"{tname}_d"]), "super_group_scale"
    assert np.array_equal(dmin, g[f"{tname}_dmin"]), "super_group_zero"
    assert np.array_equal(sq, g[f"{tname}_sq"].view(np.uint8)), "group_scale_quant"
    assert np.array_equal(zq, g[f"{tname}_zq"].view(np.uint8)), "group_zero_quant"
