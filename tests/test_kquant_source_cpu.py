"""The K-quant arithmetic of the PRODUCT's device source (gptq_gguf_toolkit_b200/csrc/kquant.cuh: scale / min search, double
quantisation of the group scales, quantise / dequantise, GGUF bit-pack) compiled for the HOST through a shim of the CUDA
intrinsics (tests/helpers/host_shim) and checked, without a GPU, against
  * the reference-generated golden of the non-block RTN path (tests/golden/rtn.npz: the reference's five tensors),
  * the reference's own packed bytes and dequantised weights of the B1 goldens (b1_a.npz), and
  * the CPU oracle on random ill-scaled weights (all five types).
This is the same source the rtn / gptq_layer kernels inline; the kernels around it (tiles, indexing, the column loop) are
covered by the `-m gpu` tests."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "helpers", "host_shim")
TYPES = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}
TS = {"Q2_K": 84, "Q3_K": 110, "Q4_K": 144, "Q5_K": 176, "Q6_K": 210}
GS = {"Q2_K": 16, "Q3_K": 16, "Q4_K": 32, "Q5_K": 32, "Q6_K": 16}
KEYS = ["qweight", "d", "sq", "dmin", "zq"]


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libkq_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", SHIM,
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(SHIM, "kquant_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    return C.CDLL(so)


def _run(lib, W, tname, want_packed=True):
    W = np.ascontiguousarray(W, dtype=np.float32)
    d_row, d_col = W.shape
    nsb, ng = d_col // 256, d_col // GS[tname]
    qw = np.zeros((d_row, d_col), np.uint8)
    d = np.zeros((d_row, nsb), np.uint16)
    dmin = np.zeros_like(d)
    sq = np.zeros((d_row, ng), np.uint8)
    zq = np.zeros_like(sq)
    pk = np.zeros((d_row, nsb * TS[tname]), np.uint8)
    wd = np.zeros((d_row, d_col), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.host_rtn(C.c_int(TYPES[tname]), p(W, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1),
                      C.c_int(20), p(qw, C.c_uint8), p(d, C.c_uint16), p(dmin, C.c_uint16), p(sq, C.c_uint8), p(zq, C.c_uint8),
                      p(pk, C.c_uint8), p(wd, C.c_float))
    assert rc == 0
    return {"qweight": qw, "d": d, "sq": sq, "dmin": dmin, "zq": zq, "packed": pk, "wdeq": wd}


def _raw(a):
    a = np.asarray(a)
    return (a.view(np.uint16) if a.dtype == np.float16 else a).view(np.uint8).reshape(a.shape[0], -1)


def test_shim_format_table_matches_common_cuh():
    pat = re.compile(r"template <> struct Fmt<GQ_(Q\d_K)> \{ static constexpr int ([^;]+); \};")
    real = dict(pat.findall(open(os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc", "common.cuh")).read()))
    shim = dict(pat.findall(open(os.path.join(SHIM, "host_shim_intrinsics.h")).read()))
    assert len(real) == 5 and real == shim


@pytest.mark.parametrize("tname", list(TYPES))
def test_device_source_matches_reference_rtn_golden(host_lib, golden_dir, tname):
    g = np.load(os.path.join(golden_dir, "rtn.npz"))
    got = _run(host_lib, g["W"], tname)
    for k in KEYS:
        assert np.array_equal(_raw(got[k]), _raw(g[f"{tname}_ieee_{k}"])), k


@pytest.mark.parametrize("tname", list(TYPES))
def test_device_source_matches_oracle_and_packs_like_the_reference(host_lib, golden_dir, tname):
    rng = np.random.default_rng(TYPES[tname])
    W = (rng.standard_normal((48, 1024)) * 0.03 * np.exp(rng.standard_normal((48, 1)))).astype(np.float32)
    W[2, 256:512] = 0.0
    W[4, :] = 0.0625
    got = _run(host_lib, W, tname)
    ref = orc.rtn_quantize(W, TYPES[tname])
    for k, r in zip(KEYS, ref):
        assert np.array_equal(_raw(got[k]), _raw(r)), k
    assert np.array_equal(got["packed"], orc.pack(TYPES[tname], *ref))
    assert np.array_equal(got["wdeq"], orc.dequantize(TYPES[tname], *ref))
    # the reference's own pack_Q*K bytes for the B1 golden's five tensors: same bytes from kq_pack_byte via the oracle's
    # packer was pinned in test_oracle_golden.py; here: the device packer == the oracle's on the golden tensors too
    g = np.load(os.path.join(golden_dir, "b1_a.npz"))
    five = [g[f"{tname}_ieee_{k}"] for k in KEYS]
    cd = np.uint8 if tname in ("Q2_K", "Q4_K", "Q5_K") else np.int8
    assert np.array_equal(orc.pack(TYPES[tname], five[0].view(cd), five[1].view(np.float16), five[2].view(cd), five[3].view(np.float16),
                                   five[4].view(cd)), g[f"{tname}_ieee_packed"])


def test_conversion_free_rint_and_square_equal_the_plain_forms(host_lib):
    """kq_rint_clamp / kq_sq_u8 (kquant.cuh) replaced rintf and the uint8 square's int conversions in the search and the column steps:
    same values as clamp(rintf(v)) for EVERY float with 2^-3 <= |v| <= 64 (all bit patterns: ties, half-integers), for zeros,
    denormals, large / infinite inputs and around the 2^22 / 2^23 / 2^24 boundaries (an exhaustive run over all 2^32 bit patterns
    found no non-NaN difference; NaNs are left out: on the GPU fmaxf(NaN, lo) is lo in both forms, the host's inlined fmaxf is not
    IEEE for signalling NaNs); the wrapped square for every code."""
    lib = host_lib
    lib.host_check_rint_clamp.restype = C.c_long
    lib.host_check_rint_clamp.argtypes = [C.POINTER(C.c_uint32), C.c_long, C.c_float, C.c_float]
    lib.host_check_rint_clamp_range.restype = C.c_long
    lib.host_check_rint_clamp_range.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float]
    special = np.array([0, 1, 2, 0x007FFFFF, 0x00800000, 0x7F800000, 0x7F7FFFFF], dtype=np.uint32)
    edges = np.concatenate([np.float32(2.0 ** e).view(np.uint32) + np.arange(-64, 65, dtype=np.int64) for e in (21, 22, 23, 24, 25, 31, 60)]).astype(np.uint32)
    rng = np.random.default_rng(0)
    rand = rng.integers(0, 2 ** 32, size=1_000_000, dtype=np.uint64).astype(np.uint32)
    allbits = np.concatenate([special, special | np.uint32(0x80000000), edges, edges | np.uint32(0x80000000), rand])
    allbits = np.ascontiguousarray(allbits[(allbits & np.uint32(0x7FFFFFFF)) <= np.uint32(0x7F800000)])
    first, last = int(np.float32(0.125).view(np.uint32)), int(np.float32(64.0).view(np.uint32))
    for lo, hi in ((0.0, 3.0), (0.0, 15.0), (0.0, 31.0), (0.0, 63.0), (-4.0, 3.0), (-32.0, 31.0)):
        bad = lib.host_check_rint_clamp(allbits.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_long(allbits.size), C.c_float(lo), C.c_float(hi))
        assert bad == 0, (lo, hi, bad)
    for lo, hi in ((0.0, 15.0), (-32.0, 31.0)):
        assert lib.host_check_rint_clamp_range(C.c_uint32(first), C.c_uint32(last), C.c_float(lo), C.c_float(hi)) == 0
    assert lib.host_check_sq_u8() == 0
