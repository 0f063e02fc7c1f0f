"""GPU: the arithmetic helpers of csrc/f32x2.cuh that the column-loop kernel's bit-exactness rests on.

  * DivBy (division by a pre-inverted divisor, Markstein correction) must equal the IEEE fp32 quotient bit for bit
    over the whole range, including zeros (sign of zero), denormals, huge/tiny operands and all-ones significands;
  * f2_mul_nofuse + f2_sub (packed pairs) must round TWICE like the reference's `w - err * u` on CPU, i.e. must not
    be contracted into an FMA by ptxas.
The oracle is numpy's float32 arithmetic (IEEE, round to nearest even)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import _lib
    lib = _lib.load()
    lib.gq_debug_divby.restype = C.c_int
    lib.gq_debug_divby.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    lib.gq_debug_mulsub2.restype = C.c_int
    lib.gq_debug_mulsub2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    lib.gq_debug_rint_clamp.restype = C.c_int
    lib.gq_debug_rint_clamp.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    return lib


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _rand_f32(rng, n, emin, emax):
    mant = rng.integers(0, 1 << 23, n, dtype=np.uint32)
    exp = rng.integers(emin + 127, emax + 128, n, dtype=np.uint32)
    sign = rng.integers(0, 2, n, dtype=np.uint32)
    return ((sign << 31) | (exp << 23) | mant).view(np.float32)


def test_divby_equals_ieee_division(lib):
    rng = np.random.default_rng(0)
    n = 1 << 22
    parts_a, parts_b = [], []
    # bulk: moderate exponents (the fast path)
    parts_a.append(_rand_f32(rng, n, -30, 30)); parts_b.append(_rand_f32(rng, n, -30, 30))
    # divisors with (nearly) all-ones significands, dividends with tiny significands
    b = _rand_f32(rng, n, -20, 20).view(np.uint32)
    b = (b | np.uint32(0x7FFF00)) | rng.integers(0, 256, n, dtype=np.uint32)
    a = _rand_f32(rng, n, -20, 20).view(np.uint32) & np.uint32(0xFF8000FF)
    parts_a.append(a.view(np.float32)); parts_b.append(b.view(np.float32))
    # full exponent range: overflow / underflow / denormal results take the IEEE path
    parts_a.append(_rand_f32(rng, n, -126, 127)); parts_b.append(_rand_f32(rng, n, -126, 127))
    # special values
    sp = np.array([0.0, -0.0, 1e-45, -1e-45, 1.17549435e-38, 3.4e38, -3.4e38, 1.0, -1.0, 1e-9, 0.5, 3.0], dtype=np.float32)
    A, B = np.meshgrid(sp, sp[sp != 0])
    parts_a.append(A.ravel()); parts_b.append(B.ravel())
    a = np.concatenate(parts_a); b = np.concatenate(parts_b)
    with np.errstate(all="ignore"):
        want = (a / b).astype(np.float32)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty_like(ta)
    assert lib.gq_debug_divby(ta.data_ptr(), tb.data_ptr(), out.data_ptr(), a.size, torch.cuda.current_stream().cuda_stream) == 0
    got = out.cpu().numpy()
    ok = (_bits(got) == _bits(want)) | (np.isnan(got) & np.isnan(want))
    assert ok.all(), f"{(~ok).sum()} of {a.size} quotients differ, e.g. a={a[~ok][:3]} b={b[~ok][:3]}"


def test_packed_mul_sub_is_not_contracted(lib):
    rng = np.random.default_rng(1)
    n = 1 << 22
    a = rng.standard_normal(n).astype(np.float32)
    e = rng.standard_normal(n).astype(np.float32)
    u = rng.standard_normal(n).astype(np.float32)
    want = a - (e * u)                                   # two float32 roundings
    fused = (a.astype(np.float64) - e.astype(np.float64) * u.astype(np.float64)).astype(np.float32)
    assert (_bits(want) != _bits(fused)).any()           # the test can tell the two apart
    ta, te, tu = (torch.from_numpy(x).cuda() for x in (a, e, u))
    out = torch.empty_like(ta)
    assert lib.gq_debug_mulsub2(ta.data_ptr(), te.data_ptr(), tu.data_ptr(), out.data_ptr(), n,
                                torch.cuda.current_stream().cuda_stream) == 0
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(want))


def test_conversion_free_rint_clamp_equals_rintf_on_every_float(lib):
    """kq_rint_clamp (two FADDs with 1.5 * 2^23 instead of FRND) and kq_sq_u8 (the uint8 square without F2I / I2F) replace rintf and
    the integer square in the scale search and the column steps: checked on the device for ALL 2^32 bit patterns (NaNs skipped) and
    the clamp ranges of the five formats and of the scale codes."""
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for lo, hi in ((0.0, 3.0), (0.0, 15.0), (0.0, 31.0), (0.0, 63.0), (-4.0, 3.0), (-32.0, 31.0)):
        assert lib.gq_debug_rint_clamp(0, 0x7FFFFFFF, lo, hi, bad.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
