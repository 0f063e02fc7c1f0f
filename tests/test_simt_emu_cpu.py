"""Kernels written without GPU access, run on the host through a small SIMT emulator (tests/helpers/simt_emu: one fiber per
CUDA thread, __syncthreads() = barrier between fibers) to catch indexing / algorithm mistakes before they cost GPU minutes.
Covered: the diagonal-block kernels of gq_prepare -- chol_diag_v2.cuh, the register-resident chol_diag_v3.cuh and the two-level
chol_diag_v4.cuh (GQ_DIAG_V2 = 1 / 3 / 4, the last one is the default) --
(128 x 128 Cholesky factor + its inverse + the inverse's transpose, against a double-precision factorisation)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "helpers", "simt_emu")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libchol_v3_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-x", "c++", "-I", EMU,
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(EMU, "chol_diag_v3_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.fixture(scope="module")
def lib_v4(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libchol_v4_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-x", "c++", "-I", EMU,
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(EMU, "chol_diag_v4_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.fixture(scope="module")
def lib_v2(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libchol_v2_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", EMU,
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(EMU, "chol_diag_v2_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.mark.parametrize("variant", ["v2", "v3", "v4_default"])
@pytest.mark.parametrize("k0", [0, 128])
def test_chol_diag_on_the_emulator(lib, lib_v2, lib_v4, k0, variant):
    rng = np.random.default_rng(k0 + 1)
    n, nb = 384, 128
    M = rng.standard_normal((nb, 3 * nb))
    H = M @ M.T / (3 * nb) + 0.05 * np.eye(nb)
    A = np.zeros((n, n), np.float32)
    A[k0:k0 + nb, k0:k0 + nb] = H.astype(np.float32)
    Binv = np.full((n, n), 7.0, np.float32)
    BinvT = np.full((n, n), 7.0, np.float32)
    flag = np.zeros(1, np.int32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    run = {"v2": lib_v2.run_chol_diag_v2, "v3": lib.run_chol_diag_v3, "v4_default": lib_v4.run_chol_diag_v4}[variant]
    run(p(A, C.c_float), p(Binv, C.c_float), p(BinvT, C.c_float), C.c_long(n), C.c_int(k0), p(flag, C.c_int))
    L = np.linalg.cholesky(H)
    X = np.linalg.inv(L)
    gL = np.tril(A[k0:k0 + nb, k0:k0 + nb].astype(np.float64))
    gX = Binv[k0:k0 + nb, k0:k0 + nb].astype(np.float64)
    gXT = BinvT[k0:k0 + nb, k0:k0 + nb].astype(np.float64)
    assert flag[0] == 0
    assert np.abs(gL - L).max() <= 1e-5 * np.abs(L).max()
    assert np.abs(gX - X).max() <= 1e-4 * np.abs(X).max()
    assert np.array_equal(gXT, gX.T), "BinvT must be the exact transpose of Binv"
    assert np.all(np.triu(gX, 1) == 0), "inv(L) is written with an explicit zero upper triangle"
    # nothing outside the block is touched
    mask = np.ones((n, n), bool)
    mask[k0:k0 + nb, k0:k0 + nb] = False
    assert np.all(Binv[mask] == 7.0) and np.all(BinvT[mask] == 7.0) and np.all(A[mask] == 0.0)


@pytest.fixture(scope="module")
def lib_upd(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libupd2_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", EMU,
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(EMU, "exact_update_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.mark.parametrize("kernel", ["run_exact_update_v1"], ids=["shipped"])
@pytest.mark.parametrize("d_row,d_col,c", [(40, 1024, 256), (33, 768, 0), (64, 1280, 512)])
def test_exact_update_on_the_emulator(lib_upd, d_row, d_col, c, kernel):
    """The SHIPPED trailing-update kernel body (csrc/rank_update.cuh: exact_update_body + rank_update<>, what
    exact_update_kernel and the left-looking loop of gptq_layer_kernel execute): the rank-256 trailing update of the exact schedule, W[:, c+256:] <- (W - chain(E[:, c:c+128], U[c:c+128, :])) - chain(E[:, c+128:c+256], U[c+128:c+256, :]) with
    E = W[:, c:c+256], every chain a single-accumulator fp32 FMA chain in ascending k -- bit for bit against a plain
    restatement, ragged row count included; columns left of c+256 and rows must be untouched."""
    import math
    rng = np.random.default_rng(d_row + d_col + c)
    W = (rng.standard_normal((d_row, d_col)) * 0.05).astype(np.float32)
    U = np.triu(rng.standard_normal((d_col, d_col)) * 0.02).astype(np.float32)
    ref = W.copy()
    E = W[:, c:c + 256].astype(np.float32)
    for half in range(2):
        acc = np.zeros((d_row, d_col - c - 256), np.float32)
        for k in range(128 * half, 128 * half + 128):                       # fmaf chain, ascending k, elementwise exact FMA
            e = E[:, k:k + 1].astype(np.float64)
            u = U[c + k:c + k + 1, c + 256:].astype(np.float64)
            acc = (e * u + acc.astype(np.float64)).astype(np.float32)       # fp64 product+sum of fp32 operands, rounded once = fmaf
        ref[:, c + 256:] = ref[:, c + 256:] - acc
    got = W.copy()
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    getattr(lib_upd, kernel)(p(got, C.c_float), p(U, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_int(c))
    assert np.array_equal(got[:, :c + 256], W[:, :c + 256]), "columns up to the finished super-block must not change"
    assert np.array_equal(got, ref)


@pytest.fixture(scope="module")
def lib_rtn(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "librtn_native_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", EMU,
           "-I", os.path.join(ROOT, "tests", "helpers", "host_shim"), "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"),
           os.path.join(EMU, "rtn_native_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_rtn_native_kernel_on_the_emulator(lib_rtn, golden_dir, tname, dtype):
    """gptq_gguf_toolkit_b200/csrc/rtn_native.cuh (gq_rtn_quantize_native, experimental): the WHOLE kernel -- tile load, search in
    the weight's own arithmetic, double quantisation, fp32 quantize(), codes, GGUF bytes, 16-bit dequantised weights -- against
    the reference golden of a bf16 / fp16 weight (five tensors), the oracle's packer and dequantiser."""
    import torch
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, f"rtn_{dtype}.npz"))
    bits = np.ascontiguousarray(g[f"W_{dtype}_bits"])
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    gs = 32 if tname in ("Q4_K", "Q5_K") else 16
    ts = {"Q2_K": 84, "Q3_K": 110, "Q4_K": 144, "Q5_K": 176, "Q6_K": 210}[tname]
    d_row, d_col = bits.shape
    nsb = d_col // 256
    qw = np.zeros((d_row, d_col), np.uint8)
    d = np.zeros((d_row, nsb), np.uint16)
    dmin = np.zeros_like(d)
    sq = np.zeros((d_row, d_col // gs), np.uint8)
    zq = np.zeros_like(sq)
    pk = np.zeros((d_row, nsb * ts), np.uint8)
    wd = np.zeros((d_row, d_col), np.uint16)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib_rtn.run_rtn_native(C.c_int(2 if dtype == "bf16" else 1), C.c_int(qt), p(bits, C.c_uint16), C.c_int(d_row), C.c_int(d_col),
                                C.c_double(-1.0), C.c_double(0.1), C.c_int(20), p(qw, C.c_uint8), p(d, C.c_uint16), p(dmin, C.c_uint16),
                                p(sq, C.c_uint8), p(zq, C.c_uint8), p(pk, C.c_uint8), p(wd, C.c_uint16))
    assert rc == 0
    for k, a in (("qweight", qw), ("d", d), ("sq", sq), ("dmin", dmin), ("zq", zq)):
        assert np.array_equal(a.view(np.uint8), g[f"{tname}_{k}"].view(np.uint8)), k
    cd = np.uint8 if tname in ("Q2_K", "Q4_K", "Q5_K") else np.int8
    five = (qw.view(cd), d.view(np.float16), sq.view(cd), dmin.view(np.float16), zq.view(cd))
    assert np.array_equal(pk, orc.pack(qt, *five)), "GGUF block bytes"
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    # compared as VALUES: a code 0 that came from rint(-0.2) keeps its sign through s * q - z (-0.0), the reference
    # dequantises the stored integer code (+0.0); torch.equal treats the two as equal, and so does every consumer
    want = torch.from_numpy(orc.dequantize(qt, *five)).to(tdt)
    got = torch.from_numpy(wd.view(np.int16).copy()).view(tdt)
    assert torch.equal(got, want), "dequantised weights in the weight's dtype"


@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_shipped_rtn_kernel_on_the_emulator(lib_rtn, golden_dir, tname):
    """The SHIPPED RTN kernel body (csrc/rtn_native.cuh: rtn_body, what rtn_kernel / gq_rtn_quantize execute): tile load, fp32
    search, double quantisation, quantize(), codes, GGUF bytes, dequantised weights -- against the reference golden of the
    non-block path (tests/golden/rtn.npz, ragged 100-row weight), the oracle's packer and dequantiser."""
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "rtn.npz"))
    W = np.ascontiguousarray(g["W"], dtype=np.float32)
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    gs = 32 if tname in ("Q4_K", "Q5_K") else 16
    ts = {"Q2_K": 84, "Q3_K": 110, "Q4_K": 144, "Q5_K": 176, "Q6_K": 210}[tname]
    d_row, d_col = W.shape
    nsb = d_col // 256
    qw = np.zeros((d_row, d_col), np.uint8)
    d = np.zeros((d_row, nsb), np.uint16)
    dmin = np.zeros_like(d)
    sq = np.zeros((d_row, d_col // gs), np.uint8)
    zq = np.zeros_like(sq)
    pk = np.zeros((d_row, nsb * ts), np.uint8)
    wd = np.zeros((d_row, d_col), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib_rtn.run_rtn_fp32(C.c_int(qt), p(W, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1), C.c_int(20),
                              p(qw, C.c_uint8), p(d, C.c_uint16), p(dmin, C.c_uint16), p(sq, C.c_uint8), p(zq, C.c_uint8),
                              p(pk, C.c_uint8), p(wd, C.c_float))
    assert rc == 0
    for k, a in (("qweight", qw), ("d", d), ("sq", sq), ("dmin", dmin), ("zq", zq)):
        r = g[f"{tname}_ieee_{k}"]
        r = r.view(np.uint16) if r.dtype == np.float16 else r
        assert np.array_equal(a.view(np.uint8), r.view(np.uint8).reshape(a.shape[0], -1)), k
    cd = np.uint8 if tname in ("Q2_K", "Q4_K", "Q5_K") else np.int8
    five = (qw.view(cd), d.view(np.float16), sq.view(cd), dmin.view(np.float16), zq.view(cd))
    assert np.array_equal(pk, orc.pack(qt, *five))
    assert np.array_equal(wd, orc.dequantize(qt, *five))        # value comparison (-0.0 == +0.0, see above)


# ------------------------------------------------------------------------------------------------
# the SHIPPED fused column-loop kernel against the reference goldens (B1), on the emulator
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def lib_layer(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libgptq_layer_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", EMU,
           "-I", os.path.join(ROOT, "tests", "helpers", "host_shim"), "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"),
           os.path.join(EMU, "gptq_layer_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.mark.parametrize("schedule", ["left_looking", "right_looking"])
@pytest.mark.parametrize("case", ["b1_a.npz", "b1_b.npz"])
@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_shipped_column_loop_kernel_matches_reference_golden(lib_layer, golden_dir, tname, case, schedule):
    """gptq_gguf_toolkit_b200/csrc/gptq_layer_kernel.cuh (gptq_layer_kernel<QT>: per-256 scale / min search, the 256 dependent
    column steps with their warp shuffles and exact divisions, in-super-block and left-looking rank-k updates, GGUF bit-pack,
    dequantised write-back) and rank_update.cuh (the trailing update of the right-looking schedule) -- the source that ships in
    libgq -- executed on the SIMT emulator on the reference's own B1 goldens (gptq.py:146-295 on CPU, ragged row counts):
    codes, the four scale tensors, the packed GGUF bytes and the dequantised weights must be bit-identical, in both schedules."""
    g = np.load(os.path.join(golden_dir, case))
    assert int(g["block_size"]) == 128
    W = np.ascontiguousarray(g["W"], dtype=np.float32).copy()
    U = np.ascontiguousarray(g["U_colmajor_T"].T, dtype=np.float32)      # the reference's U is column-major
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    gs = 32 if tname in ("Q4_K", "Q5_K") else 16
    ts = {"Q2_K": 84, "Q3_K": 110, "Q4_K": 144, "Q5_K": 176, "Q6_K": 210}[tname]
    d_row, d_col = W.shape
    nsb = d_col // 256
    qw = np.zeros((d_row, d_col), np.uint8)
    d = np.zeros((d_row, nsb), np.uint16)
    dmin = np.zeros_like(d)
    sq = np.zeros((d_row, d_col // gs), np.uint8)
    zq = np.zeros_like(sq)
    pk = np.zeros((d_row, nsb * ts), np.uint8)
    wd = np.zeros((d_row, d_col), np.float32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib_layer.run_gptq_layer(C.c_int(qt), C.c_int(1 if schedule == "right_looking" else 0), p(W, C.c_float), p(U, C.c_float),
                                  C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1), C.c_int(20), p(qw, C.c_uint8),
                                  p(d, C.c_uint16), p(sq, C.c_uint8), p(dmin, C.c_uint16), p(zq, C.c_uint8), p(pk, C.c_uint8), p(wd, C.c_float))
    assert rc == 0
    for k, a in (("qweight", qw), ("d", d), ("sq", sq), ("dmin", dmin), ("zq", zq), ("packed", pk)):
        r = g[f"{tname}_ieee_{k}"]
        r = r.view(np.uint16) if r.dtype == np.float16 else r
        assert np.array_equal(a.view(np.uint8), r.view(np.uint8).reshape(a.shape[0], -1)), f"{k} differs from the reference"
    assert np.array_equal(wd, g[f"{tname}_ieee_dequant"]), "dequantised weights"


@pytest.mark.parametrize("schedule", ["left_looking", "right_looking"])
@pytest.mark.parametrize("variant", ["static", "actorder"])
@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_shipped_column_loop_kernel_variants_match_reference_golden(lib_layer, golden_dir, tname, variant, schedule):
    """static_groups and act_order (gptq.py:184-216, 233-238, 273-277) through the shipped kernels on the emulator, in the flow
    of gq_gptq_quantize_ex / ops.gptq_quantize: scales searched up front by the RTN search over the un-permuted weights, the loop
    on W[:, perm] with the factor of H[perm][:, perm], codes un-permuted afterwards -- against the reference's own outputs
    (tests/golden/variants_a.npz).  Q3_K ignores both options like the reference (gptq.py:204-206)."""
    g = np.load(os.path.join(golden_dir, "variants_a.npz"))
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    gs = 32 if tname in ("Q4_K", "Q5_K") else 16
    W0 = np.ascontiguousarray(g["W"], dtype=np.float32)
    d_row, d_col = W0.shape
    nsb = d_col // 256
    uses_perm = variant == "actorder" and tname != "Q3_K"
    static = tname != "Q3_K"
    U = np.ascontiguousarray(g["U_perm"] if uses_perm else g["U_plain"], dtype=np.float32)
    perm = np.ascontiguousarray(g["perm"].astype(np.int32)) if uses_perm else None
    qw = np.zeros((d_row, d_col), np.uint8)
    d = np.zeros((d_row, nsb), np.uint16)
    dmin = np.zeros_like(d)
    sq = np.zeros((d_row, d_col // gs), np.uint8)
    zq = np.zeros_like(sq)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    if static:
        assert lib_layer.run_search_all(C.c_int(qt), p(W0, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1),
                                        C.c_int(20), p(d, C.c_uint16), p(dmin, C.c_uint16), p(sq, C.c_uint8), p(zq, C.c_uint8)) == 0
    W = np.ascontiguousarray(W0[:, perm] if uses_perm else W0).copy()
    rc = lib_layer.run_gptq_layer_ex(C.c_int(qt), C.c_int(1 if schedule == "right_looking" else 0), p(W, C.c_float), p(U, C.c_float),
                                     C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1), C.c_int(20),
                                     C.c_int((2 if uses_perm else 1) if static else 0), p(perm, C.c_int) if uses_perm else None,
                                     p(qw, C.c_uint8), p(d, C.c_uint16), p(sq, C.c_uint8), p(dmin, C.c_uint16), p(zq, C.c_uint8), None, None)
    assert rc == 0
    if uses_perm:                      # gptq.py:276-277: qweight[:, invperm]
        out = np.zeros_like(qw)
        out[:, perm] = qw
        qw = out
    for k, a in (("qweight", qw), ("d", d), ("sq", sq), ("dmin", dmin), ("zq", zq)):
        r = g[f"{variant}_{tname}_{k}"]
        r = r.view(np.uint16) if r.dtype == np.float16 else r
        assert np.array_equal(a.view(np.uint8), r.view(np.uint8).reshape(a.shape[0], -1)), f"{k} differs from the reference"


@pytest.mark.parametrize("schedule", ["left_looking", "right_looking"])
def test_shipped_column_loop_kernel_unsafe_division_rerun(lib_layer, schedule):
    """The rare path of the column steps: DivBy::div_fast flags quotients it cannot guarantee (a divisor whose significand is all
    ones, dividends below 2^-60 or above 2^60), the warp then re-runs the 128-column block with IEEE divisions
    (serial_block<>: __any_sync + serial_block_impl<QT, true>).  Crafted input: all-ones significands on the diagonal of U
    (scales with all-ones significands come by themselves) and a row of tiny (1e-22) weights next to ordinary ones -- the
    emulated shipped kernel must still equal the oracle bit for bit.  (Weights that overflow fp16 scales end in NaNs whose sign
    bit is implementation noise; not part of this test.)"""
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    d_row, d_col = 40, 512
    W = (rng.standard_normal((d_row, d_col)) * 0.05).astype(np.float32)
    W[3] *= np.float32(1e-22)
    W[17, 100:140] = 0.0
    U = (np.triu(rng.standard_normal((d_col, d_col)) * 0.02)).astype(np.float32)
    diag = (0.5 + rng.random(d_col)).astype(np.float32)
    diag[::3] = np.frombuffer(np.array([0x3FFFFFFF, 0x3F7FFFFF, 0x407FFFFF], np.uint32).tobytes(), np.float32)[rng.integers(0, 3, len(diag[::3]))]
    U[np.arange(d_col), np.arange(d_col)] = diag
    for tname, qt, gs, ts in (("Q4_K", 12, 32, 144), ("Q6_K", 14, 16, 210)):
        ref = orc.gptq_step(W.copy(), U, qt)
        Wk = W.copy()
        nsb = d_col // 256
        qw = np.zeros((d_row, d_col), np.uint8)
        d = np.zeros((d_row, nsb), np.uint16)
        dmin = np.zeros_like(d)
        sq = np.zeros((d_row, d_col // gs), np.uint8)
        zq = np.zeros_like(sq)
        pk = np.zeros((d_row, nsb * ts), np.uint8)
        wd = np.zeros((d_row, d_col), np.float32)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        rc = lib_layer.run_gptq_layer(C.c_int(qt), C.c_int(1 if schedule == "right_looking" else 0), p(Wk, C.c_float), p(U, C.c_float),
                                      C.c_int(d_row), C.c_int(d_col), C.c_double(-1.0), C.c_double(0.1), C.c_int(20), p(qw, C.c_uint8),
                                      p(d, C.c_uint16), p(sq, C.c_uint8), p(dmin, C.c_uint16), p(zq, C.c_uint8), p(pk, C.c_uint8), p(wd, C.c_float))
        assert rc == 0
        for k, a, r in (("qweight", qw, ref[0]), ("d", d, ref[1]), ("sq", sq, ref[2]), ("dmin", dmin, ref[3]), ("zq", zq, ref[4])):
            r = r.view(np.uint16) if r.dtype == np.float16 else r
            assert np.array_equal(a.view(np.uint8), np.ascontiguousarray(r).view(np.uint8).reshape(a.shape[0], -1)), f"{tname}.{k}"
        assert np.array_equal(pk, orc.pack(qt, *ref[:5]))


def test_shipped_simt_hessian_kernel_on_the_emulator(tmp_path_factory):
    """sg::sgemm_kernel (csrc/sgemm.cuh) in the configuration gq_hessian_update uses for fp32 activations: H <- beta H + alpha
    X^T X over upper tiles with the mirrored store -- against float64, exact symmetry, two accumulation steps (GPTQ.update's
    running average, gptq.py:108-112), a token count that is not a multiple of the k tile."""
    so = str(tmp_path_factory.mktemp("emu") / "libsgemm_emu.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", EMU,
           "-I", os.path.join(ROOT, "tests", "helpers", "host_shim"), "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"),
           os.path.join(EMU, "sgemm_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    lib = C.CDLL(so)
    lib.run_hessian_simt.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_long, C.c_int, C.c_float, C.c_float]
    rng = np.random.default_rng(5)
    d_col = 256
    H = np.zeros((d_col, d_col), np.float32)
    want = np.zeros((d_col, d_col), np.float64)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    n = 0
    for n_tok in (200, 77):
        X = (rng.standard_normal((n_tok, d_col)) * np.exp(0.3 * rng.standard_normal(d_col))).astype(np.float32)
        beta, alpha = n / (n + 1), 2.0 / (n + 1)
        lib.run_hessian_simt(p(H), p(X), n_tok, d_col, beta, alpha)
        want = beta * want + alpha * (X.astype(np.float64).T @ X.astype(np.float64))
        n += 1
    assert np.array_equal(H, H.T), "H must be exactly symmetric"
    assert np.abs(H - want).max() <= 5e-5 * np.abs(want).max()


def test_div_chain_range_bookkeeping(lib_layer):
    """DivRange / div_chain (csrc/f32x2.cuh): the column steps' Markstein quotient is only proven equal to IEEE division for a zero
    dividend or one inside (2^-60, 2^60); the two integer min / max accumulators must flag exactly the rest (the block is then re-run
    with IEEE divisions), and inside the range the quotient must be the correctly rounded one."""
    f = np.float32
    safe = [0.0, -0.0, 1.0, -3.5, np.nextafter(f(2.0 ** 60), f(0)), -np.nextafter(f(2.0 ** 60), f(0)),
            np.nextafter(f(2.0 ** -60), f(1)), -np.nextafter(f(2.0 ** -60), f(1)), 1e-10, 6e17]
    unsafe = [2.0 ** 60, -(2.0 ** 60), 2.0 ** -60, -(2.0 ** -60), 1e-30, 1e-45, -1e-45, 3e38, np.inf, -np.inf, np.nan]
    a = np.array(safe + unsafe, dtype=np.float32)
    rng = np.random.default_rng(0)
    a = np.concatenate([a, (rng.standard_normal(4096) * np.exp(8 * rng.standard_normal(4096))).astype(np.float32)])
    ok = np.zeros(a.size, np.int32)
    q = np.zeros(a.size, np.float32)
    for b in (f(0.37), f(-1.7e-3), f(911.0)):
        lib_layer.run_div_chain(a.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(a.size), C.c_float(b),
                                ok.ctypes.data_as(C.POINTER(C.c_int)), q.ctypes.data_as(C.POINTER(C.c_float)))
        aa = np.abs(a)
        want_ok = (a == 0) | ((aa > f(2.0 ** -60)) & (aa < f(2.0 ** 60)))
        assert np.array_equal(ok.astype(bool), want_ok), b
        with np.errstate(all="ignore"):
            ref = (a / b).astype(np.float32)
        m = want_ok
        assert np.array_equal(q[m], ref[m]), b            # as values: a zero quotient's sign is allowed to differ
