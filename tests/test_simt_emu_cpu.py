"""Kernels written without GPU access, run on the host through a small SIMT emulator (tests/helpers/simt_emu: one fiber per
CUDA thread, __syncthreads() = barrier between fibers) to catch indexing / algorithm mistakes before they cost GPU minutes.
Covered: gptq_gguf_toolkit_b200/csrc/chol_diag_v3.cuh, the register-resident diagonal-block kernel (GQ_DIAG_V2=3, experimental)
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


@pytest.mark.parametrize("k0", [0, 128])
def test_chol_diag_v3_on_the_emulator(lib, k0):
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
    lib.run_chol_diag_v3(p(A, C.c_float), p(Binv, C.c_float), p(BinvT, C.c_float), C.c_long(n), C.c_int(k0), p(flag, C.c_int))
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
           "-I", os.path.join(ROOT, "gptq_gguf_toolkit_b200", "csrc"), os.path.join(EMU, "exact_update_v2_host.cpp"), "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    return C.CDLL(so)


@pytest.mark.parametrize("d_row,d_col,c", [(40, 1024, 256), (33, 768, 0), (64, 1280, 512)])
def test_exact_update_v2_on_the_emulator(lib_upd, d_row, d_col, c):
    """gptq_gguf_toolkit_b200/csrc/exact_update_v2.cuh (GQ_UPDATE_V2=1, experimental): the rank-256 trailing update of the exact
    schedule, W[:, c+256:] <- (W - chain(E[:, c:c+128], U[c:c+128, :])) - chain(E[:, c+128:c+256], U[c+128:c+256, :]) with
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
    lib_upd.run_exact_update_v2(p(got, C.c_float), p(U, C.c_float), C.c_int(d_row), C.c_int(d_col), C.c_int(c))
    assert np.array_equal(got[:, :c + 256], W[:, :c + 256]), "columns up to the finished super-block must not change"
    assert np.array_equal(got, ref)
