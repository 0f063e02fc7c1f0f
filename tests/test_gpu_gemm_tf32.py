"""GPU: the two fp32-class tcgen05 GEMM building blocks -- 3xTF32 (csrc/gemm_tf32.cu) and split-fp16 (csrc/gemm_f16x3.cu: the
rank-k update of GQ_MODE_FAST) -- against fp64 matmul.
Tolerance: fp32-class -- max error <= 2e-6 of the largest |sum_k |a||b|| term (plain TF32 / fp16 would be ~5e-4)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tf32x3", "f16x3"])
def gemm(request):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import _lib
    lib = _lib.load()
    f = getattr(lib, f"gq_debug_gemm_{request.param}_nt")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_long, C.c_long, C.c_long, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    w = getattr(lib, f"gq_debug_gemm_{request.param}_workspace")
    w.restype = C.c_size_t
    w.argtypes = [C.c_int] * 4

    def run(A, B, Cm, alpha=1.0, beta=0.0, tile_mode=0, k_mode=0, batch=1, strides=(0, 0, 0), M=None, N=None, K=None):
        M = M or A.shape[-2]; N = N or B.shape[-2]; K = K or A.shape[-1]
        nws = w(M, N, K, batch)
        ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
        rc = f(A.data_ptr(), A.stride(-2), B.data_ptr(), B.stride(-2), Cm.data_ptr(), Cm.stride(-2), M, N, K, batch,
               strides[0], strides[1], strides[2], alpha, beta, tile_mode, k_mode, ws.data_ptr(), nws,
               torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.gq_last_error()
        torch.cuda.synchronize()
    return run


def _check(got, ref, scale):
    err = (got.double() - ref).abs().max().item()
    assert err <= 2e-6 * scale, (err, scale)


@pytest.mark.parametrize("shape", [(128, 128, 128), (256, 384, 96), (1024, 512, 2048), (2048, 2048, 1000), (384, 1280, 512)])
def test_gemm_full(gemm, shape):
    M, N, K = shape
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda") * torch.exp(torch.randn(M, 1, device="cuda"))
    B = torch.randn(N, K, device="cuda")
    Cm = torch.randn(M, N, device="cuda")
    ref = 0.5 * (A.double() @ B.double().T) - 2.0 * Cm.double()
    scale = (A.double().abs() @ B.double().abs().T).max().item()
    gemm(A, B, Cm, alpha=0.5, beta=-2.0)
    _check(Cm, ref, scale)


def test_gemm_strided_views_and_lower_syrk(gemm):
    torch.manual_seed(0)
    big = torch.randn(1024, 1024, device="cuda")
    P = big[256:768, 128:256]                      # (512 x 128) panel inside a larger matrix
    T = torch.randn(512, 512, device="cuda")
    ref = T.double() - P.double() @ P.double().T
    gemm(P, P, T, alpha=-1.0, beta=1.0, tile_mode=1)
    low = torch.tril(torch.ones(4, 4, device="cuda")).repeat_interleave(128, 0).repeat_interleave(128, 1).bool()
    scale = (P.double().abs() @ P.double().abs().T).max().item()
    # all-positive diagonal sums show the tensor core's truncating fp32 accumulation: 5e-6 instead of 2e-6
    assert ((T.double() - ref).abs()[low]).max().item() <= 5e-6 * scale


def test_gemm_triangular_k_ranges_and_batch(gemm):
    torch.manual_seed(1)
    n = 512
    Lt = torch.tril(torch.randn(n, n, device="cuda"))          # lower: A[m][k] = 0 for k > m
    Ut = Lt.T.contiguous()                                     # upper: A[m][k] = 0 for k < m
    X = torch.randn(384, n, device="cuda")
    for A, kmode in ((Ut, 1), (Lt, 2)):
        out = torch.empty(n, 384, device="cuda")
        gemm(A, X, out, k_mode=kmode)
        _check(out, A.double() @ X.double().T, (A.double().abs() @ X.double().abs().T).max().item())
    out = torch.empty(384, n, device="cuda")
    gemm(X, Ut.T.contiguous().T.contiguous(), out, k_mode=0)   # sanity: full
    out = torch.empty(384, n, device="cuda")
    gemm(X, Ut, out, k_mode=3)                                 # B upper triangular: k >= n
    _check(out, X.double() @ Ut.double().T, (X.double().abs() @ Ut.double().abs().T).max().item())
    # batch of 3 independent (128 x 256 x 128) problems laid out back to back
    A = torch.randn(3, 128, 128, device="cuda"); B = torch.randn(3, 256, 128, device="cuda")
    Cb = torch.zeros(3, 128, 256, device="cuda")
    gemm(A, B, Cb, batch=3, strides=(128 * 128, 256 * 128, 128 * 256))
    _check(Cb, A.double() @ B.double().transpose(1, 2), 128 * 16.0)


def test_gemm_wide_dynamic_range_inside_rows(gemm):
    """Elements of one operand row spanning 2^+-12 (the fp16 pair keeps fewer bits of the small ones): the error stays fp32-class
    relative to the row's |a||b| sum, which is what a dot product's fp32 rounding is relative to as well."""
    torch.manual_seed(5)
    M, N, K = 256, 512, 768
    A = torch.randn(M, K, device="cuda") * torch.exp2(torch.randint(-12, 13, (M, K), device="cuda").float()) * 1e-3
    B = torch.randn(N, K, device="cuda") * torch.exp2(torch.randint(-12, 13, (N, K), device="cuda").float()) * 1e3
    Cm = torch.zeros(M, N, device="cuda")
    gemm(A, B, Cm)
    # 5e-6: the tensor core's truncating fp32 accumulation shows on sums dominated by a few huge terms (measured on B200:
    # 4.0e-6 for 3xTF32, below 2e-6 for the fp16 pair)
    err = (Cm.double() - A.double() @ B.double().T).abs().max().item()
    assert err <= 5e-6 * (A.double().abs() @ B.double().abs().T).max().item()


def test_f16x3_b_is_row_prefix_of_a():
    """gemm_f16x3_nt with same_ab and N < M: B = the first N rows of A, split once (the left-looking update of the grouped
    Cholesky in csrc/linalg.cu); the debug hook sets same_ab only for M == N, so this goes through gq_prepare's own check:
    a grouped factorisation (GQ_PREPARE_GROUP=4) must agree with the ungrouped one to fp32 accuracy."""
    import os
    from gptq_gguf_toolkit_b200 import ops
    torch.manual_seed(3)
    n = 1536
    x = torch.randn(2 * n, n, device="cuda")
    H = (x.T @ x) / n
    W = torch.randn(64, n, device="cuda")
    res = {}
    for grp in ("1", "4"):
        os.environ["GQ_PREPARE_GROUP"] = grp
        U, flag = ops.prepare(H.clone(), W, 0.01)
        torch.cuda.synchronize()
        assert int(flag.item()) == 0
        res[grp] = U
    os.environ.pop("GQ_PREPARE_GROUP", None)
    assert float((res["1"] - res["4"]).abs().max() / res["1"].abs().max()) <= 2e-5
