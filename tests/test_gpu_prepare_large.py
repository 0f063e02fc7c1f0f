"""GPU: accuracy of gq_prepare (csrc/linalg.cu) at the sizes the Llama-3-8B benchmark runs -- n = 4096 (q/k/v/o, gate/up) and
n = 14336 (down_proj) -- against an fp64 factorisation on the device (reference: gptq.py:305-324 + linalg_utils.py:8-12,
U = chol(inv(H + damp I), upper)).  Both tcgen05 GEMM back-ends of the chain (GQ_PREPARE_GEMM = f16 (default) | tf32) and the
grouped / ungrouped trailing updates (GQ_PREPARE_GROUP) must meet the bounds.

Bounds (fp32 factorisation of a matrix with cond ~ 1e3..1e4, the B2 class of DESIGN.md section 2):
  * max |U - U64| <= 2e-4 max|U64|           (same bound as the small-n test in test_gpu_parity.py)
  * max |U H U^T - I| <= 2e-3                (H^-1 = U^T U  <=>  U H U^T = I)
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import ops as o
    return o


def _problem(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mix = torch.randn(n, n, device="cuda", generator=g) / n ** 0.5
    H = torch.zeros(n, n, device="cuda")
    T = 4096
    from gptq_gguf_toolkit_b200 import ops as o
    nb = (2 * n + T - 1) // T
    for b in range(nb):      # H = 2/N sum x x^T of correlated activations, accumulated like GPTQ.update does
        X = (torch.randn(T, n, device="cuda", generator=g) @ mix).to(torch.bfloat16)
        o.hessian_update(H, X, b / (b + 1.0), 2.0 / (b + 1.0) / T)
    W = torch.randn(64, n, device="cuda", generator=g) * 0.05
    W[:, 11] = 0.0
    return H, W


@pytest.mark.parametrize("backend,group", [("f16", "4"), ("f16", "1"), ("tf32", "4")])
@pytest.mark.parametrize("n", [4096, 14336])
def test_prepare_accuracy_at_llama_sizes(ops, n, backend, group, monkeypatch):
    monkeypatch.setenv("GQ_PREPARE_GEMM", backend)
    monkeypatch.setenv("GQ_PREPARE_GROUP", group)      # block columns per trailing update of the Cholesky (csrc/linalg.cu)
    H, W = _problem(n, n)
    Hd = H.clone()
    U, flag = ops.prepare(Hd, W, 0.01)
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    assert bool((torch.tril(U, -1) == 0).all()), "U must be exactly upper triangular"
    H64 = Hd.double()                                   # masked + damped in place, like the reference
    U64 = torch.linalg.cholesky(torch.linalg.inv(H64), upper=True)
    rel = float((U.double() - U64).abs().max() / U64.abs().max())
    R = U.double() @ H64 @ U.double().T
    R.diagonal().sub_(1.0)
    resid = float(R.abs().max())
    print(f"gq_prepare n={n} backend={backend} group={group}: max|U-U64|/max|U64| = {rel:.2e}, max|U H U^T - I| = {resid:.2e}")
    assert rel <= 2e-4, rel
    assert resid <= 2e-3, resid
