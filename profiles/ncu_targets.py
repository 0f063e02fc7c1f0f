#!/usr/bin/env python
"""Small stand-alone launches of the hot kernels for `ncu --set full` captures (profiles/profile.sh):
    python profiles/ncu_targets.py gptq      # o_proj shape (4096 x 4096, Q4_K): 16 panel launches + 15 exact_update launches
    python profiles/ncu_targets.py prepare   # one Cholesky chain at n = 4096 (chol_diag_v2_kernel, gemm_tf32x3_kernel)
    python profiles/ncu_targets.py hessian   # one tcgen05 SYRK update, d_col 4096, 16384 tokens
    python profiles/ncu_targets.py fast      # down_proj shape (4096 x 14336, Q4_K) in GQ_MODE_FAST: gemm_f16x3_kernel launches
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gptq_gguf_toolkit_b200 import ops  # noqa: E402

what = set(sys.argv[1:]) or {"gptq", "prepare", "hessian"}      # "fast" only on request
torch.manual_seed(0)
n = 4096
if "gptq" in what:
    U = torch.triu(torch.randn(n, n, device="cuda") * 0.01) + torch.eye(n, device="cuda")
    W = torch.randn(4096, n, device="cuda") * 0.02
    ops.gptq_quantize(W, U, 12, wdeq_dtype=torch.bfloat16)
if "fast" in what:
    n2 = 14336
    U = torch.triu(torch.randn(n2, n2, device="cuda") * 0.01) + torch.eye(n2, device="cuda")
    W = torch.randn(4096, n2, device="cuda") * 0.02
    ops.gptq_quantize(W, U, 12, wdeq_dtype=torch.bfloat16, mode=1)
    del U, W
if "prepare" in what:
    x = torch.randn(2 * n, n, device="cuda")
    H = (x.T @ x) / n
    ops.prepare(H, torch.randn(256, n, device="cuda"), 0.01)
if "prepare_big" in what:      # one Cholesky chain at down_proj's width (813 launches)
    nb = 14336
    x = torch.randn(nb // 2, nb, device="cuda")
    H = (x.T @ x) / nb + 0.5 * torch.eye(nb, device="cuda")
    del x
    ops.prepare(H, torch.randn(256, nb, device="cuda"), 0.01)
if "hessian" in what:
    X = torch.randn(16384, n, device="cuda").to(torch.bfloat16)
    H = torch.zeros(n, n, device="cuda")
    ops.hessian_update(H, X, 0.0, 2.0)
    ops.hessian_update(H, X, 0.5, 1.0)
torch.cuda.synchronize()
