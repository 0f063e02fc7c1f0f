#!/usr/bin/env python
"""Build-container check (needs /root/reference): the oracle against the UNMODIFIED reference at layer sizes beyond the
committed goldens, and the speed of the oracle PORT against the reference's own CPU implementation (what `cpu_baseline.kind =
"port"` in bench.py stands in for).

    python tests/golden/check_oracle_vs_reference_large.py

Recorded on 2026-10-17 (8 vCPU, torch 2.11, 8 threads, Q4_K, reference with an IEEE sqrt as in make_golden.py):
    1024 x 2048: oracle == reference bit for bit (codes and scales); reference step 1.20 s, oracle port 0.29 s
    2048 x 4096: oracle == reference bit for bit (codes and scales); reference step 3.51 s, oracle port 1.61 s
    unpatched torch.sqrt (1 ulp off IEEE on ~0.7 % of the inputs): 99.8 % identical rows at 1024 x 2048
so the port under-estimates the reference's CPU time for the column loop by 2.2-4.1x: the CPU baseline bench.py reports is
conservative."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/reference/quant/gptq")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from src.gptq import GPTQ  # noqa: E402
from src import quant_utils as qu  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    torch.set_num_threads(8)
    orig = torch.sqrt
    for ieee in (True, False):
        torch.sqrt = (lambda t: t.double().sqrt().float()) if ieee else orig
        for d_row, d_col in ((1024, 2048), (2048, 4096)):
            torch.manual_seed(0)
            layer = torch.nn.Linear(d_col, d_row, bias=False)
            h = GPTQ(layer, rel_damp=0.01, block_size=128)
            for _ in range(4):
                h.update(torch.randn(1, 2048, d_col).to(torch.bfloat16).float())
            h.quantization_pre_step()
            U = h._prepare().clone()
            W0 = h.W.clone()
            h._prepare = lambda: U
            t0 = time.perf_counter()
            out = h.step(qu.GGMLQuantizationType.Q4_K)
            t_ref = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref = orc.gptq_step(W0.numpy().copy(), np.ascontiguousarray(U.numpy()), 12)
            t_port = time.perf_counter() - t0
            q = out[0].numpy()
            print(f"ieee_sqrt={ieee} {d_row}x{d_col}: rows identical {(q == ref[0]).all(1).mean():.4f}, scales identical "
                  f"{np.array_equal(out[1].numpy().view(np.uint16), ref[1].view(np.uint16))}; reference step {t_ref:.2f} s, "
                  f"oracle port {t_port:.2f} s")


if __name__ == "__main__":
    main()
