#!/usr/bin/env python
"""Golden vectors for GPTQ.step with --block_size other than 128 (gptq.py:55,219-270 of the reference): the UNMODIFIED reference on
CPU (IEEE sqrt, see make_golden.py) on one seeded layer, for block_size 32, 64, 256 and all five types.
    python tests/golden/make_golden_blocksize.py        (build container only: needs /root/reference)
Stored per (block_size, type): the five outputs, the GGUF bytes and the dequantised weights; W and U once (U does not depend on
the block size)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts the reference on sys.path)


def main():
    d_row, d_col, seed = 20, 768, 11
    gen = torch.Generator().manual_seed(seed)
    W = torch.randn(d_row, d_col, generator=gen) * 0.05
    W = W * torch.exp(0.5 * torch.randn(d_row, 1, generator=gen))
    xs = mg.correlated_x(gen, 6, 96, d_col)
    out = {"W": W.numpy()}
    U0 = None
    for bs in (32, 64, 256):
        for q_type in mg.TYPES:
            mg.set_sqrt(True)
            five, U, H = mg.run_ref_gptq(W, xs, q_type, bs)
            mg.set_sqrt(False)
            if U0 is None:
                U0 = U.clone()
                out["U_colmajor_T"] = U.t().contiguous().numpy()
            assert torch.equal(U, U0)
            for k, v in mg.np5(five).items():
                out[f"bs{bs}_{q_type.name}_{k}"] = v
            out[f"bs{bs}_{q_type.name}_packed"] = np.asarray(mg.pack_ref(q_type, five))
            out[f"bs{bs}_{q_type.name}_dequant"] = mg.qu.dequantize_linear_weight(q_type, *five).numpy()
    np.savez_compressed(os.path.join(HERE, "blocksize_a.npz"), **out)
    print("wrote blocksize_a.npz", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
