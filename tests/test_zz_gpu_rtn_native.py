"""GPU: gq_rtn_quantize_native -- embed_tokens / lm_head of a 16-bit model with the scale search in the weight's own arithmetic,
as the reference computes it (quant/gptq/src/quantizer.py:303-305 passes the weight un-widened, so every torch op of
get_scale_and_zero rounds to bf16 / fp16) -- against the reference's own output (tests/golden/rtn_bf16.npz, rtn_f16.npz, made by
tests/golden/make_golden_rtn_bf16.py).  The kernel was written after round 1's GPU budget was spent; before this first run on
hardware its whole body was checked on the SIMT emulator of the CPU suite against the same goldens (tests/test_simt_emu_cpu.py)
and by 1034 random cases against the oracle.  The file name makes it run after the other GPU tests."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.test_gpu_parity import KEYS, TYPES, assert_five_equal, raw

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("tname", list(TYPES))
def test_rtn_native_arithmetic_matches_reference_golden(golden_dir, tname, dtype):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import ops
    g = np.load(os.path.join(golden_dir, f"rtn_{dtype}.npz"))
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    W = torch.from_numpy(g[f"W_{dtype}_bits"].view(np.int16).copy()).view(tdt).cuda()
    out = ops.rtn_quantize(W, TYPES[tname], wdeq_dtype=tdt, native_arith=True)
    torch.cuda.synchronize()
    ref5 = [g[f"{tname}_{k}"] for k in KEYS]
    assert_five_equal(out[:5], ref5, f"rtn native {dtype}/{tname}")
    cd = np.uint8 if tname in ("Q2_K", "Q4_K", "Q5_K") else np.int8
    five = (ref5[0].view(cd), ref5[1].view(np.float16), ref5[2].view(cd), ref5[3].view(np.float16), ref5[4].view(cd))
    assert np.array_equal(raw(out[5]), orc.pack(TYPES[tname], *five)), "GGUF block bytes"
    assert torch.equal(out[6].cpu(), torch.from_numpy(orc.dequantize(TYPES[tname], *five)).to(tdt)), "dequantised weights"
