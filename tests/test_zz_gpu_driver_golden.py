"""GPU: BASELINE.json configs[0] end to end through libgq against the UNMODIFIED reference driver's output
(tests/golden/driver_tiny.npz, see tests/golden/make_golden_driver.py and tests/test_driver_golden_cpu.py, which runs the
same comparison with the CPU oracle behind the host logic).  Boundary B3: embed_tokens / lm_head (RTN) bit-exact, the GPTQ
layers statistical -- Hessian GEMM and Cholesky chain differ from MKL / LAPACK in rounding, and the reference does not
reproduce itself across thread counts there either.  With the CPU oracle the same run gives 100 % identical codes on six of
the seven layers of block 0 and >= 85.8 % everywhere; the thresholds below are deliberately wider for the CUDA
factorisation (3xTF32 GEMMs inside gq_prepare)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.test_driver_golden_cpu import CFG, KEYS, REGEX, _deq, _raw

pytestmark = pytest.mark.gpu


def test_cuda_driver_against_reference_driver_golden(golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from transformers import LlamaConfig, LlamaForCausalLM
    from gptq_gguf_toolkit_b200.quant_utils import GGMLQuantizationType as T
    from gptq_gguf_toolkit_b200.quantizer import Quantizer
    g = np.load(os.path.join(golden_dir, "driver_tiny.npz"))
    golden = {str(n): {k: g[f"{n}|{k}"] for k in KEYS} for n in g["names"]}
    torch.manual_seed(0)
    model = LlamaForCausalLM(LlamaConfig(**CFG)).float().eval()
    pristine = {n: p.data.clone() for n, p in model.named_parameters()}
    model = model.cuda()
    gen = torch.Generator().manual_seed(1)
    loader = [([], {"input_ids": torch.randint(0, CFG["vocab_size"], (1, 128), generator=gen).cuda()}) for _ in range(8)]
    quant_config = {k: T.Q4_K for k in ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj",
                                        "embed_tokens", "lm_head")}
    q = Quantizer(model, data_loader=loader, quantizable_modules=REGEX,
                  quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                        static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                  pre_block_modules=["model.embed_tokens"], post_block_modules=["lm_head"], block_modules="model.layers",
                  save_dir=None, quant_non_block_modules=True, device=torch.device("cuda"), keep_results=True,
                  calibration_batch_size=4)
    q.quantize(quant_config)
    torch.cuda.synchronize()
    assert q.non_invertible_modules() == []
    res = {n: {k: _raw(d[k].numpy()) for k in KEYS} for n, d in q.results.items()}
    assert sorted(res) == sorted(golden)
    for n in ("model.embed_tokens", "lm_head"):
        for k in KEYS:
            assert np.array_equal(res[n][k], golden[n][k]), f"{n}.{k}: RTN of the pristine weights must be bit-exact"
    for n, ref in golden.items():
        if "layers" not in n:
            continue
        same = float((res[n]["qweight"] == ref["qweight"]).mean())
        wr, wo = _deq(ref), _deq(res[n])
        w0 = pristine[n + ".weight"].numpy()
        er, eo = float(np.linalg.norm(wr - w0)), float(np.linalg.norm(wo - w0))
        print(f"{n}: identical codes {same:.4f}, quant. error ours/ref {eo / er:.4f}")
        assert same >= (0.95 if ".layers.0." in n else 0.6), (n, same)
        assert abs(eo / er - 1.0) <= 0.02, (n, eo / er)
        # the layer weight left in the model is the dequantisation of what was emitted (quantizer.py:257-264)
        assert np.array_equal(model.get_submodule(n).weight.data.cpu().numpy(), wo), n
