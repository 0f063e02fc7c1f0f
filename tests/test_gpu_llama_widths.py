"""GPU parity at the widths the Llama-3-8B benchmark runs (d_col = 4096 and 14336), through the C ABI:

  * all five K-quant types against the oracle, bit for bit (codes, the four scale tensors, GGUF bytes, dequantised weights),
    on row slabs of full-width layers -- rows of a GPTQ problem are independent given U (gptq.py:146-295), so a slab checks
    the very arithmetic every row of the full layer goes through, in seconds of CPU time;
  * BASELINE.json configs[2]: a per-projection MIXED bit-width configuration (quant.py:203-217) through the whole driver on one
    Llama-3-8B-shaped block (hidden 4096, intermediate 14336, 32 / 8 heads); every column-loop launch the driver makes is
    recorded as (W, U, q_type) and replayed through the oracle, and the bytes the driver emitted must be identical.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.test_gpu_parity import TYPES, assert_five_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import ops as o
    return o


def _problem(rows, d_col, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    W = torch.randn(rows, d_col, device="cuda", generator=g) * 0.02
    U = torch.triu(torch.randn(d_col, d_col, device="cuda", generator=g) * (0.3 / d_col ** 0.5))
    U.diagonal().copy_(1.0 + 0.1 * torch.rand(d_col, device="cuda", generator=g))
    return W, U


def _check_against_oracle(out, W_np, U_np, qt, what):
    ref = orc.gptq_step(W_np, U_np, qt)
    assert_five_equal(out[:5], ref[:5], what)
    assert np.array_equal(out[5].cpu().numpy(), orc.pack(qt, *ref[:5])), f"{what}: GGUF bytes"
    assert torch.equal(out[6].cpu(), torch.from_numpy(ref[5]).to(out[6].dtype)), f"{what}: dequantised weights"


@pytest.mark.parametrize("tname", list(TYPES))
@pytest.mark.parametrize("rows,d_col", [(64, 4096), (32, 14336)])
def test_all_types_bit_exact_at_llama_widths(ops, tname, rows, d_col):
    W, U = _problem(rows, d_col, rows + d_col + TYPES[tname])
    W_np, U_np = W.cpu().numpy(), U.cpu().numpy()
    out = ops.gptq_quantize(W.clone(), U, TYPES[tname], wdeq_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    _check_against_oracle(out, W_np, U_np, TYPES[tname], f"{tname} {rows}x{d_col}")


def test_mixed_configuration_llama_block_replayed_through_oracle(ops, monkeypatch):
    from transformers import LlamaConfig, LlamaForCausalLM
    from gptq_gguf_toolkit_b200 import quantizer as Q
    from gptq_gguf_toolkit_b200.quant_utils import GGMLQuantizationType as T

    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=2048, hidden_size=4096, intermediate_size=14336, num_hidden_layers=1, num_attention_heads=32,
                      num_key_value_heads=8, max_position_embeddings=512, tie_word_embeddings=False)
    with torch.device("cuda"):
        model = LlamaForCausalLM(cfg).to(torch.bfloat16).eval()
    g = torch.Generator().manual_seed(1)
    loader = [([], {"input_ids": torch.randint(0, 2048, (1, 256), generator=g).cuda()}) for _ in range(8)]
    quant_config = {"q_proj": T.Q3_K, "k_proj": T.Q2_K, "v_proj": T.Q5_K, "o_proj": T.Q4_K, "gate_proj": T.Q3_K, "up_proj": T.Q4_K,
                    "down_proj": T.Q6_K}
    ROWS = 48
    calls = []
    real = Q.ops.gptq_quantize

    def recording(W, U, q_type, *a, **kw):
        rec = {"W": W[:ROWS].clone(), "U": U, "qt": int(q_type), "rows": W.shape[0]}
        out = real(W, U, q_type, *a, **kw)
        rec["out"] = out
        calls.append(rec)
        return out

    monkeypatch.setattr(Q.ops, "gptq_quantize", recording)
    q = Q.Quantizer(model, data_loader=loader, quantizable_modules=r".*layers.*((q|k|v|o|gate|up|down)_proj)$",
                    quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax", static_groups=False,
                                          rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                    pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                    quant_non_block_modules=False, device=torch.device("cuda"), save_dir=None, keep_results=True,
                    calibration_batch_size=4)
    q.quantize(quant_config)
    torch.cuda.synchronize()
    assert not q.non_invertible_modules()
    # one launch per (shared input, q_type): q / k / v have three different types here, gate / up two
    assert sorted(c["qt"] for c in calls) == sorted(int(v) for v in quant_config.values())
    shapes = {10: (1024, 4096), 11: None, 12: None, 13: (1024, 4096), 14: (4096, 14336)}
    for c in calls:
        if shapes[c["qt"]] is not None:
            assert (c["rows"], c["U"].shape[0]) == shapes[c["qt"]]
        out = [t[:ROWS] for t in c["out"][:7]]
        _check_against_oracle(out, c["W"].cpu().numpy(), c["U"].cpu().numpy(), c["qt"], f"replay q_type {c['qt']} d_col {c['U'].shape[0]}")
    # and what the driver emitted per module is what those launches produced
    want = {"q_proj": 11, "k_proj": 10, "v_proj": 13, "o_proj": 12, "gate_proj": 11, "up_proj": 12, "down_proj": 14}
    for name, d in q.results.items():
        assert d["q_type"] == want[name.split(".")[-1]], name
        layer = model.get_submodule(name)
        deq = ops.dequantize(d["q_type"], d["qweight"].cuda(), d["super_group_scale"].cuda(), d["group_scale_quant"].cuda(),
                             d["super_group_zero"].cuda(), d["group_zero_quant"].cuda(), torch.bfloat16)
        assert torch.equal(deq, layer.weight.data), f"{name}: layer weight must be the dequantised result (quantizer.py:257-264)"
        assert torch.equal(ops.pack(d["q_type"], d["qweight"].cuda(), d["super_group_scale"].cuda(), d["group_scale_quant"].cuda(),
                                    d["super_group_zero"].cuda(), d["group_zero_quant"].cuda()).cpu(), d["packed"])
