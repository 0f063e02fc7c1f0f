"""GPU: the fused element-wise pieces of the block forwards (csrc/fwd_ops.cu, fused_forward.py) against the HF modules
they replace: SiLU*up and the rotary embedding bit-identical, RMSNorm within one 16-bit ulp; the context manager
installs, verifies and restores; a whole Llama forward with the pieces installed stays within bf16 noise of HF's."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(dtype):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=512, hidden_size=512, intermediate_size=1536, num_hidden_layers=2, num_attention_heads=8,
                      num_key_value_heads=2, max_position_embeddings=256, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).to("cuda", dtype).eval()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_pieces_match_hf(dtype):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from transformers.models.llama import modeling_llama as ml
    from gptq_gguf_toolkit_b200 import fused_forward as ff
    torch.manual_seed(1)
    model = _model(dtype)
    blk = model.model.layers[0]
    # SiLU * up: bit-identical
    g = torch.randn(64, 1536, device="cuda", dtype=dtype) * 4
    u = torch.randn(64, 1536, device="cuda", dtype=dtype)
    assert torch.equal(ff.silu_mul(g, u), blk.mlp.act_fn(g) * u)
    # rotary embedding on the strided views the attention module produces: bit-identical, same strides
    B, L, H, Hk, hd = 3, 40, 8, 2, 64
    q = torch.randn(B, L, H * hd, device="cuda", dtype=dtype).view(B, L, H, hd).transpose(1, 2)
    k = torch.randn(B, L, Hk * hd, device="cuda", dtype=dtype).view(B, L, Hk, hd).transpose(1, 2)
    pos = torch.arange(L, device="cuda")[None]
    cos, sin = model.model.rotary_emb(q, pos)
    wq, wk = ml.apply_rotary_pos_emb(q, k, cos, sin)
    gq, gk = ff.rope(q, cos, sin), ff.rope(k, cos, sin)
    assert torch.equal(gq, wq) and torch.equal(gk, wk) and gq.stride() == wq.stride()
    # RMSNorm: within one ulp of the 16-bit result (fp32 mean summed in another order)
    x = torch.randn(5, 33, 512, device="cuda", dtype=dtype) * 2.5
    blk.input_layernorm.weight.data = (1 + 0.1 * torch.randn(512, device="cuda")).to(dtype)
    want = blk.input_layernorm(x)
    got = ff.rmsnorm(x, blk.input_layernorm.weight, blk.input_layernorm.variance_epsilon)
    ulp = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    assert ((got.float() - want.float()).abs() <= ulp * want.float().abs()).all()
    assert (got == want).float().mean() > 0.99


def test_context_manager_installs_and_restores():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from transformers.models.llama import modeling_llama as ml
    from gptq_gguf_toolkit_b200.fused_forward import fused_forward
    model = _model(torch.bfloat16)
    ids = torch.randint(0, 512, (2, 64), device="cuda")
    orig_rope = ml.apply_rotary_pos_emb
    with torch.no_grad():
        ref = model(ids).logits
        with fused_forward(model) as installed:
            assert any("RMSNorm" in s for s in installed) and any("MLP" in s for s in installed)
            assert any("apply_rotary_pos_emb" in s for s in installed)
            assert ml.apply_rotary_pos_emb is not orig_rope
            got = model(ids).logits
        assert ml.apply_rotary_pos_emb is orig_rope
        assert all("forward" not in m.__dict__ for m in model.modules())
        again = model(ids).logits
    assert torch.equal(again, ref)
    rel = (got.float() - ref.float()).abs().max() / ref.float().abs().max()
    assert rel < 2e-2, rel                      # bf16 noise (RMSNorm's last-bit differences propagate through 2 blocks)
    # fp32 models are left alone
    with fused_forward(model.float()) as installed:
        assert installed == []
