"""GPU: the block-sequential driver (gptq_gguf_toolkit_b200/quantizer.py) on a tiny random Llama, through libgq.

The scheduling options added over the reference -- pass-1 early exit, Cholesky chains on side streams (eager / staged;
started from the forward hook of the last calibration batch or after pass 1),
the deferred-tail split of pass 2 (last quantised layer's chain + column loop overlapped with the block forwards) --
must not change a single bit of any result: same kernels, same inputs, only the order of independent work differs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

REGEX = r".*layers.*((q|k|v|o|gate|up|down)_proj)$"


def _model(dtype):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=768, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).to("cuda", dtype).eval()


def _run(dtype, qname="Q4_K", **kw):
    from gptq_gguf_toolkit_b200.quant import build_quant_config
    from gptq_gguf_toolkit_b200.quantizer import Quantizer
    model = _model(dtype)
    g = torch.Generator().manual_seed(1)
    loader = [([], {"input_ids": torch.randint(0, 512, (1, 64), generator=g).cuda()}) for _ in range(8)]
    q = Quantizer(model, data_loader=loader, quantizable_modules=REGEX,
                  quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                        static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                  pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                  quant_non_block_modules=True, device=torch.device("cuda"), save_dir=None, keep_results=True,
                  calibration_batch_size=4, **kw)
    q.quantize(build_quant_config(qname, None))
    torch.cuda.synchronize()
    assert q.non_invertible_modules() == []
    return model, q


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_scheduling_options_are_bit_neutral(dtype):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    base_model, base = _run(dtype, overlap_prepare=False, defer_last_layer=False, early_exit_pass1=False)
    assert len(base.results) == 3 * 7 + 2
    variants = {
        "default (eager + deferred tail + early exit)": {},
        "staged": dict(overlap_prepare="staged"),
        "eager, no deferred tail": dict(defer_last_layer=False),
        "no early exit": dict(early_exit_pass1=False),
        "no early prepare (chains start after pass 1)": dict(early_prepare=False),
        "column loops one after the other on the main stream": dict(concurrent_groups=False),
        "fast mode scratch slots (concurrent groups)": dict(concurrent_groups=True, defer_last_layer=False),
    }
    for name, kw in variants.items():
        model, q = _run(dtype, **kw)
        if "deferred tail" in name and "no deferred" not in name:
            assert q._split_ok is True, "the deferred-tail split must be recognised as valid for a Llama block"
        assert q.results.keys() == base.results.keys()
        for mod, ref in base.results.items():
            got = q.results[mod]
            for key, t in ref.items():
                if isinstance(t, torch.Tensor):
                    assert torch.equal(t, got[key]), f"{name}: {mod}.{key} differs"
        for (n1, p1), (n2, p2) in zip(base_model.named_parameters(), model.named_parameters()):
            assert torch.equal(p1, p2), f"{name}: weight {n1} differs after quantisation"


def test_deferred_tail_mixed_types():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _, a = _run(torch.bfloat16, qname="Q6_K", overlap_prepare=False, defer_last_layer=False)
    _, b = _run(torch.bfloat16, qname="Q6_K")
    for mod, ref in a.results.items():
        for key, t in ref.items():
            if isinstance(t, torch.Tensor):
                assert torch.equal(t, b.results[mod][key]), f"{mod}.{key}"
