"""GPU: the artefacts downstream stages consume, produced from libgq results (the CPU twins run the same
host code on the oracle backend):

  * quant.py's driver on a tiny Llama with the real kernels -> data.pth files -> `.gguf` (pack_gptq_into_gguf.py:282-349 of the
    reference, here gptq_gguf_toolkit_b200/pack_gptq_into_gguf.py): read back with gguf-py's own reader, its dequantiser must
    reproduce the dequantised weights the kernels wrote into the model, bit for bit (q/k rows permuted like llama.cpp does);
  * the EvoPress layer database (mapper/gguf_splitter.py:373-404, here ep_database.py): layers-gguf/<tensor>/<bw>-<Q>.pth must
    hold exactly the tensor's bytes in that .gguf (what the reference's splitter extracts), layers-hf the fp16 HF-order weights.
"""
import os

import numpy as np
import pytest
import torch

import gguf

pytestmark = pytest.mark.gpu
REGEX = r".*layers.*((q|k|v|o|gate|up|down)_proj)$"


def _quantise(tmp_path, qname, dtype):
    from transformers import LlamaConfig, LlamaForCausalLM
    from gptq_gguf_toolkit_b200.quant import build_quant_config
    from gptq_gguf_toolkit_b200.quantizer import Quantizer
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, tie_word_embeddings=False)
    model = LlamaForCausalLM(cfg).to("cuda", dtype).eval()
    g = torch.Generator().manual_seed(1)
    loader = [([], {"input_ids": torch.randint(0, 512, (1, 96), generator=g).cuda()}) for _ in range(8)]
    save_dir = str(tmp_path / f"quant_{qname}")
    q = Quantizer(model, data_loader=loader, quantizable_modules=REGEX,
                  quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax", static_groups=False,
                                        rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                  pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                  quant_non_block_modules=True, device=torch.device("cuda"), save_dir=save_dir, calibration_batch_size=4)
    q.quantize(build_quant_config(qname, None))
    torch.cuda.synchronize()
    return model, cfg, save_dir


@pytest.mark.parametrize("qname,dtype", [("Q4_K", torch.bfloat16), ("Q6_K", torch.float16), ("Q3_K", torch.float32)])
def test_gguf_and_layer_database_from_libgq_results(tmp_path, qname, dtype):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200.ep_database import EXACT_BITS, emit_database
    from gptq_gguf_toolkit_b200.pack_gptq_into_gguf import llama_permute, write_gguf
    model, cfg, save_dir = _quantise(tmp_path, qname, dtype)
    assert len(os.listdir(save_dir)) == 2 * 7 + 2
    out = str(tmp_path / "tiny.gguf")
    written = write_gguf(model, cfg, save_dir, out, outtype="f16")
    assert sum(v == qname for v in written.values()) == 2 * 7 + 2
    reader = gguf.GGUFReader(out)
    tensors = {t.name: t for t in reader.tensors}
    tmap = gguf.get_tensor_name_map(gguf.MODEL_ARCH.LLAMA, cfg.num_hidden_layers)
    qtype = getattr(gguf.GGMLQuantizationType, qname)
    n_checked = 0
    for hf_name, p in model.state_dict().items():
        base = hf_name.removesuffix(".weight")
        if base not in os.listdir(save_dir):
            continue
        t = tensors[tmap.get_name(hf_name, try_suffixes=(".weight",))]
        want = p.detach().float().cpu()
        if hf_name.endswith("q_proj.weight"):
            want = llama_permute(want, cfg.num_attention_heads, cfg.num_attention_heads)
        if hf_name.endswith("k_proj.weight"):
            want = llama_permute(want, cfg.num_attention_heads, cfg.num_key_value_heads)
        assert t.tensor_type == qtype, hf_name
        deq = gguf.quants.dequantize(np.asarray(t.data), qtype)          # gguf-py's own dequantiser on the file's bytes
        got = torch.from_numpy(deq).to(dtype).float()                    # the layer holds the dequantised weight in the model dtype
        assert torch.equal(got, want), f"{hf_name}: gguf-py dequantisation differs from the weight the kernels wrote back"
        n_checked += 1
    assert n_checked == 2 * 7 + 2

    db = str(tmp_path / "db")
    counts = emit_database(save_dir, cfg, db)
    assert counts == {"gguf": 2 * 7 + 2, "hf": 2 * 7}
    prefix = f"{EXACT_BITS[qname] if EXACT_BITS[qname] != int(EXACT_BITS[qname]) else int(EXACT_BITS[qname])}-{qname}"
    for name, t in tensors.items():
        f = os.path.join(db, "layers-gguf", name, f"{prefix}.pth")
        if t.tensor_type != qtype:
            assert not os.path.exists(f)
            continue
        assert open(f, "rb").read() == np.asarray(t.data).tobytes(), f"{name}: database bytes differ from the .gguf tensor"
    w = torch.load(os.path.join(db, "layers-hf", "model.layers.1.mlp.down_proj", f"{prefix}.pth"))
    t = tensors[tmap.get_name("model.layers.1.mlp.down_proj.weight", try_suffixes=(".weight",))]      # down_proj: no row permutation
    want = torch.from_numpy(gguf.quants.dequantize(np.asarray(t.data), qtype)).to(torch.float16)
    assert w.dtype == torch.float16 and torch.equal(w, want)
