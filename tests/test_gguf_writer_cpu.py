"""CPU: the .gguf writer (gptq_gguf_toolkit_b200/pack_gptq_into_gguf.py) on a tiny random Llama whose
layers were quantised through the driver with the oracle standing in for the kernels (tests/_oracle_backend.py).

Checks, all through gguf-py's OWN reader and dequantiser (an oracle independent of this repository):
  * every quantised module arrives as a K-quant tensor of the right ggml type / shape under llama.cpp's tensor name;
  * gguf.quants.dequantize(bytes) == the dequantised layer weight the driver wrote back (bit-exact), with the Llama
    q/k row permutation of the reference's converter (pack_gptq_into_gguf.py:2177-2183) applied to rows of BYTES;
  * norms stay F32; metadata carries the architecture hyper-parameters;
  * a data.pth WITHOUT `packed` (the reference's own schema) is refused loudly on a machine without CUDA."""
import os

import numpy as np
import pytest
import torch

import gguf

from tests.test_host_cpu import _run_driver


def _read(path):
    r = gguf.GGUFReader(path)
    return r, {t.name: t for t in r.tensors}


@pytest.mark.parametrize("qname", ["Q4_K", "Q6_K", "Q2_K"])
def test_gguf_roundtrip_matches_driver_weights(monkeypatch, tmp_path, qname):
    from gptq_gguf_toolkit_b200.pack_gptq_into_gguf import llama_permute, write_gguf
    from gptq_gguf_toolkit_b200.quant import build_quant_config
    model, q, save_dir = _run_driver(monkeypatch, tmp_path, "w", quant_config=build_quant_config(qname, None))
    out = str(tmp_path / "tiny.gguf")
    written = write_gguf(model, model.config, save_dir, out, outtype="f16")
    assert sum(v == qname for v in written.values()) == 2 * 7 + 2

    reader, tensors = _read(out)
    cfg = model.config
    tmap = gguf.get_tensor_name_map(gguf.MODEL_ARCH.LLAMA, cfg.num_hidden_layers)
    qtype = getattr(gguf.GGMLQuantizationType, qname)
    for hf_name, p in model.state_dict().items():
        t = tensors[tmap.get_name(hf_name, try_suffixes=(".weight",))]
        want = p.detach().float()
        if hf_name.endswith("q_proj.weight"):
            want = llama_permute(want, cfg.num_attention_heads, cfg.num_attention_heads)
        if hf_name.endswith("k_proj.weight"):
            want = llama_permute(want, cfg.num_attention_heads, cfg.num_key_value_heads)
        if hf_name.removesuffix(".weight") in os.listdir(save_dir):
            assert t.tensor_type == qtype, hf_name
            assert [int(x) for x in t.shape] == [p.shape[1], p.shape[0]], hf_name        # ggml order: (ne0 = cols, ne1 = rows)
            deq = gguf.quants.dequantize(np.asarray(t.data), qtype)
            assert np.array_equal(deq, want.numpy()), f"{hf_name}: gguf-py dequantisation differs from the layer weight"
        elif p.dim() == 1:
            assert t.tensor_type == gguf.GGMLQuantizationType.F32
            assert np.array_equal(np.asarray(t.data), want.numpy())
    kv = {f.name: f for f in reader.fields.values()}
    assert int(kv["llama.block_count"].parts[-1][0]) == cfg.num_hidden_layers
    assert int(kv["llama.attention.head_count_kv"].parts[-1][0]) == cfg.num_key_value_heads
    assert int(kv["llama.embedding_length"].parts[-1][0]) == cfg.hidden_size


def test_reference_schema_without_packed_needs_cuda(monkeypatch, tmp_path):
    from gptq_gguf_toolkit_b200._lib import GQError
    from gptq_gguf_toolkit_b200.pack_gptq_into_gguf import write_gguf
    model, q, save_dir = _run_driver(monkeypatch, tmp_path, "w")
    path = os.path.join(save_dir, "model.layers.0.mlp.down_proj", "data.pth")
    d = torch.load(path)
    d.pop("packed")
    torch.save(d, path)
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the GPU packer handles this case")
    with pytest.raises((GQError, RuntimeError, AssertionError)):
        write_gguf(model, model.config, save_dir, str(tmp_path / "x.gguf"))


def test_llama3_rope_scaling_and_unknown_types(monkeypatch, tmp_path):
    """rope_scaling of type llama3 (Llama 3.1 / 3.2) must produce rope_freqs.weight like the reference's generate_extra_tensors
    (pack_gptq_into_gguf.py:2259-2287); linear scaling the two KVs (:2172-2175); unknown types are refused."""
    from gptq_gguf_toolkit_b200.pack_gptq_into_gguf import llama3_rope_factors, write_gguf
    model, q, save_dir = _run_driver(monkeypatch, tmp_path, "w")
    cfg = model.config.to_dict()
    cfg.pop("rope_parameters", None)
    cfg["rope_theta"] = 500000.0
    cfg["rope_scaling"] = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                           "original_max_position_embeddings": 64}
    out = str(tmp_path / "l3.gguf")
    written = write_gguf(model, cfg, save_dir, out)
    assert written["rope_freqs.weight"] == "F32"
    reader, tensors = _read(out)
    f = np.asarray(tensors["rope_freqs.weight"].data)
    want = llama3_rope_factors(cfg, cfg["rope_scaling"]).numpy()
    assert f.shape == (cfg["hidden_size"] // cfg["num_attention_heads"] // 2,) and np.array_equal(f, want)
    assert f[0] == 1.0 and f[-1] == 8.0 and np.all(np.diff(f) >= 0)          # high frequencies untouched, low ones stretched
    # the published rule, restated independently for one mid-band frequency
    dim, base = cfg["hidden_size"] // cfg["num_attention_heads"], 500000.0
    mid = [i for i in range(dim // 2) if 1.0 < f[i] < 8.0]
    for i in mid[:2]:
        wl = 2 * np.pi * base ** (2 * i / dim)
        smooth = (64 / wl - 1.0) / (4.0 - 1.0)
        assert abs(f[i] - 1 / ((1 - smooth) / 8.0 + smooth)) < 1e-5
    cfg["rope_scaling"] = {"rope_type": "linear", "factor": 2.0}
    write_gguf(model, cfg, save_dir, str(tmp_path / "lin.gguf"))
    r2, _ = _read(str(tmp_path / "lin.gguf"))
    kv = {x.name: x for x in r2.fields.values()}
    assert float(kv["llama.rope.scaling.factor"].parts[-1][0]) == 2.0
    cfg["rope_scaling"] = {"rope_type": "longrope", "factor": 2.0}
    with pytest.raises(NotImplementedError):
        write_gguf(model, cfg, save_dir, str(tmp_path / "bad.gguf"))
