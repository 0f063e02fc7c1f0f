"""CPU: the EvoPress layer database emitted straight from the GPTQ results (gptq_gguf_toolkit_b200/ep_database.py,
SURVEY §8f N2) against the REFERENCE's own pipeline run in this container: results -> our .gguf writer ->
`mapper/gguf_splitter.py --gguf-layers --exact` (imported read-only from /root/reference; the test is skipped where the
reference is not mounted, e.g. on the GPU box).  Raw tensor bytes must be identical file for file, the metadata fields
the stitcher / EvoPress read must agree, and the fp16 HF-layout copies must equal gguf-py's dequantisation."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import gguf

from tests.test_host_cpu import _run_driver

REF = "/root/reference/mapper"


def _emit(monkeypatch, tmp_path, qname):
    from gptq_gguf_toolkit_b200.ep_database import emit_database
    from gptq_gguf_toolkit_b200.quant import build_quant_config
    model, q, save_dir = _run_driver(monkeypatch, tmp_path, f"r_{qname}", quant_config=build_quant_config(qname, None))
    out_root = str(tmp_path / "db")
    counts = emit_database(save_dir, model.config, out_root)
    return model, save_dir, out_root, counts


def test_two_levels_side_by_side_and_hf_copies(monkeypatch, tmp_path):
    from gptq_gguf_toolkit_b200.ep_database import EXACT_BITS
    model, save_dir, out_root, counts = _emit(monkeypatch, tmp_path, "Q4_K")
    assert counts == {"gguf": 2 * 7 + 2, "hf": 2 * 7}
    _, save_dir6, _, _ = _emit(monkeypatch, tmp_path, "Q6_K")
    d = os.path.join(out_root, "layers-hf", "model.layers.1.self_attn.k_proj")
    assert sorted(f for f in os.listdir(d) if f.endswith(".pth")) == ["4.5-Q4_K.pth", "6.5625-Q6_K.pth"]
    # EvoPress parses the float before the first '-' (evo_quant_search.py:39-44)
    assert [float(f.split("-")[0]) for f in sorted(os.listdir(d)) if f.endswith(".pth")] == [EXACT_BITS["Q4_K"], EXACT_BITS["Q6_K"]]
    for qname, sd in (("Q4_K", save_dir), ("Q6_K", save_dir6)):
        obj = torch.load(os.path.join(sd, "model.layers.1.self_attn.k_proj", "data.pth"))
        want = gguf.quants.dequantize(obj["packed"].numpy(), getattr(gguf.GGMLQuantizationType, qname)).astype(np.float16)
        got = torch.load(os.path.join(d, f"{EXACT_BITS[qname]}-{qname}.pth"))
        assert got.dtype == torch.float16 and np.array_equal(got.numpy(), want)      # HF row order: no permutation


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
@pytest.mark.parametrize("qname", ["Q4_K", "Q3_K"])
def test_layers_gguf_equal_reference_splitter(monkeypatch, tmp_path, qname):
    from gptq_gguf_toolkit_b200.pack_gptq_into_gguf import write_gguf
    model, save_dir, out_root, _ = _emit(monkeypatch, tmp_path, qname)
    gguf_path = str(tmp_path / "m.gguf")
    write_gguf(model, model.config, save_dir, gguf_path, outtype="f16")
    sys.path.insert(0, REF)
    try:
        from gguf_splitter import GGUFSplitter
    finally:
        sys.path.remove(REF)
    ref_dir = tmp_path / "ref_split"
    GGUFSplitter(gguf_path, str(ref_dir), use_exact_bitwidth=True).split_gguf_model(None)
    ours = os.path.join(out_root, "layers-gguf")
    checked = 0
    for tname in sorted(os.listdir(ours)):
        if not os.path.isdir(os.path.join(ours, tname)):
            continue
        for f in os.listdir(os.path.join(ours, tname)):
            a, b = os.path.join(ours, tname, f), os.path.join(str(ref_dir), tname, f)
            assert os.path.isfile(b), f"reference splitter did not produce {tname}/{f}"
            if f.endswith(".pth"):
                assert open(a, "rb").read() == open(b, "rb").read(), f"{tname}/{f}: raw bytes differ"
                checked += 1
            else:
                ia, ib = json.load(open(a))["tensor_info"], json.load(open(b))["tensor_info"]
                for key in ("name", "type", "quantization", "bitwidth", "exact_bitwidth", "shape", "n_elements", "n_bytes",
                            "data_filename", "np_dtype", "np_shape"):
                    assert ia[key] == ib[key], f"{tname}/{f}: {key}: {ia[key]!r} != {ib[key]!r}"
    assert checked == 2 * 7 + 2
