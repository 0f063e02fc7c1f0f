"""EvoPress layer database straight from the GPTQ results.

The reference builds its per-layer database by writing a .gguf per quantisation level and splitting it again
(`mapper/build_ep_database.sh:127-175` -> `mapper/gguf_splitter.py --gguf-layers --hf-layers --exact`):

    layers-gguf/<gguf tensor name>/<bw>-<Q>.pth            raw GGUF block bytes of the tensor     (gguf_splitter.py:373-375)
    layers-gguf/<gguf tensor name>/<bw>-<Q>-metadata.json  tensor_info                           (:381-404)
    layers-hf/<hf module name>/<bw>-<Q>.pth                torch.save(fp16 dequantised weight)    (:556-572)
    layers-hf/<hf module name>/<bw>-<Q>-metadata.json

with <bw> the exact fractional bit width of the K-quant type (`get_tensor_bit_width`, :52-93) -- EvoPress only
parses the float in front of the first '-' (evopress/evo_quant_search.py:39-44) and `gguf_stitcher.py` copies the raw
bytes back into a .gguf.  The data.pth files of our driver already contain those bytes (`packed`), so the database
is emitted directly, one call per quantisation level, byte-compatible with the splitter's output
(tests/test_ep_database_cpu.py compares with the reference splitter run on the .gguf of the same results).
Host-side only: file formats, no arithmetic except gguf-py's dequantiser for the fp16 copies.
"""
from __future__ import annotations

import json
import os
import re
import time
from typing import Dict, Optional

import numpy as np
import torch

import gguf

from .pack_gptq_into_gguf import K_QUANTS, _packed_bytes, llama_permute

# exact bits per weight of the K-quant block layouts (block bytes * 8 / 256), as tabulated by the reference splitter
EXACT_BITS = {"Q2_K": 2.5625, "Q3_K": 3.4375, "Q4_K": 4.5, "Q5_K": 5.5, "Q6_K": 6.5625}
HF_LAYER_RE = re.compile(r"^model\.layers\..*\.(q_proj|k_proj|v_proj|o_proj|gate_proj|up_proj|down_proj)$")


def _prefix(qname: str) -> str:
    bw = EXACT_BITS[qname]
    return f"{bw if bw != int(bw) else int(bw)}-{qname}"


def emit_database(dir_model_quant: str, config, out_root: str, dtype: str = "float16", hf_layers: bool = True) -> Dict[str, int]:
    """Emit `<out_root>/layers-gguf` and `<out_root>/layers-hf` for every module directory of `dir_model_quant`
    (the save_dir of quant.py).  Calling it once per quantisation level with the same `out_root` accumulates the levels
    side by side, like the reference's loop over models.  hf_layers=False skips the fp16 HF-layout copies (the splitter's
    --hf-layers switch; 16 GB per level for an 8B model).  Returns counts of files written."""
    cfg = config if isinstance(config, dict) else config.to_dict()
    n_layer, n_head = cfg["num_hidden_layers"], cfg["num_attention_heads"]
    n_kv = cfg.get("num_key_value_heads") or n_head
    tmap = gguf.get_tensor_name_map(gguf.MODEL_ARCH.LLAMA, n_layer)
    gdir, hdir = os.path.join(out_root, "layers-gguf"), os.path.join(out_root, "layers-hf")
    os.makedirs(gdir, exist_ok=True)
    os.makedirs(hdir, exist_ok=True)
    tdtype = torch.float16 if dtype == "float16" else torch.float32
    g_manifest = _load_json(os.path.join(gdir, "manifest.json"), {"model_info": {"use_exact_bitwidth": True}, "layers": {}})
    g_db = _load_json(os.path.join(gdir, "gguf_layer_database.json"), {})
    h_manifest = _load_json(os.path.join(hdir, "manifest.json"), {"model_info": {"dtype": dtype, "use_exact_bitwidth": True}, "layers": {}})
    mapping = _load_json(os.path.join(hdir, "layer_mapping.json"), {})
    counts = {"gguf": 0, "hf": 0}
    for module in sorted(os.listdir(dir_model_quant)):
        path = os.path.join(dir_model_quant, module, "data.pth")
        if not os.path.isfile(path):
            continue
        obj = torch.load(path, map_location="cpu", weights_only=True)
        qtype = K_QUANTS[int(obj["q_type"])]
        qname = qtype.name
        hf_name = module + ".weight"
        gname = tmap.get_name(hf_name, try_suffixes=(".weight",))
        if gname is None:
            raise ValueError(f"cannot map module {module!r} to a GGUF tensor name")
        d_row, d_col = obj["qweight"].shape
        packed = _packed_bytes(obj)
        gbytes = packed
        if module.endswith("q_proj"):
            gbytes = llama_permute(packed, n_head, n_head)
        elif module.endswith("k_proj"):
            gbytes = llama_permute(packed, n_head, n_kv)
        raw = np.ascontiguousarray(gbytes.numpy())
        prefix = _prefix(qname)
        # ---- layers-gguf: raw tensor bytes + metadata (gguf_splitter.py:373-404)
        ldir = os.path.join(gdir, gname)
        os.makedirs(ldir, exist_ok=True)
        with open(os.path.join(ldir, f"{prefix}.pth"), "wb") as f:
            f.write(raw.tobytes())
        info = {"name": gname, "type": int(qtype), "quantization": qname, "bitwidth": EXACT_BITS[qname],
                "exact_bitwidth": EXACT_BITS[qname], "shape": [int(d_col), int(d_row)], "n_elements": int(d_row * d_col),
                "n_bytes": int(raw.nbytes), "data_offset_original": None, "data_filename": f"{prefix}.pth",
                "np_dtype": "uint8", "np_shape": [int(raw.shape[0]), int(raw.shape[1])]}
        with open(os.path.join(ldir, f"{prefix}-metadata.json"), "w") as f:
            json.dump({"tensor_info": info}, f, indent=2)
        g_manifest["layers"].setdefault(gname, {"original_name": gname, "dims": info["shape"], "bitwidths": {}})["bitwidths"][
            str(EXACT_BITS[qname])] = {"filename": f"{prefix}.pth", "metadata_filename": f"{prefix}-metadata.json",
                                      "type": int(qtype), "quantization": qname, "bitwidth": EXACT_BITS[qname],
                                      "exact_bitwidth": EXACT_BITS[qname], "size_bytes": int(raw.nbytes), "shape": info["shape"],
                                      "n_elements": info["n_elements"], "data_offset": None}
        g_db[gname] = {"tensor_type": int(qtype), "quantization": qname, "bitwidth": EXACT_BITS[qname],
                       "exact_bitwidth": EXACT_BITS[qname], "shape": info["shape"], "n_elements": info["n_elements"],
                       "n_bytes": int(raw.nbytes), "data_offset": None}
        counts["gguf"] += 1
        # ---- layers-hf: dequantised weight in HF row order (the splitter gets it from transformers' GGUF loader,
        # which undoes the q/k permutation), fp16; only the block projections (gguf_splitter.py:487-490)
        if hf_layers and HF_LAYER_RE.match(module):
            w = torch.from_numpy(gguf.quants.dequantize(np.ascontiguousarray(packed.numpy()), qtype)).to(tdtype)
            mdir = os.path.join(hdir, module)
            os.makedirs(mdir, exist_ok=True)
            torch.save(w, os.path.join(mdir, f"{prefix}.pth"))
            hinfo = {"name": hf_name, "gguf_mapped_name": gname, "bitwidth": EXACT_BITS[qname], "dtype": str(w.dtype),
                     "shape": list(w.shape), "n_elements": w.numel(), "n_bytes": w.numel() * w.element_size(),
                     "data_filename": f"{prefix}.pth", "requires_grad": False}
            with open(os.path.join(mdir, f"{prefix}-metadata.json"), "w") as f:
                json.dump({"tensor_info": hinfo, "gguf_info": g_db[gname]}, f, indent=2)
            h_manifest["layers"][hf_name] = {"original_name": hf_name, "gguf_mapped_name": gname, "layer_directory": module,
                                             "dims": list(w.shape), "bitwidth": EXACT_BITS[qname], "filename": f"{prefix}.pth",
                                             "metadata_filename": f"{prefix}-metadata.json", "dtype": str(w.dtype),
                                             "size_bytes": w.numel() * w.element_size(), "shape": list(w.shape),
                                             "n_elements": w.numel()}
            mapping[hf_name] = gname
            counts["hf"] += 1
    for man in (g_manifest, h_manifest):
        man["model_info"]["split_timestamp"] = time.time()
    _dump_json(os.path.join(gdir, "manifest.json"), g_manifest)
    _dump_json(os.path.join(gdir, "gguf_layer_database.json"), g_db)
    _dump_json(os.path.join(hdir, "manifest.json"), h_manifest)
    _dump_json(os.path.join(hdir, "layer_mapping.json"), mapping)
    return counts


def _load_json(path: str, default):
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return default


def _dump_json(path: str, obj) -> None:
    with open(path, "w") as f:
        json.dump(obj, f, indent=2)


def main(argv: Optional[list] = None):
    import argparse
    ap = argparse.ArgumentParser(description="Emit the EvoPress layer database (layers-gguf / layers-hf) from GPTQ results")
    ap.add_argument("model", type=str, help="HF model directory (for config.json)")
    ap.add_argument("--dir_model_quant", type=str, required=True, action="append",
                    help="save_dir of quant.py; repeat the flag for several quantisation levels")
    ap.add_argument("--output_dir", type=str, required=True)
    ap.add_argument("--dtype", choices=["float16", "float32"], default="float16")
    ap.add_argument("--no_hf_layers", action="store_true", help="only layers-gguf (the splitter without --hf-layers)")
    args = ap.parse_args(argv)
    from transformers import AutoConfig
    cfg = AutoConfig.from_pretrained(args.model)
    total = {"gguf": 0, "hf": 0}
    for d in args.dir_model_quant:
        c = emit_database(d, cfg, args.output_dir, args.dtype, hf_layers=not args.no_hf_layers)
        total = {k: total[k] + c[k] for k in total}
    print(json.dumps(total))


if __name__ == "__main__":
    main()
