"""CPU tests (pytest -m "not gpu"): the C ABI loads and exports every symbol include/gq.h declares, the host
mirror fails loudly without CUDA, and the host-side logic (driver, sharing, stacking, row sharding, CLI
config) behaves -- with the oracle standing in for the kernels (tests/_oracle_backend.py, tests only)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}


# ------------------------------------------------------------------------------------------------
# C ABI
# ------------------------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol():
    from gptq_gguf_toolkit_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gq.h")).read()
    declared = set(re.findall(r"GQ_API\s+[\w\s\*]+?\b(gq_\w+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gq.h but not exported by libgq.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == declared


def test_abi_no_compute_without_gpu_and_error_text():
    from gptq_gguf_toolkit_b200 import _lib
    lib = _lib.load()
    assert lib.gq_abi_version() == 1
    for name, t in T.items():
        f = _lib.format_info(t)
        assert f == orc.fmt(t), name
    with pytest.raises(_lib.GQError) as e:
        _lib.format_info(9)
    assert e.value.status == _lib.GQ_ERR_INVALID and b"q_type" in lib.gq_last_error()
    assert lib.gq_pre_step(None, None, 4, 256, None) == _lib.GQ_ERR_INVALID
    assert lib.gq_prepare_workspace_bytes(256) >= 2 * 256 * 256 * 4
    assert lib.gq_launch_count() == 0


def test_product_path_refuses_cpu_tensors():
    """There is no CPU fallback: the mirror must raise, not compute, when handed CPU tensors."""
    from gptq_gguf_toolkit_b200 import ops, packing_utils, quant_utils
    from gptq_gguf_toolkit_b200._lib import GQError
    with pytest.raises(GQError):
        ops.rtn_quantize(torch.zeros(32, 256), 12)
    with pytest.raises(GQError):
        ops.gptq_quantize(torch.zeros(32, 256), torch.eye(256), 12)
    with pytest.raises(GQError):
        quant_utils.Quantizer().get_scale_and_zero(torch.zeros(4, 256), quant_utils.GGMLQuantizationType.Q4_K)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            packing_utils.pack_Q4K(torch.zeros(4, 256, dtype=torch.uint8), torch.zeros(4, 1), torch.zeros(4, 8, dtype=torch.uint8),
                                   torch.zeros(4, 1), torch.zeros(4, 8, dtype=torch.uint8))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gptq_gguf_toolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("# oracle", ""), f"{f} mentions the oracle"


def test_format_registry_matches_reference_table():
    from gptq_gguf_toolkit_b200.quant_utils import GGML_QUANT_SIZES, GGMLQuantizationType, GGUF_TYPE_SIZE
    for qt in GGMLQuantizationType:
        bits, clamp, smq, gs, sgs, sdt, qdt = GGML_QUANT_SIZES[qt]
        f = orc.fmt(int(qt))
        assert (bits, clamp[0], clamp[1], smq, gs, sgs) == (f["bits"], f["qmin"], f["qmax"], f["scale_maxq"], f["group_size"], 256)
        assert GGUF_TYPE_SIZE[qt] == f["type_size"]
        assert (sdt == torch.uint8) == bool(f["asym"])


# ------------------------------------------------------------------------------------------------
# CLI surface
# ------------------------------------------------------------------------------------------------
def test_quant_config_builder(tmp_path):
    from gptq_gguf_toolkit_b200.quant import build_quant_config, parse_args
    from gptq_gguf_toolkit_b200.quant_utils import GGMLQuantizationType as Q
    cfg = build_quant_config("Q4_K", None)
    assert cfg["q_proj"] == Q.Q4_K and cfg["lm_head"] == Q.Q4_K and len(cfg) == 9
    p = tmp_path / "config.json"
    p.write_text(json.dumps({"q_proj": "Q2_K", "down_proj": "Q6_K"}))
    cfg = build_quant_config("Q4_K", str(p))                 # JSON replaces the uniform config wholesale (quant.py:203-217)
    assert cfg == {"q_proj": Q.Q2_K, "down_proj": Q.Q6_K}
    with pytest.raises(ValueError):
        build_quant_config("Q8_0", None)
    with pytest.raises(ValueError):
        build_quant_config("Q4_K", str(tmp_path / "missing.json"))
    a = parse_args(["--model_name_or_path", "m", "--quantizable_modules", "x", "--pre_block_modules", "model.embed_tokens",
                    "--block_modules", "model.layers", "--calibration_data", "c", "--save_dir", "s"])
    assert (a.rel_damp, a.block_size, a.rmin, a.rdelta, a.nstep, a.default_bit_width) == (1e-2, 128, -1.0, 0.1, 20, "Q4_K")


def test_select_layers_regex():
    from transformers import LlamaConfig, LlamaForCausalLM
    from gptq_gguf_toolkit_b200.model_utils import LINEAR_LAYERS, select_layers
    m = LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                     num_attention_heads=2, num_key_value_heads=1))
    got = select_layers(m, "model.layers.1.", r".*layers.*((q|k|v|o|gate|up|down)_proj)$", LINEAR_LAYERS)
    assert [n.split(".")[-1] for n in got] == ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
    assert all(n.startswith("model.layers.1.") for n in got)


# ------------------------------------------------------------------------------------------------
# driver logic with the oracle standing in for the kernels
# ------------------------------------------------------------------------------------------------
def _tiny_model(seed=0):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    cfg = LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).float().eval()


def _run_driver(monkeypatch, tmp_path, tag, quant_config=None, quantizer_kwargs_extra=None, **kw):
    from tests import _oracle_backend as ob
    ob.install(monkeypatch)
    from gptq_gguf_toolkit_b200.quant import build_quant_config
    from gptq_gguf_toolkit_b200.quantizer import Quantizer
    model = _tiny_model()
    g = torch.Generator().manual_seed(1)
    loader = [([], {"input_ids": torch.randint(0, 512, (1, 64), generator=g)}) for _ in range(4)]
    save_dir = str(tmp_path / tag)
    q = Quantizer(model, data_loader=loader, quantizable_modules=r".*layers.*((q|k|v|o|gate|up|down)_proj)$",
                  quantizer_kwargs={**dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                           static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                                    **(quantizer_kwargs_extra or {})},
                  pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                  quant_non_block_modules=True, device="cpu", save_dir=save_dir, keep_results=True, **kw)
    q.quantize(quant_config or build_quant_config("Q4_K", None))
    return model, q, save_dir


def test_driver_writes_reference_schema_and_rtn_is_bit_exact(monkeypatch, tmp_path):
    pristine = _tiny_model()
    model, q, save_dir = _run_driver(monkeypatch, tmp_path, "a", calibration_batch_size=2)
    names = sorted(os.listdir(save_dir))
    assert len(names) == 2 * 7 + 2
    for n in names:
        d = torch.load(os.path.join(save_dir, n, "data.pth"))
        assert set(d) >= {"q_type", "qweight", "super_group_scale", "super_group_zero", "group_scale_quant", "group_zero_quant"}
        assert d["q_type"] == 12 and d["qweight"].dtype == torch.uint8 and d["super_group_scale"].dtype == torch.float16
        w = model.get_submodule(n).weight.data
        assert d["qweight"].shape == w.shape
        deq = orc.dequantize(12, d["qweight"].numpy(), d["super_group_scale"].numpy(), d["group_scale_quant"].numpy(),
                             d["super_group_zero"].numpy(), d["group_zero_quant"].numpy())
        assert np.array_equal(deq, w.numpy()), f"{n}: layer weight must be the dequantised result (quantizer.py:257-264)"
        assert np.array_equal(d["packed"].numpy(), orc.pack(12, d["qweight"].numpy(), d["super_group_scale"].numpy(),
                                                            d["group_scale_quant"].numpy(), d["super_group_zero"].numpy(),
                                                            d["group_zero_quant"].numpy()))
    for n in ("model.embed_tokens", "lm_head"):        # RTN on the pristine fp32 weights: bit-exact
        d = torch.load(os.path.join(save_dir, n, "data.pth"))
        ref = orc.rtn_quantize(pristine.get_submodule(n).weight.data.numpy(), 12)
        assert np.array_equal(d["qweight"].numpy(), ref[0])


def test_sharing_stacking_and_batching_do_not_change_results(monkeypatch, tmp_path):
    _, qa, _ = _run_driver(monkeypatch, tmp_path, "a", calibration_batch_size=1, share_hessians=False)
    _, qb, _ = _run_driver(monkeypatch, tmp_path, "b", calibration_batch_size=1, share_hessians=True)
    _, qc, _ = _run_driver(monkeypatch, tmp_path, "c", calibration_batch_size=4, share_hessians=True)
    for n in qa.results:
        for k in ("qweight", "super_group_scale", "group_scale_quant", "super_group_zero", "group_zero_quant", "packed"):
            assert torch.equal(qa.results[n][k], qb.results[n][k]), f"sharing changed {n}.{k}"
    # batching only changes the fp summation order of H (statistical boundary): most rows must still agree
    same = np.mean([float(torch.equal(qb.results[n]["qweight"][r], qc.results[n]["qweight"][r]))
                    for n in qb.results for r in range(0, qb.results[n]["qweight"].shape[0], 7)])
    assert same > 0.6


def test_mixed_bit_width_configuration(monkeypatch, tmp_path):
    from gptq_gguf_toolkit_b200.quant_utils import GGMLQuantizationType as Q
    cfg = {"q_proj": Q.Q2_K, "k_proj": Q.Q4_K, "v_proj": Q.Q4_K, "o_proj": Q.Q6_K, "gate_proj": Q.Q3_K, "up_proj": Q.Q5_K,
           "down_proj": Q.Q4_K}
    _, q, save_dir = _run_driver(monkeypatch, tmp_path, "m", quant_config=cfg, calibration_batch_size=2)
    want = {"q_proj": 10, "k_proj": 12, "v_proj": 12, "o_proj": 14, "gate_proj": 11, "up_proj": 13, "down_proj": 12,
            "embed_tokens": 14, "lm_head": 14}     # missing non-block keys default to Q6_K (quantizer.py:106-108)
    for n, d in q.results.items():
        assert d["q_type"] == want[n.split(".")[-1]], n
    assert q.results["model.layers.0.mlp.gate_proj"]["qweight"].dtype == torch.int8


def test_gptq_handle_protocol(monkeypatch):
    from tests import _oracle_backend as ob
    ob.install(monkeypatch)
    from gptq_gguf_toolkit_b200.gptq import GPTQ, HessianAccumulator
    torch.manual_seed(0)
    layer = torch.nn.Linear(256, 16, bias=False)
    h = GPTQ(layer, rel_damp=0.01, block_size=128)
    for _ in range(3):
        h.update(torch.randn(2, 20, 256))
    assert h.num_samples == 6 and h.H.shape == (256, 256)
    out = h.quantize(12)
    assert len(out) == 5 and out[0].shape == (16, 256) and h.packed.shape == (16, 144) and h.wdeq.dtype == layer.weight.dtype
    h.reset()
    assert h.H is None and h.num_samples == 0
    with pytest.raises(AssertionError):
        GPTQ(layer, act_order=True, static_groups=False)        # gptq.py:45-46
    acc = HessianAccumulator(256)
    a, b = GPTQ(layer, block_size=128, hessian=acc), GPTQ(torch.nn.Linear(256, 8, bias=False), block_size=128, hessian=acc)
    a.update(torch.randn(1, 40, 256))
    assert b.H is a.H and acc.users == 2
    a.quantize(12); a.reset()
    assert acc.H is not None            # still owned by b
    b.quantize(12); b.reset()
    assert acc.H is None


def test_act_order_and_static_groups_host_logic(monkeypatch, tmp_path):
    """GPTQ handle + driver with act_order / static_groups (gptq.py:184-216): permutation from diag(H), factor of the
    permuted Hessian, Q3_K exemption (:204-206); handle and driver must agree with a direct oracle call."""
    from tests import _oracle_backend as ob
    ob.install(monkeypatch)
    from gptq_gguf_toolkit_b200.gptq import GPTQ
    torch.manual_seed(0)
    layer = torch.nn.Linear(256, 16, bias=False)
    W0 = layer.weight.data.clone()
    xs = [torch.randn(2, 40, 256) * torch.linspace(0.2, 3.0, 256) for _ in range(3)]
    for qt in (12, 11):
        layer.weight.data = W0.clone()
        h = GPTQ(layer, rel_damp=0.01, block_size=128, act_order=True, static_groups=True)
        for x in xs:
            h.update(x)
        H = h.H.clone()
        five = h.quantize(qt)
        perm = None if qt == 11 else torch.argsort(torch.diag(H), descending=True)
        Hp = H if perm is None else H[perm][:, perm]
        Wp = W0 if perm is None else W0[:, perm]
        U = orc.prepare(Hp.numpy().copy(), Wp.numpy().copy(), 0.01)[0]
        ref = orc.gptq_step(W0.numpy(), U, qt, static_groups=True, perm=None if perm is None else perm.numpy())
        for got, want in zip(five, ref[:5]):
            assert np.array_equal(got.numpy().view(np.uint8), np.ascontiguousarray(want).view(np.uint8))
        if qt == 12:
            assert not torch.equal(perm, torch.arange(256))      # the test really permutes
    # driver: act_order changes the result, keeps the schema, and the written weights are the dequantised codes
    kw = dict(quantizer_kwargs_extra=dict(act_order=True, static_groups=True))
    model_a, qa, _ = _run_driver(monkeypatch, tmp_path, "ao", **kw)
    model_b, qb, _ = _run_driver(monkeypatch, tmp_path, "plain")
    name = "model.layers.0.mlp.down_proj"
    assert not torch.equal(qa.results[name]["qweight"], qb.results[name]["qweight"])
    d = qa.results[name]
    deq = orc.dequantize(12, d["qweight"].numpy(), d["super_group_scale"].numpy(), d["group_scale_quant"].numpy(),
                         d["super_group_zero"].numpy(), d["group_zero_quant"].numpy())
    assert np.array_equal(deq, model_a.get_submodule(name).weight.data.numpy())
    assert np.array_equal(d["packed"].numpy(), orc.pack(12, d["qweight"].numpy(), d["super_group_scale"].numpy(),
                                                        d["group_scale_quant"].numpy(), d["super_group_zero"].numpy(),
                                                        d["group_zero_quant"].numpy()))


# ------------------------------------------------------------------------------------------------
# world_size = 2 over gloo: row sharding + all-gather and the Hessian all-reduce
# ------------------------------------------------------------------------------------------------
_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import pytest
from tests import _oracle_backend as ob
from gptq_gguf_toolkit_b200 import gptq as G, quantizer as Q
G.ops = ob; Q.ops = ob
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(0)
W = torch.from_numpy((rng.standard_normal((72, 512)) * 0.05).astype(np.float32))
U = torch.from_numpy((np.triu(rng.standard_normal((512, 512)) * 0.01) + np.eye(512)).astype(np.float32))
qz = Q.Quantizer(None, [], "", dict(block_size=128), [], [], "", None)
full = qz._sharded_gptq(W.clone(), U, 12, torch.float32, rank, world)
ref = ob.gptq_quantize(W.clone(), U, 12, wdeq_dtype=torch.float32)[:7]
ok = all(torch.equal(a, b) for a, b in zip(full, ref))
acc = G.HessianAccumulator(256)
acc.update(torch.full((1, 4, 256), float(rank + 1)))
acc.all_reduce()
want = 2.0 * 4 * (1.0 + 4.0) / 2
ok = ok and bool(torch.allclose(acc.H, torch.full((256, 256), want)))
dist.barrier()
if rank == 0:
    print(json.dumps({{"ok": bool(ok)}}))
dist.destroy_process_group()
"""


def test_world_size_2_gloo_row_sharding_and_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"] is True


def test_conv_layer_handle(monkeypatch):
    """_ConvNd layers (model_utils.py:12, gptq.py:97-107, 139-140 of the reference): the Hessian is accumulated over unfolded
    patches, the weight is flattened to (out_channels, in_channels x kernel), the batch counts images."""
    from tests import _oracle_backend as ob
    ob.install(monkeypatch)
    from gptq_gguf_toolkit_b200.gptq import GPTQ
    from gptq_gguf_toolkit_b200.model_utils import LINEAR_LAYERS, get_number_of_rows_and_cols
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(64, 24, kernel_size=2, stride=1, padding=1, bias=False)
    assert isinstance(conv, LINEAR_LAYERS) and get_number_of_rows_and_cols(conv) == (24, 256)
    h = GPTQ(conv, rel_damp=0.01, block_size=128)
    assert (h.d_row, h.d_col) == (24, 256)
    xs = [torch.randn(3, 64, 6, 5) for _ in range(2)]
    H_ref = torch.zeros(256, 256, dtype=torch.float64)
    n = 0
    for x in xs:
        h.update(x)
        p = torch.nn.functional.unfold(x.double(), 2, padding=1).transpose(1, 2).flatten(0, 1)
        H_ref = H_ref * (n / (n + 3)) + (2.0 / (n + 3)) * (p.T @ p)       # gptq.py:108-112 with batch_size = images
        n += 3
    assert h.num_samples == 6
    assert torch.allclose(h.H.double(), H_ref, rtol=1e-5, atol=1e-5)
    out = h.quantize(12)
    assert out[0].shape == (24, 256) and h.wdeq.shape == (24, 256)
    ref = orc.gptq_step(conv.weight.data.float().flatten(1).numpy(), h.hessian.U.numpy(), 12)
    assert np.array_equal(out[0].numpy(), ref[0])
