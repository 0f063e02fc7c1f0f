"""CPU: pin the oracle (oracle/gq_oracle.c) to vectors produced by the reference itself
(tests/golden/make_golden.py ran quant/gptq/src of the reference on CPU)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

TYPES = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}
KEYS = ["qweight", "d", "sq", "dmin", "zq"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _U(g):
    # stored as U^T contiguous == the reference's column-major U; hand the oracle the strided view
    return g["U_colmajor_T"].T


def _five_raw(o):
    return dict(qweight=o[0], d=o[1].view(np.uint16), sq=o[2], dmin=o[3].view(np.uint16), zq=o[4])


@pytest.mark.parametrize("case", ["b1_a.npz", "b1_b.npz"])
@pytest.mark.parametrize("tname", list(TYPES))
def test_step_matches_reference_bit_exact(golden_dir, case, tname):
    g = _load(golden_dir, case)
    o = orc.gptq_step(g["W"], _U(g), TYPES[tname], block_size=int(g["block_size"]))
    got = _five_raw(o)
    for k in KEYS:
        ref = g[f"{tname}_ieee_{k}"]
        assert got[k].dtype == ref.dtype and got[k].shape == ref.shape, k
        assert np.array_equal(got[k], ref), f"{case} {tname} {k}"
    # w written back by the reference loop == dequantize_linear_weight of its outputs
    assert np.array_equal(o[5], g[f"{tname}_ieee_dequant"])
    # unpatched reference (torch's non-IEEE sqrt): report, require >= 95% identical rows
    bad = np.zeros(g["W"].shape[0], bool)
    for k in KEYS:
        bad |= (got[k] != g[f"{tname}_raw_{k}"]).reshape(bad.size, -1).any(1)
    assert bad.mean() <= 0.05


@pytest.mark.parametrize("case", ["b1_a.npz", "b1_b.npz"])
@pytest.mark.parametrize("tname", list(TYPES))
def test_pack_and_dequant_match_reference(golden_dir, case, tname):
    g = _load(golden_dir, case)
    five = [g[f"{tname}_ieee_{k}"] for k in KEYS]
    five[1] = five[1].view(np.float16)
    five[3] = five[3].view(np.float16)
    packed = orc.pack(TYPES[tname], *five)
    assert np.array_equal(packed, g[f"{tname}_ieee_packed"])
    deq = orc.dequantize(TYPES[tname], *five)
    assert np.array_equal(deq, g[f"{tname}_ieee_dequant"])


@pytest.mark.parametrize("tname", list(TYPES))
def test_packed_bytes_decode_with_gguf_py(golden_dir, tname):
    """independent cross-check: gguf-py's dequantize of the packed bytes == the reference dequant."""
    gguf = pytest.importorskip("gguf")
    g = _load(golden_dir, "b1_a.npz")
    five = [g[f"{tname}_ieee_{k}"] for k in KEYS]
    five[1] = five[1].view(np.float16)
    five[3] = five[3].view(np.float16)
    packed = orc.pack(TYPES[tname], *five)
    deq = gguf.quants.dequantize(packed, gguf.GGMLQuantizationType(TYPES[tname]))
    assert np.array_equal(deq.astype(np.float32), g[f"{tname}_ieee_dequant"])


@pytest.mark.parametrize("tname", list(TYPES))
def test_search_edge_cases(golden_dir, tname):
    g = _load(golden_dir, "search_edge.npz")
    d, sq, dmin, zq = orc.get_scale_and_zero(g["x"], TYPES[tname])
    assert np.array_equal(d.view(np.uint16), g[f"{tname}_ieee_d"])
    assert np.array_equal(dmin.view(np.uint16), g[f"{tname}_ieee_dmin"])
    assert np.array_equal(sq, g[f"{tname}_ieee_sq"])
    assert np.array_equal(zq, g[f"{tname}_ieee_zq"])


@pytest.mark.parametrize("tname", list(TYPES))
def test_rtn_matches_reference(golden_dir, tname):
    g = _load(golden_dir, "rtn.npz")
    got = _five_raw(orc.rtn_quantize(g["W"], TYPES[tname]))
    for k in KEYS:
        assert np.array_equal(got[k], g[f"{tname}_ieee_{k}"]), k


def test_large_validation_was_clean(golden_dir):
    """make_golden.py's 1024x1024 oracle-vs-reference run (not stored) must have had 0 mismatching rows."""
    s = json.load(open(os.path.join(golden_dir, "golden_stats.json")))
    for t, row in s["oracle_vs_reference_1024x1024"].items():
        assert row["ieee"]["mismatching_rows"] == 0, t


def test_prepare_factorisation_properties():
    rng = np.random.default_rng(0)
    n, d_row = 96, 8
    X = rng.standard_normal((400, n)).astype(np.float32)
    H = np.zeros((n, n), np.float32)
    orc.hessian_update(H, X, 0.0, 2.0)
    assert np.allclose(H, 2.0 * X.T.astype(np.float64) @ X.astype(np.float64), rtol=1e-5, atol=1e-4)
    W = rng.standard_normal((d_row, n)).astype(np.float32)
    W[:, 3] = 0.0
    U, Hd, Wm, bad = orc.prepare(H, W, 0.01)
    assert not bad
    assert np.allclose(np.tril(U, -1), 0)
    inv = np.linalg.inv(Hd.astype(np.float64))
    assert np.allclose(U.T.astype(np.float64) @ U.astype(np.float64), inv, rtol=1e-3, atol=1e-6)

    assert np.all(Hd[3, :3] == 0) and np.all(Hd[:3, 3] == 0)


@pytest.mark.parametrize("variant", ["static", "actorder"])
@pytest.mark.parametrize("tname", list(TYPES))
def test_static_groups_and_act_order_match_reference(golden_dir, variant, tname):
    """gptq.py:184-216, 233-238, 273-277: scales searched up front on the original W (static_groups), columns visited
    in descending diag(H) order with per-column group lookup (act_order); Q3_K ignores both (:204-206)."""
    g = _load(golden_dir, "variants_a.npz")
    uses_perm = variant == "actorder" and tname != "Q3_K"
    o = orc.gptq_step(g["W"], g["U_perm"] if uses_perm else g["U_plain"], TYPES[tname], static_groups=True,
                      perm=g["perm"] if uses_perm else None)
    got = _five_raw(o)
    for k in KEYS:
        ref = g[f"{variant}_{tname}_{k}"]
        assert np.array_equal(got[k].view(ref.dtype), ref), f"{variant} {tname} {k}"
    # the weights handed back are the dequantisation of those outputs, in the original column order
    five = [got["qweight"].view(o[0].dtype), o[1], o[2], o[3], o[4]]
    assert np.array_equal(o[5], orc.dequantize(TYPES[tname], *five))


# ------------------------------------------------------------------------------------------------
# non-block RTN on a BF16 weight: the reference searches the scales in bf16 arithmetic (quantizer.py:303-305)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_rtn_bf16_arithmetic_matches_reference_golden(golden_dir, tname):
    """tests/golden/rtn_bf16.npz (make_golden_rtn_bf16.py: the reference's _quant_non_block_module on a bf16 weight with zero,
    constant and negative-constant groups): the oracle's bf16 mode reproduces all five tensors bit for bit; the fp32-arithmetic
    search, which the CUDA path uses by default for bf16 weights, does NOT (documented deviation)."""
    g = np.load(os.path.join(golden_dir, "rtn_bf16.npz"))
    W = (g["W_bf16_bits"].astype(np.uint32) << 16).view(np.float32)
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    out = orc.rtn_quantize(W, qt, bf16=True)
    for k, a in zip(("qweight", "d", "sq", "dmin", "zq"), out):
        a = a.view(np.uint16) if a.dtype == np.float16 else a
        assert np.array_equal(a.view(np.uint8), g[f"{tname}_{k}"].view(np.uint8)), f"{tname}.{k}"
    out32 = orc.rtn_quantize(W, qt, bf16=False)
    same = float((out32[0] == out[0]).mean())
    assert same < 1.0, "the fp32-arithmetic search is expected to differ from the bf16 one on this input"
    assert same > 0.85, same          # measured 0.896 (Q4_K) ... 0.990 (Q3_K)


@pytest.mark.parametrize("tname", ["Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"])
def test_rtn_fp16_arithmetic_matches_reference_golden(golden_dir, tname):
    """The same for an FP16 weight (tests/golden/rtn_f16.npz): same rule set, rounding to fp16."""
    g = np.load(os.path.join(golden_dir, "rtn_f16.npz"))
    W = g["W_f16_bits"].view(np.float16).astype(np.float32)
    qt = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}[tname]
    out = orc.rtn_quantize(W, qt, fp16=True)
    for k, a in zip(("qweight", "d", "sq", "dmin", "zq"), out):
        a = a.view(np.uint16) if a.dtype == np.float16 else a
        assert np.array_equal(a.view(np.uint8), g[f"{tname}_{k}"].view(np.uint8)), f"{tname}.{k}"


@pytest.mark.parametrize("bs", [32, 64, 256])
@pytest.mark.parametrize("tname", list(TYPES))
def test_step_other_block_sizes_match_reference(golden_dir, tname, bs):
    """GPTQ.step with --block_size 32 / 64 / 256 (gptq.py:55, 219-270): the oracle against the reference's own outputs
    (tests/golden/blocksize_a.npz, generated by make_golden_blocksize.py), bit for bit."""
    g = _load(golden_dir, "blocksize_a.npz")
    o = orc.gptq_step(g["W"], _U(g), TYPES[tname], block_size=bs)
    got = _five_raw(o)
    for k in KEYS:
        assert np.array_equal(got[k], g[f"bs{bs}_{tname}_{k}"]), f"bs{bs} {tname} {k}"
    assert np.array_equal(o[5], g[f"bs{bs}_{tname}_dequant"])
    assert np.array_equal(orc.pack(TYPES[tname], *o[:5]), g[f"bs{bs}_{tname}_packed"])
