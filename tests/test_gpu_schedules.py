"""GPU: the two schedules of the exact arithmetic (include/gq.h, gq_mode) are bit-identical.

GQ_MODE_EXACT_LEFT  = one left-looking launch per layer (every 32-row CTA applies all earlier blocks to its own tile);
GQ_MODE_EXACT_RIGHT = per 256-column super-block a panel launch + exact_update_kernel over the whole trailing part
                      (used for row slices of wide projections on several GPUs, where the left-looking kernel leaves
                      most SMs idle);  GQ_MODE_EXACT picks right-looking from 4 super-blocks up (measured faster at
                      every Llama-3-8B shape), left-looking for narrower layers.
Both must reproduce the reference goldens (gptq.py:146-295) and the oracle bit for bit, for all five types, ragged row
counts, the static_groups / act_order variants, and a d_col = 14336 slice like the one a rank of an 8-GPU run gets."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.test_gpu_parity import KEYS, TYPES, assert_five_equal, dev, raw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import ops as _ops
    return _ops


def _modes():
    from gptq_gguf_toolkit_b200._lib import GQ_MODE_EXACT, GQ_MODE_EXACT_LEFT, GQ_MODE_EXACT_RIGHT
    return {"auto": GQ_MODE_EXACT, "left": GQ_MODE_EXACT_LEFT, "right": GQ_MODE_EXACT_RIGHT}


@pytest.mark.parametrize("sched", ["left", "right"])
@pytest.mark.parametrize("case", ["b1_a.npz", "b1_b.npz"])
@pytest.mark.parametrize("tname", list(TYPES))
def test_both_schedules_match_reference_golden(ops, golden_dir, case, tname, sched):
    g = np.load(os.path.join(golden_dir, case))
    out = ops.gptq_quantize(dev(g["W"]), dev(g["U_colmajor_T"].T), TYPES[tname], block_size=int(g["block_size"]),
                            wdeq_dtype=torch.float32, mode=_modes()[sched])
    torch.cuda.synchronize()
    assert_five_equal(out[:5], [g[f"{tname}_ieee_{k}"] for k in KEYS], f"{case}/{tname}/{sched}")
    assert np.array_equal(raw(out[5]), g[f"{tname}_ieee_packed"]), "packed GGUF bytes"
    assert np.array_equal(raw(out[6]), g[f"{tname}_ieee_dequant"]), "dequantised weights"


@pytest.mark.parametrize("tname", ["Q2_K", "Q4_K", "Q6_K"])
@pytest.mark.parametrize("shape", [(100, 1280), (33, 256), (40, 2048)])
def test_right_looking_matches_oracle_ragged(ops, tname, shape):
    d_row, d_col = shape
    rng = np.random.default_rng(d_row * 11 + d_col)
    W = (rng.standard_normal((d_row, d_col)) * 0.05 * np.exp(0.5 * rng.standard_normal((d_row, 1)))).astype(np.float32)
    X = (rng.standard_normal((2 * d_col, d_col)) @ (rng.standard_normal((d_col, d_col)) / np.sqrt(d_col))
         * np.exp(rng.standard_normal(d_col))).astype(np.float32)
    H = np.zeros((d_col, d_col), np.float32)
    orc.hessian_update(H, X, 0.0, 2.0 / 4)
    U, _, _, bad = orc.prepare(H, W, 0.01)
    assert not bad
    ref = orc.gptq_step(W, U, TYPES[tname])
    out = ops.gptq_quantize(dev(W), dev(U), TYPES[tname], wdeq_dtype=torch.float32, mode=_modes()["right"])
    torch.cuda.synchronize()
    assert_five_equal(out[:5], ref[:5], f"{shape}/{tname}")
    assert np.array_equal(raw(out[5]), orc.pack(TYPES[tname], *ref[:5]))
    assert np.array_equal(raw(out[6]), ref[5])


@pytest.mark.parametrize("variant", ["static_groups", "act_order"])
def test_right_looking_variants_match_left(ops, variant):
    rng = np.random.default_rng(3)
    d_row, d_col = 72, 1024
    W = torch.from_numpy((rng.standard_normal((d_row, d_col)) * 0.05).astype(np.float32)).cuda()
    Un = np.triu(rng.standard_normal((d_col, d_col)) * 0.02) + np.eye(d_col)
    U = torch.from_numpy(Un.astype(np.float32)).cuda()
    perm = torch.randperm(d_col, generator=torch.Generator().manual_seed(0)).cuda() if variant == "act_order" else None
    outs = {}
    for sched in ("left", "right"):
        outs[sched] = ops.gptq_quantize(W.clone(), U, 12, wdeq_dtype=torch.float32, mode=_modes()[sched],
                                        static_groups=True, perm=perm)
    torch.cuda.synchronize()
    for a, b in zip(outs["left"][:7], outs["right"][:7]):
        assert torch.equal(a, b), variant


@pytest.mark.parametrize("rows", [512, 96])
def test_down_proj_slice_all_schedules_identical(ops, rows):
    """d_col = 14336 (56 super-blocks), a row slice as one rank of a multi-GPU run sees it: auto (-> right-looking here)
    == left == right on every output, including the propagated errors left in W."""
    d_col = 14336
    g = torch.Generator(device="cuda").manual_seed(rows)
    W = torch.randn(rows, d_col, device="cuda", generator=g) * 0.02
    U = torch.triu(torch.randn(d_col, d_col, device="cuda", generator=g) * (0.3 / d_col ** 0.5))
    U.diagonal().copy_(1.0 + 0.1 * torch.rand(d_col, device="cuda", generator=g))
    got = {}
    for sched, mode in _modes().items():
        Wk = W.clone()
        out = ops.gptq_quantize(Wk, U, 12, wdeq_dtype=torch.bfloat16, mode=mode)
        torch.cuda.synchronize()
        got[sched] = tuple(out[:7]) + (Wk,)
    for sched in ("left", "auto"):
        for a, b in zip(got[sched], got["right"]):
            assert torch.equal(a, b), f"{sched} vs right"
    # and the oracle on the first 32 rows (the CPU restatement needs ~10 s for this width)
    ref = orc.gptq_step(W[:32].cpu().numpy(), U.cpu().numpy(), 12)
    assert_five_equal([t[:32] for t in got["right"][:5]], ref[:5], "down slice vs oracle")
