"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libgq.so via
gptq_gguf_toolkit_b200.ops) and is compared with (a) the committed reference-generated golden vectors and
(b) the CPU oracle on the same seeded inputs.  Integer / byte outputs must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TYPES = {"Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}
KEYS = ["qweight", "d", "sq", "dmin", "zq"]


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gptq_gguf_toolkit_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def raw(t):
    a = t.detach().cpu().numpy()
    return a.view(np.uint16) if a.dtype == np.float16 else a


def assert_five_equal(got5, ref5, what):
    for k, g, r in zip(KEYS, got5, ref5):
        g, r = raw(g), (r.view(np.uint16) if r.dtype == np.float16 else r)
        assert g.shape == r.shape, (what, k, g.shape, r.shape)
        bad = g.view(np.uint8).reshape(g.shape[0], -1) != r.view(np.uint8).reshape(r.shape[0], -1)
        assert not bad.any(), f"{what}: {k} differs in {int(bad.any(1).sum())}/{g.shape[0]} rows"


# ------------------------------------------------------------------------------------------------
# B1: (W, U) -> five tensors, packed bytes, dequantised weights; golden vectors from the reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["b1_a.npz", "b1_b.npz"])
@pytest.mark.parametrize("tname", list(TYPES))
def test_step_matches_reference_golden(ops, golden_dir, case, tname):
    g = np.load(os.path.join(golden_dir, case))
    W = dev(g["W"])
    U = dev(g["U_colmajor_T"].T)          # reference U is column-major; the ABI takes row-major
    out = ops.gptq_quantize(W, U, TYPES[tname], block_size=int(g["block_size"]), wdeq_dtype=torch.float32,
                            search_flags=True)
    torch.cuda.synchronize()
    assert_five_equal(out[:5], [g[f"{tname}_ieee_{k}"] for k in KEYS], f"{case}/{tname}")
    assert np.array_equal(raw(out[5]), g[f"{tname}_ieee_packed"]), "packed GGUF bytes"
    assert np.array_equal(raw(out[6]), g[f"{tname}_ieee_dequant"]), "dequantised weights"
    flags = raw(out[7]).astype(np.uint32)
    assert not (flags[:, 1] & ~flags[:, 0]).any(), "degenerate-search marker must be clear on regular inputs"


@pytest.mark.parametrize("tname", list(TYPES))
@pytest.mark.parametrize("shape", [(100, 512), (64, 1280), (33, 256)])
def test_step_matches_oracle(ops, tname, shape):
    d_row, d_col = shape
    rng = np.random.default_rng(d_row * 7 + d_col)
    W = (rng.standard_normal((d_row, d_col)) * 0.05 * np.exp(0.5 * rng.standard_normal((d_row, 1)))).astype(np.float32)
    X = (rng.standard_normal((2 * d_col, d_col)) @ (rng.standard_normal((d_col, d_col)) / np.sqrt(d_col))
         * np.exp(rng.standard_normal(d_col))).astype(np.float32)
    H = np.zeros((d_col, d_col), np.float32)
    orc.hessian_update(H, X, 0.0, 2.0 / 4)
    U, _, _, bad = orc.prepare(H, W, 0.01)
    assert not bad
    ref = orc.gptq_step(W, U, TYPES[tname])
    out = ops.gptq_quantize(dev(W), dev(U), TYPES[tname], wdeq_dtype=torch.float32)
    torch.cuda.synchronize()
    assert_five_equal(out[:5], ref[:5], f"{shape}/{tname}")
    assert np.array_equal(raw(out[5]), orc.pack(TYPES[tname], *ref[:5]))
    assert np.array_equal(raw(out[6]), ref[5])


def test_step_wdeq_dtypes_and_w_clobbered(ops):
    rng = np.random.default_rng(5)
    W = (rng.standard_normal((64, 512)) * 0.05).astype(np.float32)
    U = np.triu(rng.standard_normal((512, 512)).astype(np.float32) * 0.01) + np.eye(512, dtype=np.float32)
    ref = orc.gptq_step(W, U, 12)
    for dt in (torch.bfloat16, torch.float16):
        out = ops.gptq_quantize(dev(W), dev(U), 12, wdeq_dtype=dt)
        torch.cuda.synchronize()
        want = torch.from_numpy(ref[5]).to(dt)
        assert torch.equal(out[6].cpu(), want)


def test_step_rejects_unsupported(ops):
    from gptq_gguf_toolkit_b200._lib import GQError, GQ_ERR_UNSUPPORTED, GQ_ERR_INVALID
    W = torch.zeros(32, 256, device="cuda")
    U = torch.eye(256, device="cuda")
    with pytest.raises(GQError) as e:
        ops.gptq_quantize(W, U, 12, block_size=96)          # 32 / 64 / 128 / 256 are implemented
    assert e.value.status == GQ_ERR_UNSUPPORTED
    with pytest.raises(GQError) as e:
        ops.gptq_quantize(W, U, 12, block_size=64, mode=1)  # the other block sizes: exact arithmetic only
    assert e.value.status == GQ_ERR_UNSUPPORTED
    with pytest.raises(GQError) as e:
        ops.gptq_quantize(W, U, 99)
    assert e.value.status == GQ_ERR_INVALID
    with pytest.raises(GQError):
        ops.gptq_quantize(torch.zeros(32, 256), torch.eye(256), 12)     # CPU tensors: no fallback


@pytest.mark.parametrize("bs", [32, 64, 256])
@pytest.mark.parametrize("tname", list(TYPES))
def test_step_other_block_sizes_golden(ops, golden_dir, tname, bs):
    """--block_size 32 / 64 / 256 (gptq.py:55, 219-270): csrc/gptq_blocksize.cu against the reference's own outputs
    (tests/golden/blocksize_a.npz, make_golden_blocksize.py): five tensors, GGUF bytes and dequantised weights, bit for bit;
    and against the oracle on a wider, ragged problem."""
    g = np.load(os.path.join(golden_dir, "blocksize_a.npz"))
    W, U = dev(g["W"]), dev(np.ascontiguousarray(g["U_colmajor_T"].T))
    out = ops.gptq_quantize(W, U, TYPES[tname], block_size=bs, wdeq_dtype=torch.float32)
    torch.cuda.synchronize()
    want = [g[f"bs{bs}_{tname}_{k}"] for k in ("qweight", "d", "sq", "dmin", "zq")]
    for got, w, k in zip(out[:5], want, KEYS):
        assert np.array_equal(raw(got), w), f"bs{bs}/{tname}: {k}"
    assert np.array_equal(out[5].cpu().numpy(), g[f"bs{bs}_{tname}_packed"])
    assert np.array_equal(out[6].cpu().numpy(), g[f"bs{bs}_{tname}_dequant"])
    rng = np.random.default_rng(bs)
    Wn = (rng.standard_normal((45, 1280)) * 0.05).astype(np.float32)
    Un = (np.triu(rng.standard_normal((1280, 1280)) * 0.01) + np.eye(1280)).astype(np.float32)
    ref = orc.gptq_step(Wn, Un, TYPES[tname], bs)
    out = ops.gptq_quantize(dev(Wn), dev(Un), TYPES[tname], block_size=bs, wdeq_dtype=torch.bfloat16, static_groups=False)
    torch.cuda.synchronize()
    assert_five_equal(out[:5], ref[:5], f"bs{bs}/{tname} vs oracle")
    assert torch.equal(out[6].cpu(), torch.from_numpy(ref[5]).to(torch.bfloat16))


def test_step_all_zero_weights(ops):
    W = np.zeros((32, 512), np.float32)
    U = np.eye(512, dtype=np.float32)
    for t in TYPES.values():
        ref = orc.gptq_step(W, U, t)
        out = ops.gptq_quantize(dev(W), dev(U), t, wdeq_dtype=torch.float32)
        torch.cuda.synchronize()
        assert_five_equal(out[:5], ref[:5], f"zeros/{t}")


# ------------------------------------------------------------------------------------------------
# GQ_MODE_FAST: rank-k updates between super-blocks on tcgen05 (split-fp16 GEMM, csrc/gemm_f16x3.cu).  Not bit-identical by construction
# (tensor cores do not reproduce a sequentially rounded fp32 chain): statistical check against exact mode.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("group,pair", [("1", "0"), ("2", "0"), ("4", "0"), ("2", "1"), ("4", "1")])
@pytest.mark.parametrize("shape,tname", [((100, 1024), "Q4_K"), ((256, 1536), "Q6_K"), ((1024, 4096), "Q4_K")])
def test_fast_mode_statistical(ops, shape, tname, group, pair, monkeypatch):
    from gptq_gguf_toolkit_b200._lib import GQ_MODE_FAST
    monkeypatch.setenv("GQ_FAST_GROUP", group)      # super-blocks per trailing update (read per layer call)
    monkeypatch.setenv("GQ_GEMM_2CTA", pair)        # one CTA per 128 x 256 tile / CTA pairs on 256 x 256 tiles (cta_group::2)
    d_row, d_col = shape
    torch.manual_seed(d_row + d_col)
    W = (torch.randn(d_row, d_col, device="cuda") * 0.03).contiguous()
    X = (torch.randn(3 * d_col, d_col, device="cuda") @ (torch.randn(d_col, d_col, device="cuda") / d_col ** 0.5)).to(torch.bfloat16)
    H = torch.zeros(d_col, d_col, device="cuda")
    ops.hessian_update(H, X, 0.0, 2.0 / 3)
    U, flag = ops.prepare(H, W.clone(), 0.01)
    exact = ops.gptq_quantize(W.clone(), U, TYPES[tname], wdeq_dtype=torch.float32)
    fast = ops.gptq_quantize(W.clone(), U, TYPES[tname], wdeq_dtype=torch.float32, mode=GQ_MODE_FAST)
    torch.cuda.synchronize()
    same_rows = (exact[0] == fast[0]).all(dim=1).float().mean().item()
    same_codes = (exact[0] == fast[0]).float().mean().item()

    def obj(wq):
        dW = (wq - W).double()
        return float(((dW @ H.double()) * dW).sum())

    o_e, o_f = obj(exact[6]), obj(fast[6])
    print(f"fast (group {group}) vs exact {shape} {tname}: identical rows {same_rows:.3f}, identical codes {same_codes:.5f}, objective ratio {o_f / o_e:.6f}")
    # GPTQ is chaotic per row: one flipped rounding changes the rest of that row, so long rows match less often
    assert same_codes >= 0.9, same_codes
    assert abs(o_f - o_e) <= 2e-3 * o_e, (o_f, o_e)
    # internal consistency of the fast outputs: packed bytes and dequantised weights match its own five tensors
    assert torch.equal(fast[5], ops.pack(TYPES[tname], *fast[:5]))
    assert torch.equal(fast[6], ops.dequantize(TYPES[tname], *fast[:5]))


# ------------------------------------------------------------------------------------------------
# scale search, RTN, pack, dequant
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tname", list(TYPES))
def test_search_edge_cases_golden(ops, golden_dir, tname):
    g = np.load(os.path.join(golden_dir, "search_edge.npz"))
    d, sq, dmin, zq = ops.get_scale_and_zero(dev(g["x"]), TYPES[tname])
    torch.cuda.synchronize()
    assert np.array_equal(raw(d), g[f"{tname}_ieee_d"])
    assert np.array_equal(raw(dmin), g[f"{tname}_ieee_dmin"])
    assert np.array_equal(raw(sq), g[f"{tname}_ieee_sq"])
    assert np.array_equal(raw(zq), g[f"{tname}_ieee_zq"])


def test_search_strided_input_and_nstep(ops):
    rng = np.random.default_rng(1)
    big = rng.standard_normal((37, 1024)).astype(np.float32) * 0.1
    x = dev(big)[:, 256:512]                  # row stride 1024
    for nstep in (0, 5, 20):
        d, sq, dmin, zq = ops.get_scale_and_zero(x, 12, nstep=nstep)
        rd, rsq, rdmin, rzq = orc.get_scale_and_zero(big[:, 256:512], 12, nstep=nstep)
        assert np.array_equal(raw(d), rd.view(np.uint16)) and np.array_equal(raw(sq), rsq)
        assert np.array_equal(raw(dmin), rdmin.view(np.uint16)) and np.array_equal(raw(zq), rzq)


@pytest.mark.parametrize("tname", list(TYPES))
def test_rtn_golden_and_oracle(ops, golden_dir, tname):
    g = np.load(os.path.join(golden_dir, "rtn.npz"))
    out = ops.rtn_quantize(dev(g["W"]), TYPES[tname], wdeq_dtype=torch.float32)
    torch.cuda.synchronize()
    ref5 = [g[f"{tname}_ieee_{k}"] for k in KEYS]
    assert_five_equal(out[:5], ref5, f"rtn/{tname}")
    five = [ref5[0], ref5[1].view(np.float16), ref5[2], ref5[3].view(np.float16), ref5[4]]
    assert np.array_equal(raw(out[5]), orc.pack(TYPES[tname], *five))
    assert np.array_equal(raw(out[6]), orc.dequantize(TYPES[tname], *five))
    # ragged rows, bf16 weights (arithmetic is fp32 on the widened values)
    rng = np.random.default_rng(2)
    Wb = torch.from_numpy((rng.standard_normal((77, 768)) * 0.03).astype(np.float32)).to(torch.bfloat16)
    out = ops.rtn_quantize(Wb.cuda(), TYPES[tname], wdeq_dtype=torch.bfloat16)
    ref = orc.rtn_quantize(Wb.float().numpy(), TYPES[tname])
    torch.cuda.synchronize()
    assert_five_equal(out[:5], ref, f"rtn-bf16/{tname}")
    assert torch.equal(out[6].cpu(), torch.from_numpy(orc.dequantize(TYPES[tname], *ref)).to(torch.bfloat16))


@pytest.mark.parametrize("tname", list(TYPES))
def test_pack_dequant_golden_and_gguf_py(ops, golden_dir, tname):
    g = np.load(os.path.join(golden_dir, "b1_a.npz"))
    five = [dev(g[f"{tname}_ieee_{k}"]) for k in KEYS]
    five[1] = five[1].view(torch.float16)
    five[3] = five[3].view(torch.float16)
    packed = ops.pack(TYPES[tname], *five)
    assert np.array_equal(raw(packed), g[f"{tname}_ieee_packed"])
    deq = ops.dequantize(TYPES[tname], *five)
    assert np.array_equal(raw(deq), g[f"{tname}_ieee_dequant"])
    if TYPES[tname] in (11, 14):      # Q3_K / Q6_K packers take three tensors
        assert np.array_equal(raw(ops.pack(TYPES[tname], five[0], five[1], five[2])), g[f"{tname}_ieee_packed"])
    gguf = pytest.importorskip("gguf")
    back = gguf.quants.dequantize(raw(packed), gguf.GGMLQuantizationType(TYPES[tname]))
    assert np.array_equal(back.astype(np.float32), raw(deq))


# ------------------------------------------------------------------------------------------------
# Hessian and Cholesky chain (floating point: tolerance stated in each test)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(params=[("1", "1"), ("1", "0"), ("0", "0")], ids=["mn_major_2cta", "mn_major_1cta", "transposed_copy"])
def hessian_path(request, monkeypatch):
    """The three kernels of hessian_tc.cu: X fed as MN-major operands straight from the activations to a cta_group::2 kernel (CTA
    pairs on 256 x 256 tiles; the default since round 2), the same operands on one CTA per 128 x 256 tile (GQ_HESSIAN_2CTA=0), and
    round 1's transposed K-major copy (GQ_HESSIAN_MN=0); the library reads the variables on every call."""
    monkeypatch.setenv("GQ_HESSIAN_MN", request.param[0])
    monkeypatch.setenv("GQ_HESSIAN_2CTA", request.param[1])
    return request.param


def test_hessian_tensor_core_path_large(ops, hessian_path):
    """tcgen05 SYRK at a Llama-sized d_col: several tiles per CTA, both TMEM accumulator buffers, k tail padding."""
    torch.manual_seed(3)
    d_col, T = 4096, 4096 + 40
    x = (torch.randn(T, d_col, device="cuda") * torch.linspace(0.1, 3.0, d_col, device="cuda")).to(torch.bfloat16)
    H = torch.randn(d_col, d_col, device="cuda")
    H = (H + H.T).contiguous()
    ref = 0.25 * H.double() + 0.5 * (x.double().T @ x.double())
    ops.hessian_update(H, x, 0.25, 0.5)
    torch.cuda.synchronize()
    # exact bf16 products, fp32 accumulation over 4136 tokens: 5e-5 of the matrix scale
    assert (H.double() - ref).abs().max() <= 5e-5 * ref.abs().max()
    assert torch.equal(H, H.T)


@pytest.mark.parametrize("d_col", [384, 512, 1280])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_hessian_update(ops, dt, d_col, hessian_path):
    torch.manual_seed(0)
    H = torch.zeros(d_col, d_col, device="cuda")
    ref = torch.zeros(d_col, d_col, dtype=torch.float64)
    n = 0
    for b, t in [(1, 200), (3, 77), (2, 128)]:
        x = (torch.randn(b, t, d_col) * 0.5).to(dt)
        beta, alpha = n / (n + b), 2.0 / (n + b)
        ops.hessian_update(H, x.reshape(-1, d_col).cuda().contiguous(), beta, alpha)
        xd = x.reshape(-1, d_col).double()
        ref = beta * ref + alpha * (xd.T @ xd)
        n += b
    torch.cuda.synchronize()
    Hc = H.cpu().double()
    # fp32 accumulation of exact products: relative error to the fp64 result <= 1e-5 of the matrix scale
    assert (Hc - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert torch.equal(H, H.T), "H must be exactly symmetric"


def _spd_problem(n, d_row, seed):
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((3 * n, n)) @ (rng.standard_normal((n, n)) / np.sqrt(n))).astype(np.float32)
    H = (2.0 / 3 * (X.T.astype(np.float64) @ X.astype(np.float64)) / n).astype(np.float32)
    H = ((H + H.T) / 2).astype(np.float32)
    W = (rng.standard_normal((d_row, n)) * 0.05).astype(np.float32)
    return H, W


_DIAG = [("4", "diag_v4"), ("3", "diag_v3"), ("1", "diag_v2"), ("0", "diag_v1")]


@pytest.fixture(params=[v for v, _ in _DIAG], ids=[n for _, n in _DIAG])
def diag_variant(request, monkeypatch):
    """All diagonal-block kernels of gq_prepare (csrc/linalg.cu: the two-level chol_diag_v4_kernel, the default since round 2,
    the register-resident chol_diag_v3_kernel, chol_diag_v2_kernel and the original chol_diag_kernel) must meet the same B2 bounds; the library reads
    GQ_DIAG_V2 on every call."""
    monkeypatch.setenv("GQ_DIAG_V2", request.param)
    return request.param


@pytest.mark.parametrize("n", [128, 384, 896, 2048])
def test_prepare_matches_double_precision(ops, n, diag_variant):
    H, W = _spd_problem(n, 16, n)
    W[:, 7] = 0.0                              # an all-zero weight column (gptq.py:308-313)
    Hd, Wd = dev(H), dev(W)
    U, flag = ops.prepare(Hd, Wd, 0.01)
    torch.cuda.synchronize()
    Uo, Ho, _, bad = orc.prepare(H, W, 0.01)
    assert not bad and int(flag.item()) == 0
    Ug = U.cpu().numpy()
    assert np.all(np.tril(Ug, -1) == 0), "U must be exactly upper triangular"
    # H was masked and damped in place like the reference does
    assert np.allclose(Hd.cpu().numpy(), Ho, rtol=1e-6, atol=1e-7)
    # fp32 factorisation vs fp64: 2e-4 relative to the largest entry (cond(H) ~ 1e3..1e4 here)
    assert np.abs(Ug - Uo).max() <= 2e-4 * np.abs(Uo).max()
    inv = np.linalg.inv(Ho.astype(np.float64))
    got = Ug.T.astype(np.float64) @ Ug.astype(np.float64)
    assert np.abs(got - inv).max() <= 5e-4 * np.abs(inv).max()


def test_prepare_not_positive_definite_falls_back_to_identity(ops, diag_variant):
    n = 256
    H = -np.eye(n, dtype=np.float32)
    W = np.ones((8, n), np.float32)
    U, flag = ops.prepare(dev(H), dev(W), 0.01)
    torch.cuda.synchronize()
    assert int(flag.item()) == 1
    assert torch.equal(U.cpu(), torch.eye(n))


def test_pre_step_dead_channels(ops):
    H, W = _spd_problem(256, 8, 3)
    H[5, :] = 0
    H[:, 5] = 0
    Hd, Wd = dev(H), dev(W)
    ops.pre_step(Hd, Wd)
    torch.cuda.synchronize()
    assert Hd[5, 5].item() == 1.0 and (Wd[:, 5] == 0).all()
    assert torch.equal(Wd[:, :5].cpu(), torch.from_numpy(W[:, :5]))


# ------------------------------------------------------------------------------------------------
# B2: whole handle (GPU Hessian + GPU Cholesky + exact loop) vs the oracle fed the oracle's own U.
# The reference does not reproduce itself at this boundary (LAPACK bits differ between thread counts), so
# the check is statistical: most rows identical, layer objective tr(dW H dW^T) within 1% of the oracle's.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q2_K"])
def test_handle_end_to_end_statistical(ops, tname):
    from gptq_gguf_toolkit_b200.gptq import GPTQ
    torch.manual_seed(1)
    d_row, d_col = 128, 512
    layer = torch.nn.Linear(d_col, d_row, bias=False).cuda()
    layer.weight.data.mul_(0.5)
    mix = torch.randn(d_col, d_col) / d_col ** 0.5
    xs = [((torch.randn(1, 96, d_col) @ mix) * torch.exp(0.5 * torch.randn(d_col))) for _ in range(6)]
    h = GPTQ(layer, rel_damp=0.01, block_size=128)
    Hn = np.zeros((d_col, d_col), np.float32)
    n = 0
    for x in xs:
        h.update(x.cuda())
        orc.hessian_update(Hn, x.reshape(-1, d_col).numpy(), n / (n + 1), 2.0 / (n + 1))
        n += 1
    W0 = layer.weight.data.float().cpu().numpy()
    five = h.quantize(TYPES[tname])
    torch.cuda.synchronize()
    Uo, Hdamped, _, bad = orc.prepare(Hn, W0, 0.01)
    ref = orc.gptq_step(W0, Uo, TYPES[tname])
    same_rows = np.mean([(raw(five[0])[r] == ref[0][r]).all() for r in range(d_row)])
    assert same_rows >= 0.5, f"only {same_rows:.2%} rows identical to the oracle"

    def objective(wq):
        dW = (wq - W0).astype(np.float64)
        return float(np.einsum("ij,jk,ik->", dW, Hdamped.astype(np.float64), dW))

    o_gpu, o_ref = objective(h.wdeq.float().cpu().numpy()), objective(ref[5])
    assert abs(o_gpu - o_ref) <= 0.01 * o_ref, (o_gpu, o_ref)
    assert h.packed.shape == (d_row, d_col // 256 * orc.fmt(TYPES[tname])["type_size"])
    assert not h.non_invertible()


# ------------------------------------------------------------------------------------------------
# Full-size (Llama-3-8B q_proj shape) size-independent properties
# ------------------------------------------------------------------------------------------------
def test_full_size_properties(ops):
    torch.manual_seed(0)
    d_row = d_col = 4096
    W = (torch.randn(d_row, d_col, device="cuda") * 0.02).to(torch.bfloat16).float()
    X = torch.randn(8192, d_col, device="cuda").to(torch.bfloat16)
    H = torch.zeros(d_col, d_col, device="cuda")
    ops.hessian_update(H, X, 0.0, 2.0 / 4)
    Wc = W.clone()
    U, flag = ops.prepare(H, Wc, 0.01)
    out = ops.gptq_quantize(Wc, U, 12, wdeq_dtype=torch.float32)
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    qweight, d, sq, dmin, zq, packed, wdeq, _ = out
    assert qweight.dtype == torch.uint8 and int(qweight.max()) <= 15
    assert int(sq.max()) <= 63 and int(zq.max()) <= 63
    # packed bytes == standalone pack of the five tensors; gguf-py decodes them to exactly wdeq
    assert torch.equal(packed, ops.pack(12, qweight, d, sq, dmin, zq))
    assert torch.equal(wdeq, ops.dequantize(12, qweight, d, sq, dmin, zq))
    gguf = pytest.importorskip("gguf")
    rows = slice(0, 64)
    back = gguf.quants.dequantize(packed[rows].cpu().numpy(), gguf.GGMLQuantizationType.Q4_K)
    assert np.array_equal(back.astype(np.float32), wdeq[rows].cpu().numpy())
    # GPTQ must beat RTN on the layer objective tr(dW H dW^T)
    rtn = ops.rtn_quantize(W, 12, wdeq_dtype=torch.float32)[6]

    def obj(wq):
        dW = wq - W
        return float(((dW @ H) * dW).sum())

    assert obj(wdeq) < obj(rtn)
    # a slice of rows checked bit-exactly against the oracle with the GPU's own U
    ref = orc.gptq_step(W[:64].cpu().numpy(), U.cpu().numpy(), 12)
    assert np.array_equal(raw(qweight[:64]), ref[0]) and np.array_equal(raw(d[:64]), ref[1].view(np.uint16))
