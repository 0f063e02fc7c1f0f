#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference (IST-DASLab/gptq-gguf-toolkit @7a38bc5) has no tests or fixtures of its own, so
these vectors -- outputs of the reference's own code on seeded inputs -- are what pins the oracle
(oracle/gq_oracle.c) and, through it, the CUDA path.

Two variants are stored for every case:
  *_ieee : reference with torch.sqrt replaced by a correctly rounded sqrt (t.double().sqrt().float())
           -- torch's CPU sqrt is ~0.7% of the time 1 ulp off IEEE, CUDA's sqrtf is exact.
  *_raw  : reference completely unpatched.
The oracle must equal *_ieee bit for bit; tests report the row match rate against *_raw.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/quant/gptq"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from src.gptq import GPTQ  # noqa: E402
from src import quant_utils as qu  # noqa: E402
from src import packing_utils as pu  # noqa: E402
from src.quantizer import Quantizer as RefQuantizer  # noqa: E402

T = qu.GGMLQuantizationType
TYPES = [T.Q2_K, T.Q3_K, T.Q4_K, T.Q5_K, T.Q6_K]
_orig_sqrt = torch.sqrt


def set_sqrt(ieee: bool):
    torch.sqrt = (lambda t: t.double().sqrt().float()) if ieee else _orig_sqrt


def correlated_x(gen, n_seq, seq_len, d_col):
    """ill-conditioned activations: randn @ mix * colscale."""
    mix = torch.randn(d_col, d_col, generator=gen) / d_col ** 0.5
    colscale = torch.exp(torch.randn(d_col, generator=gen))
    return [(torch.randn(1, seq_len, d_col, generator=gen) @ mix + 0.1 * torch.randn(1, seq_len, d_col, generator=gen))
            * colscale for _ in range(n_seq)]


def run_ref_gptq(W, xs, q_type, block_size=128):
    layer = torch.nn.Linear(W.shape[1], W.shape[0], bias=False)
    layer.weight.data = W.clone()
    h = GPTQ(layer, rel_damp=0.01, block_size=block_size)
    for x in xs:
        h.update(x)
    keep = {}
    orig = h._prepare
    h._prepare = lambda: keep.setdefault("U", orig())
    out = h.quantize(q_type)
    return [t.clone() for t in out], keep["U"], h.H.clone()


def pack_ref(q_type, five):
    qweight, d, sq, dmin, zq = [t.clone() for t in five]
    if q_type == T.Q2_K:
        return pu.pack_Q2K(qweight, d, sq, dmin, zq)
    if q_type == T.Q3_K:
        return pu.pack_Q3K(qweight, d, sq)
    if q_type == T.Q4_K:
        return pu.pack_Q4K(qweight, d, sq, dmin, zq)
    if q_type == T.Q5_K:
        return pu.pack_Q5K(qweight, d, sq, dmin, zq)
    return pu.pack_Q6K(qweight, d, sq)


def np5(five):
    qweight, d, sq, dmin, zq = five
    return dict(qweight=qweight.numpy(), d=d.numpy().view(np.uint16), sq=sq.numpy(),
                dmin=dmin.numpy().view(np.uint16), zq=zq.numpy())


def b1_case(name, d_row, d_col, seed, n_seq=6, seq_len=96, block_size=128, wscale=0.05, dead_col=None):
    gen = torch.Generator().manual_seed(seed)
    W = torch.randn(d_row, d_col, generator=gen) * wscale
    W = W * torch.exp(0.5 * torch.randn(d_row, 1, generator=gen))
    if dead_col is not None:
        W[:, dead_col] = 0.0
    xs = correlated_x(gen, n_seq, seq_len, d_col)
    out = {"W": W.numpy(), "block_size": np.int32(block_size)}
    U0 = None
    for q_type in TYPES:
        for variant in ("ieee", "raw"):
            set_sqrt(variant == "ieee")
            five, U, H = run_ref_gptq(W, xs, q_type, block_size)
            set_sqrt(False)
            assert U.stride() == (1, d_col), U.stride()   # column-major, as SURVEY 8a notes
            if U0 is None:
                U0 = U.clone()
                out["U_colmajor_T"] = U.t().contiguous().numpy()   # row i of this array = column i of U
                out["H_damped"] = H.numpy()
            assert torch.equal(U, U0)
            # the reference returns (qweight, d, sq, dmin, zq)
            five = [five[0], five[1], five[2], five[3], five[4]]
            for k, v in np5(five).items():
                out[f"{q_type.name}_{variant}_{k}"] = v
            if variant == "ieee":
                out[f"{q_type.name}_ieee_packed"] = np.asarray(pack_ref(q_type, five))
                deq = qu.dequantize_linear_weight(q_type, five[0], five[1], five[2], five[3], five[4])
                out[f"{q_type.name}_ieee_dequant"] = deq.numpy()
    np.savez_compressed(os.path.join(HERE, f"b1_{name}.npz"), **out)
    return out


def variants_case(name, d_row, d_col, seed, n_seq=6, seq_len=96):
    """static_groups and act_order (+ static_groups) variants of GPTQ.step (gptq.py:184-216, 233-238, 273-277).
    Stored per (variant, type): the permutation the reference derived from diag(H) (act_order), the U it factored
    from the permuted Hessian, and its five outputs (IEEE-sqrt reference only)."""
    gen = torch.Generator().manual_seed(seed)
    W = torch.randn(d_row, d_col, generator=gen) * 0.05
    W = W * torch.exp(0.5 * torch.randn(d_row, 1, generator=gen))
    xs = correlated_x(gen, n_seq, seq_len, d_col)
    out = {"W": W.numpy()}
    for vname, kw in (("static", dict(static_groups=True)), ("actorder", dict(static_groups=True, act_order=True))):
        for q_type in TYPES:
            set_sqrt(True)
            layer = torch.nn.Linear(d_col, d_row, bias=False)
            layer.weight.data = W.clone()
            h = GPTQ(layer, rel_damp=0.01, block_size=128, **kw)
            for x in xs:
                h.update(x)
            keep = {}
            orig = h._prepare

            def prep():
                # called after the act_order permutation of W / H (gptq.py:209-216): record what the loop will see
                keep["H_diag_order"] = torch.argsort(torch.diag(h.H), descending=True)
                return keep.setdefault("U", orig())
            h._prepare = prep
            H0 = None

            def pre_step_hook(orig_pre=h.quantization_pre_step):
                orig_pre()
                keep["H0"] = h.H.clone()
            h.quantization_pre_step = pre_step_hook
            five = [t.clone() for t in h.quantize(q_type)]
            set_sqrt(False)
            U = keep["U"]
            uses_perm = vname == "actorder" and q_type != T.Q3_K            # Q3_K forces both options off (gptq.py:204-206)
            # U and perm are the same for every type of a kind: stored once ("U_plain", "U_perm" + "perm")
            if uses_perm:
                perm = torch.argsort(torch.diag(keep["H0"]), descending=True).numpy().astype(np.int32)
                assert np.array_equal(out.setdefault("perm", perm), perm)
                assert np.array_equal(out.setdefault("U_perm", U.contiguous().numpy()), U.contiguous().numpy())
            else:
                assert np.array_equal(out.setdefault("U_plain", U.contiguous().numpy()), U.contiguous().numpy())
            for k, v in np5(five).items():
                out[f"{vname}_{q_type.name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, f"variants_{name}.npz"), **out)
    return out


def search_case():
    """Edge cases for get_scale_and_zero (quant_utils.py:90-145)."""
    gen = torch.Generator().manual_seed(7)
    rows = []
    rows.append(torch.randn(24, 256, generator=gen) * 0.02)
    rows.append(torch.randn(8, 256, generator=gen).abs() * 0.3)            # all non-negative groups
    rows.append(-torch.randn(8, 256, generator=gen).abs() * 0.3)           # all negative
    rows.append(torch.zeros(2, 256))                                       # all-zero rows
    rows.append(torch.full((2, 256), 0.125))                               # constant positive
    rows.append(torch.full((2, 256), -0.5))                                # constant negative
    r = torch.randn(4, 256, generator=gen)
    r[:, 32:64] = 0.0                                                      # one zero group
    r[:, 64:96] = 3.0                                                      # one const group
    rows.append(r)
    rows.append(torch.randn(6, 256, generator=gen) * 50.0)                 # large magnitudes
    rows.append(torch.randn(6, 256, generator=gen) * 1e-6)                 # tiny magnitudes
    r = torch.randn(4, 256, generator=gen) * 0.01
    r[:, ::17] = 1.5                                                       # outliers
    rows.append(r)
    x = torch.cat(rows, 0).contiguous()
    out = {"x": x.numpy()}
    for q_type in TYPES:
        bits, clamp, smq, gs, sgs, gtype, qtype_ = qu.GGML_QUANT_SIZES[q_type]
        for variant in ("ieee", "raw"):
            set_sqrt(variant == "ieee")
            qz = qu.Quantizer()
            qz.configure(bits=bits, scale_maxq=smq, super_group_size=sgs, group_size=gs, group_type=gtype,
                         rmin=-1.0, rdelta=0.1, nstep=20)
            d, sq, dmin, zq = qz.get_scale_and_zero(x.clone(), q_type)
            set_sqrt(False)
            out[f"{q_type.name}_{variant}_d"] = d.numpy().view(np.uint16)
            out[f"{q_type.name}_{variant}_sq"] = sq.numpy()
            out[f"{q_type.name}_{variant}_dmin"] = dmin.numpy().view(np.uint16)
            out[f"{q_type.name}_{variant}_zq"] = zq.numpy()
    np.savez_compressed(os.path.join(HERE, "search_edge.npz"), **out)


def rtn_case():
    """Quantizer._quant_non_block_module (quantizer.py:278-330), fp32 weights."""
    gen = torch.Generator().manual_seed(11)
    W = torch.randn(40, 512, generator=gen) * 0.02
    out = {"W": W.numpy()}

    class _Self:
        quantizer_kwargs = dict(quant_scale="absmax", rmin=-1.0, rdelta=0.1, nstep=20)

    for q_type in TYPES:
        set_sqrt(True)
        five = RefQuantizer._quant_non_block_module(_Self(), W.clone(), q_type)
        set_sqrt(False)
        for k, v in np5(five).items():
            out[f"{q_type.name}_ieee_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "rtn.npz"), **out)


def validate_oracle_large():
    """Not stored: oracle vs reference on a 1024x1024 layer, all five types; stats go to golden_stats.json."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import oracle as orc
    gen = torch.Generator().manual_seed(3)
    d_row = d_col = 1024
    W = torch.randn(d_row, d_col, generator=gen) * 0.03
    xs = correlated_x(gen, 8, 512, d_col)
    stats = {}
    for q_type in TYPES:
        row = {}
        for variant in ("ieee", "raw"):
            set_sqrt(variant == "ieee")
            five, U, _ = run_ref_gptq(W, xs, q_type)
            set_sqrt(False)
            o = orc.gptq_step(W.numpy(), U.numpy(), int(q_type))
            ref = np5(five)
            bad_rows = np.zeros(d_row, bool)
            for k, got in zip(["qweight", "d", "sq", "dmin", "zq"],
                              [o[0], o[1].view(np.uint16), o[2], o[3].view(np.uint16), o[4]]):
                bad_rows |= (got != ref[k]).reshape(d_row, -1).any(1)
            row[variant] = {"rows": d_row, "mismatching_rows": int(bad_rows.sum())}
        stats[q_type.name] = row
        print(q_type.name, row, flush=True)
    return stats


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]


if __name__ == "__main__":
    torch.set_num_threads(8)
    b1_case("a", d_row=48, d_col=512, seed=0)
    b1_case("b", d_row=20, d_col=768, seed=1, dead_col=5)
    variants_case("a", d_row=24, d_col=512, seed=5)
    search_case()
    rtn_case()
    stats = {"torch": torch.__version__, "threads": torch.get_num_threads(),
             "reference": "IST-DASLab/gptq-gguf-toolkit@7a38bc5"}
    if "--no-large" not in sys.argv:
        stats["oracle_vs_reference_1024x1024"] = validate_oracle_large()
    stats["files"] = {f: sha(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if f.endswith(".npz")}
    json.dump(stats, open(os.path.join(HERE, "golden_stats.json"), "w"), indent=1)
    print(json.dumps(stats, indent=1))
