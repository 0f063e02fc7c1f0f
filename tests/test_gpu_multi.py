"""GPU, world_size = 2 over NCCL (skipped on a one-GPU box): the multi-rank path of the driver with the real kernels.

  * row sharding + all-gather (`Quantizer._sharded_gptq`), on the main stream and on a side stream (the deferred-tail
    variant), is bit-identical to the unsharded launch: rows of a GPTQ problem are independent given U;
  * a full driver run with the calibration sequences split over the ranks (quant.py:177-179 of the reference) leaves
    IDENTICAL weights on every rank (pass 2 of every block runs on the all-gathered dequantised weights), deals the results
    out over the ranks (spread_emission: every module emitted exactly once), and agrees with the single-rank run up to the rounding of the Hessian all-reduce (gptq.py:131-132 —
    boundary B3, statistical: the summation order of H differs).
"""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from gptq_gguf_toolkit_b200 import ops, gptq as G, quantizer as Q
from gptq_gguf_toolkit_b200.quant import build_quant_config

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{{local}}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", init_method="env://", device_id=dev)
world = dist.get_world_size()
res = {{}}

# ---- 1. row sharding == unsharded, main stream and side stream, ragged row count, two types -------------------
rng = np.random.default_rng(0)
ok = True
for qt, rows in ((12, 200), (14, 96), (10, 520)):
    W = torch.from_numpy((rng.standard_normal((rows, 1024)) * 0.05).astype(np.float32)).to(dev)
    Un = np.triu(rng.standard_normal((1024, 1024)) * 0.01) + np.eye(1024)
    U = torch.from_numpy(Un.astype(np.float32)).to(dev)
    qz = Q.Quantizer(None, [], "", dict(block_size=128), [], [], "", None)
    ref = ops.gptq_quantize(W.clone(), U, qt, wdeq_dtype=torch.bfloat16)[:7]
    a = qz._sharded_gptq(W.clone(), U, qt, torch.bfloat16, rank, world)
    side = torch.cuda.Stream()
    b = qz._sharded_gptq(W.clone(), U, qt, torch.bfloat16, rank, world, side)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ok = ok and all(torch.equal(x, y) for x, y in zip(a, ref)) and all(torch.equal(x, y) for x, y in zip(b, ref))
res["sharding_bit_exact"] = bool(ok)

# ---- 2. the whole driver ------------------------------------------------------------------------------------
REGEX = r".*layers.*((q|k|v|o|gate|up|down)_proj)$"
def model():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=768, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).to(dev, torch.bfloat16).eval()

g = torch.Generator().manual_seed(1)
seqs = [torch.randint(0, 512, (1, 128), generator=g).to(dev) for _ in range(16)]     # 2048 tokens > d_col 768

def run(my_seqs, **kw):
    m = model()
    q = Q.Quantizer(m, data_loader=[([], {{"input_ids": t}}) for t in my_seqs], quantizable_modules=REGEX,
                    quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                          static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                    pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                    quant_non_block_modules=True, device=dev, save_dir=None, keep_results=True,
                    calibration_batch_size=2, **kw)
    q.quantize(build_quant_config("Q4_K", None))
    torch.cuda.synchronize()
    return m, q

per = len(seqs) // world
m2, q2 = run(seqs[rank * per:(rank + 1) * per])                                  # default: eager chains + deferred tail
m3, q3 = run(seqs[rank * per:(rank + 1) * per], defer_last_layer=False, overlap_prepare=False)
res["deferred_tail_used"] = bool(q2._split_ok)
keys = [None] * world
dist.all_gather_object(keys, sorted(q2.results))      # spread_emission: the modules' results are dealt out over the ranks
res["results_partitioned_over_ranks"] = (len(set(sum(keys, []))) == 3 * 7 + 2 and sum(len(k) for k in keys) == 3 * 7 + 2
                                          and all(len(k) >= (3 * 7 + 2) // world for k in keys))
# every rank must hold the same weights afterwards, and the scheduling options must be bit-neutral at world 2 too
same = True
for (n, p), (_, p3) in zip(m2.named_parameters(), m3.named_parameters()):
    ref = p.data.clone()
    dist.broadcast(ref, src=0)
    same = same and bool(torch.equal(ref, p.data)) and bool(torch.equal(p.data, p3.data))
flag = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
res["weights_identical_on_all_ranks"] = bool(flag.item())

# single-rank run of the same job inside this process (collectives switched off), all 8 sequences
saved = (Q._world, Q._rank, Q._dist_on, G.HessianAccumulator.all_reduce)
Q._world, Q._rank, Q._dist_on = (lambda: 1), (lambda: 0), (lambda: False)
G.HessianAccumulator.all_reduce = lambda self: setattr(self, "synced", True)
m1, q1 = run(seqs)
Q._world, Q._rank, Q._dist_on, G.HessianAccumulator.all_reduce = saved
eq = tot = 0
num = den = 0.0
for name, r2 in q2.results.items():          # this rank's share of the modules against the single-rank run
    a, b = q1.results[name]["qweight"], r2["qweight"]
    eq += int((a == b).sum()); tot += a.numel()
    w1, w2 = m1.get_submodule(name).weight.data.float(), m2.get_submodule(name).weight.data.float()
    num += float((w1 - w2).pow(2).sum()); den += float(w1.pow(2).sum())
acc = torch.tensor([eq, tot, num, den], dtype=torch.float64, device=dev)
dist.all_reduce(acc)
res["code_match_rate_vs_single_rank"] = float(acc[0] / acc[1])
res["rel_weight_diff_vs_single_rank"] = float((acc[2] / acc[3]) ** 0.5)
res["non_invertible"] = q2.non_invertible_modules()
dist.barrier()
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
"""


def test_world_size_2_nccl_driver_and_row_sharding(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    print(res)
    assert res["sharding_bit_exact"] is True
    assert res["deferred_tail_used"] is True
    assert res["results_partitioned_over_ranks"] is True
    assert res["weights_identical_on_all_ranks"] is True
    assert res["non_invertible"] == []
    # B3-class agreement with the single-rank run: only the rounding of the Hessian average differs
    # (measured on 2xB200: 94 % identical codes / 6 % relative weight difference with only 512 calibration tokens, where
    # the d_col = 768 Hessian is rank deficient and rests on the damping term; the test uses 2048 tokens)
    assert res["code_match_rate_vs_single_rank"] > 0.9, res
    assert res["rel_weight_diff_vs_single_rank"] < 0.1, res
