"""world_size = 2 over gloo, the WHOLE driver (oracle behind the host logic, CPU): calibration sequences split over the ranks
(quant.py:177-179 of the reference), Hessian all-reduce (gptq.py:131-132), row-sharded column loops + all-gather, results dealt
out over the ranks (spread_emission).  Checked: both ranks end with bit-identical model weights (pass 2 of every block runs on the
all-gathered dequantised weights), every module is emitted exactly once, and the run agrees with the single-rank run of the same 8 sequences at the
statistical boundary B3 (the average of two half-Hessians rounds differently from one running average).
The NCCL twin of this test with the real kernels is tests/test_gpu_multi.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from tests import _oracle_backend as ob
from gptq_gguf_toolkit_b200 import gptq as G, quantizer as Q
G.ops = ob; Q.ops = ob
from gptq_gguf_toolkit_b200.quant import build_quant_config
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
REGEX = r".*layers.*((q|k|v|o|gate|up|down)_proj)$"

def model():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).float().eval()

g = torch.Generator().manual_seed(1)
seqs = [torch.randint(0, 512, (1, 96), generator=g) for _ in range(8)]

def run(my):
    m = model()
    q = Q.Quantizer(m, data_loader=[([], {{"input_ids": t}}) for t in my], quantizable_modules=REGEX,
                    quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                          static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False),
                    pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                    quant_non_block_modules=True, device="cpu", save_dir=None, keep_results=True, calibration_batch_size=2)
    q.quantize(build_quant_config("Q4_K", None))
    return m, q

per = len(seqs) // world
m2, q2 = run(seqs[rank * per:(rank + 1) * per])
keys = [None] * world
dist.all_gather_object(keys, sorted(q2.results))      # spread_emission: the modules' results are dealt out over the ranks
res = {{"results_partitioned_over_ranks": len(set(sum(keys, []))) == 2 * 7 + 2 and sum(len(k) for k in keys) == 2 * 7 + 2
       and all(len(k) == (2 * 7 + 2) // world for k in keys)}}
same = True
for n, p in m2.named_parameters():
    ref = p.data.clone()
    dist.broadcast(ref, src=0)
    same = same and bool(torch.equal(ref, p.data))
flag = torch.tensor([1 if same else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
res["weights_identical_on_all_ranks"] = bool(flag.item())
saved = (Q._world, Q._rank, Q._dist_on, G.HessianAccumulator.all_reduce)
Q._world, Q._rank, Q._dist_on = (lambda: 1), (lambda: 0), (lambda: False)
G.HessianAccumulator.all_reduce = lambda self: setattr(self, "synced", True)
m1, q1 = run(seqs)
Q._world, Q._rank, Q._dist_on, G.HessianAccumulator.all_reduce = saved
eq = tot = 0
rtn_ok = 1
for name, r2 in q2.results.items():          # this rank's share of the modules against the single-rank run
    a, b = q1.results[name]["qweight"], r2["qweight"]
    eq += int((a == b).sum()); tot += a.numel()
    if name in ("model.embed_tokens", "lm_head"):      # RTN modules do not depend on calibration data at all
        rtn_ok = rtn_ok and int(torch.equal(a, b))
acc = torch.tensor([eq, tot, 1 - rtn_ok], dtype=torch.float64)
dist.all_reduce(acc)
res["code_match_rate_vs_single_rank"] = float(acc[0] / acc[1])
res["rtn_identical"] = bool(acc[2] == 0)
dist.barrier()
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
"""


def test_world_size_2_gloo_whole_driver(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    print(res)
    assert res["results_partitioned_over_ranks"] is True
    assert res["weights_identical_on_all_ranks"] is True
    assert res["rtn_identical"] is True
    assert res["code_match_rate_vs_single_rank"] > 0.9, res
