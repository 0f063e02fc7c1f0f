#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: wall-clock seconds to GPTQ-quantise a random-init
Llama-3-8B (bf16) to uniform Q4_K with 128 x 2048 synthetic calibration tokens, on N B200s.

    python bench.py --gpus 1 --steps 1 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

One "step" = one complete quantisation of the model through gptq_gguf_toolkit_b200.quantizer.Quantizer
(capture block inputs, per block: forward pass 1 + Hessians, Cholesky chain, fused column loop, forward
pass 2; embed_tokens / lm_head RTN), weights restored to the pristine random init before every step.
  value : weights and token ids resident in HBM when the timed region starts, results left in HBM.
  e2e   : the same call with HOST buffers: pristine weights are copied from pinned host memory inside the
          timed region and every result (five tensors + GGUF bytes per module) is read back to pinned host memory.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "llama3_8b_q4k_quantize_wall_clock_s"
# ncu --set full, one launch of exact_update_kernel: dram__bytes_read.sum + dram__bytes_write.sum (round 2's capture,
# profiles/r02/r02ag_ncu_details_colloop.csv; round 1's, profiles/r01_final2_ncu_details_exact_update_kernel.csv: 53.2 + 2.3 MB)
EXACT_UPDATE_DRAM_BYTES = 54.7e6
EXACT_UPDATE_TRAFFIC_NOTE = ("ncu --set full of ONE launch (profiles/r02/r02ag_ncu_details_colloop.csv): o_proj shape, "
                             "4096 rows x 11 windows of 256 columns, 154.0 us, 53.2 MB read + 1.5 MB written; algorithmic "
                             "bytes of that launch 99.4 MB (W window read + written 92.3, E 4.2, U slab 2.9): the window was "
                             "written by the previous launch and is still in the 126 MB L2 (the write-back happens later); a constant "
                             "of that capture, not a measurement of this run")
REGEX = r".*layers.*((q|k|v|o|gate|up|down)_proj)$"

WORKLOADS = {
    # BASELINE.json configs[1]
    "llama3-8b": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                      num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, n_seq=128, seq_len=2048,
                      dtype="bfloat16"),
    # development only (NOT the reported metric): same layer shapes, fewer blocks / sequences
    "llama3-8b-dev": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=2, num_attention_heads=32,
                          num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, n_seq=32, seq_len=2048,
                          dtype="bfloat16"),
    # profiling only: ONE block (about 2300 kernel launches per step; an ncu launch list costs ~0.17 s per launch)
    "llama3-8b-dev1": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=1, num_attention_heads=32,
                           num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, n_seq=16, seq_len=2048,
                           dtype="bfloat16"),
    # BASELINE.json configs[0] (plumbing)
    "tiny": dict(hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, vocab_size=1024, max_position_embeddings=256, n_seq=8, seq_len=128,
                 dtype="float32"),
}


def layer_shapes(w):
    h, i = w["hidden_size"], w["intermediate_size"]
    kv = h // w["num_attention_heads"] * w["num_key_value_heads"]
    return [("q_proj", h, h), ("k_proj", kv, h), ("v_proj", kv, h), ("o_proj", h, h),
            ("gate_proj", i, h), ("up_proj", i, h), ("down_proj", h, i)]


def algorithmic_work(w):
    """rank-k flops d_row*d_col*(d_col-128) per layer; Hessian flops 2*T*d_col^2 per layer as the
    reference computes it (7 per block)."""
    T = w["n_seq"] * w["seq_len"]
    rk = sum(r * c * (c - 128) for _, r, c in layer_shapes(w)) * w["num_hidden_layers"]
    hs = sum(2 * T * c * c for _, r, c in layer_shapes(w)) * w["num_hidden_layers"]
    return rk, hs


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tf=1400.0, tf_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =================================================================================================
# reference arm / cpu_baseline: the reference's CPU implementation of the hot path.
# The reference is pure Python on torch and cannot travel to the GPU box, so this is the oracle PORT:
# torch CPU ops for exactly the library calls the reference makes (H.addmm_, cholesky / cholesky_inverse /
# cholesky(upper)), the C restatement (oracle/gq_oracle.c, OpenMP over rows) for the column loop.
# =================================================================================================
# The reference's OWN functions timed in the build container (tests/golden/check_oracle_vs_reference_large.py, 8 threads,
# bit-identical outputs): GPTQ.step of the reference against the oracle port on the same (W, U).  The port is the kinder
# baseline; the reference itself cannot travel to the GPU box (pure Python on /root/reference).
REFERENCE_MEASURED = {
    "where": "build container (no GPU), 8 CPU threads, tests/golden/check_oracle_vs_reference_large.py",
    "gptq_step_s": {"1024x2048": {"reference": 1.20, "oracle_port": 0.29}, "2048x4096": {"reference": 3.51, "oracle_port": 1.61}},
    "port_is_kinder_by": "2.2-4.1x on the column loop",
}


_CPU_BLOCK = {}


def _cpu_block(w):
    """ONE decoder block of the workload's shape on the CPU, in the model's dtype (built once per process: constructing and
    initialising 218 M parameters costs more than the sampled forward)."""
    key = (w["hidden_size"], w["intermediate_size"], w["dtype"])
    if key not in _CPU_BLOCK:
        from transformers import LlamaConfig, LlamaModel
        cfg = LlamaConfig(**{k: w[k] for k in ("hidden_size", "intermediate_size", "num_attention_heads", "num_key_value_heads",
                                               "max_position_embeddings")}, num_hidden_layers=1, vocab_size=1024)
        old = torch.get_default_dtype()
        torch.set_default_dtype(getattr(torch, w["dtype"]))
        try:
            _CPU_BLOCK[key] = LlamaModel(cfg).eval()
        finally:
            torch.set_default_dtype(old)
    return _CPU_BLOCK[key]


def cpu_reference_sample(w, threads: int):
    """Times a bounded sample (~10-30 s of CPU work, depending on the host's cores) and extrapolates each phase by its algorithmic
    work to the whole model.  Returns (whole_model_hot_path_seconds, detail).  Sample: 16 sequences of H.addmm_ and one Cholesky chain at every distinct
    d_col, the column loop on WHOLE-WIDTH slabs (2048 rows at the narrow width, 512 rows at the widest)."""
    from oracle import oracle as orc
    import numpy as np
    torch.set_num_threads(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    g = torch.Generator().manual_seed(0)
    shapes = layer_shapes(w)
    dcols = sorted({c for _, _, c in shapes})
    L, nblk, nseq = w["seq_len"], w["num_hidden_layers"], w["n_seq"]
    detail, sample_s = {}, {}
    # 1. Hessian: H.addmm_ per calibration sequence (gptq.py:108-112), 8 sequences at every distinct d_col
    t_h = {}
    n_h = min(16, nseq)
    for c in dcols:
        x = torch.randn(L, c, generator=g).to(torch.bfloat16).float()
        H = torch.zeros(c, c)
        H.addmm_(x.T, x, beta=0.0, alpha=2.0)            # warm
        t0 = time.perf_counter()
        for i in range(n_h):
            H.addmm_(x.T, x, beta=(i + 1.0) / (i + 2.0), alpha=2.0 / (i + 2.0))
        t_h[c] = (time.perf_counter() - t0) / n_h
    hess = sum(t_h[c] for _, _, c in shapes) * nseq * nblk
    detail["hessian_s_per_seq"] = {str(k): round(v, 4) for k, v in t_h.items()}
    sample_s["hessian"] = round(sum(t_h.values()) * n_h, 2)
    # 2. Cholesky chain, measured at every distinct d_col (gptq.py:305-324)
    t_p = {}
    for c in dcols:
        x = torch.randn(c // 4, c, generator=g)          # SPD, well conditioned; the factorisation's cost does not depend on the values
        H = (x.T @ x) / c + 0.5 * torch.eye(c)
        t0 = time.perf_counter()
        Hi = torch.cholesky_inverse(torch.linalg.cholesky(H))
        torch.linalg.cholesky(Hi, upper=True)
        t_p[c] = time.perf_counter() - t0
    prep = sum(t_p[c] for _, _, c in shapes) * nblk
    detail["prepare_s"] = {str(k): round(v, 3) for k, v in t_p.items()}
    sample_s["prepare"] = round(sum(t_p.values()), 2)
    # 3. column loop (gptq.py:146-295): rows are independent, so time(d_row, d_col) = d_row * t_row(d_col); t_row is measured
    #    on a slab of the layer's FULL width at every distinct d_col
    t_row = {}
    for c in dcols:
        rows = 2048 if c <= 4096 else 512
        rng = np.random.default_rng(c)
        W = (rng.standard_normal((rows, c)) * 0.02).astype(np.float32)
        U = np.triu(rng.standard_normal((c, c)).astype(np.float32) * 0.01) + np.eye(c, dtype=np.float32)
        t0 = time.perf_counter()
        orc.gptq_step(W, U, 12)
        t_row[c] = (time.perf_counter() - t0) / rows
        detail.setdefault("step_slab_rows", {})[str(c)] = rows
    step = sum(r * t_row[c] for _, r, c in shapes) * nblk
    detail["step_s_per_row"] = {str(k): round(v, 5) for k, v in t_row.items()}
    sample_s["step"] = round(sum(t_row[c] * detail["step_slab_rows"][str(c)] for c in dcols), 2)
    # 4. the two forward passes per block the reference driver runs on its device (quantizer.py:150-151, 161-172; here: the CPU), one
    #    sequence per forward, in the model's dtype: ONE decoder block of the workload's shape on a piece of a sequence, scaled
    #    linearly to 2 passes x n_seq sequences x all blocks (the attention's quadratic term is under-counted: kinder to the CPU)
    t_f, L_f = None, min(L, 512)
    try:
        blk = _cpu_block(w)
        ids = torch.randint(0, 1024, (1, L_f), generator=g)
        with torch.no_grad():
            blk(input_ids=ids)
            t0 = time.perf_counter()
            blk(input_ids=ids)
            t_f = time.perf_counter() - t0
    except Exception as ex:  # noqa: BLE001
        detail["forward_sample_failed"] = f"{type(ex).__name__}: {str(ex)[:120]}"
    fwd = (t_f or 0.0) * (L / L_f) * nseq * 2 * nblk
    detail["forward_s_per_block_and_piece"] = {"tokens": L_f, "seconds": round(t_f, 4) if t_f else None, "dtype": w["dtype"]}
    sample_s["forwards"] = round(2 * (t_f or 0.0), 2)
    # 5. embed_tokens / lm_head round-to-nearest (quantizer.py:278-330): rows are independent, a 1024-row slab of each
    t0 = time.perf_counter()
    orc.rtn_quantize((np.random.default_rng(7).standard_normal((1024, w["hidden_size"])) * 0.02).astype(np.float32), 12)
    t_r = time.perf_counter() - t0
    rtn = t_r * (2 * w["vocab_size"] / 1024)
    sample_s["rtn"] = round(t_r, 2)
    hot = hess + prep + step
    total = hot + fwd + rtn
    detail.update(hessian_s=round(hess, 1), prepare_s_total=round(prep, 1), step_s=round(step, 1), forwards_s=round(fwd, 1), rtn_s=round(rtn, 1),
                  hot_path_only_s=round(hot, 1), sample_seconds=sample_s, sample_total_s=round(sum(sample_s.values()), 2))
    return total, detail


CPU_SAMPLE_NOTE = ("per phase: H.addmm_ of 16 sequences of 2048 tokens at each d_col, one Cholesky chain (cholesky, cholesky_inverse, "
                   "cholesky upper) at each d_col, the column loop on whole-width slabs (2048 rows x 4096, 512 rows x 14336), one decoder "
                   "block forward on 512 tokens in the model dtype, RTN of a 1024-row slab; each scaled by its algorithmic work to the whole "
                   "job (32 blocks x 7 projections x 128 sequences, 2 forward passes per block, embed_tokens + lm_head) -- EXTRAPOLATED; "
                   "detail.hot_path_only_s is the hot path alone (Hessian + Cholesky chain + column loop); oracle PORT of the reference "
                   "(see reference_measured for the reference's own functions)")


def bench_config(args, w, world, qdesc, main_mode):
    """The `config` object of the JSON line -- the SAME for both arms (the reference arm runs the workload the own arm names)."""
    return {"workload": f"{args.workload}: random-init {w['dtype']} Llama ({w['num_hidden_layers']} blocks, d_model {w['hidden_size']}), "
                        f"{w['n_seq']} calib seqs x {w['seq_len']}, {qdesc}, {main_mode} mode",
            "calibration_batch": args.batch, "l2": "inputs (16 GB weights + 2 GB activations) exceed the 126 MB L2; no flush needed",
            "parallelism": f"dp{world} over calibration sequences + row-sharded quantisation" if world > 1 else "single GPU"}


def run_reference_arm(args, w):
    """`value` = the metric (seconds for the whole job) EXTRAPOLATED from the bounded sample; `ms_per_step` = the time one timed
    step (one sample) really took, so that steps x ms_per_step is the run's own duration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(max(0, args.warmup - 2)):      # the sample is deterministic CPU work; one warm-up is plenty
        cpu_reference_sample(w, threads)
    vals, walls, detail = [], [], None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, detail = cpu_reference_sample(w, threads)
        walls.append(time.perf_counter() - t0)
        vals.append(v)
    v = sum(vals) / len(vals)
    main_mode = "fast" if args.mode == "fast" else "exact"
    qdesc = f"uniform {args.qtype}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, w, args.gpus, qdesc, main_mode),
        "value_note": "whole-job seconds EXTRAPOLATED from the bounded CPU sample each step times (ms_per_step is the sample's own wall "
                      "time); runs on rank 0's host cores only, whatever --gpus says",
        "cpu_baseline": {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": CPU_SAMPLE_NOTE, "extrapolated": True,
                         "sample_seconds": detail.get("sample_seconds"), "detail": detail, "reference_measured": REFERENCE_MEASURED},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# =================================================================================================
# own arm
# =================================================================================================
def build_model(w, device, seed=0):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(**{k: w[k] for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                           "num_key_value_heads", "vocab_size", "max_position_embeddings")},
                      tie_word_embeddings=False)
    dtype = getattr(torch, w["dtype"])
    torch.manual_seed(seed)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    return model.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="llama3-8b", choices=list(WORKLOADS))
    ap.add_argument("--qtype", type=str, default="Q4_K")
    ap.add_argument("--batch", type=int, default=8, help="calibration sequences per block forward")
    ap.add_argument("--mode", type=str, default="both", choices=["exact", "fast", "both"],
                    help="exact: bit-identical fp32 rank-k (headline); fast: tcgen05 split-fp16 rank-k; both: headline exact + one fast step")
    ap.add_argument("--bit-width-configuration", type=str, default=None,
                    help="JSON file {projection name: Q2_K..Q6_K} (the reference's --bit_width_configuration, quant.py:203-217); "
                         "BASELINE.json configs[2]: profiles/configs/mixed_q2k_q6k.json")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="ablation: Cholesky chains on the main stream")
    ap.add_argument("--overlap", type=str, default="eager", choices=["eager", "staged"],
                    help="side-stream Cholesky chains: overlap the column loops too (eager) or run before them (staged)")
    ap.add_argument("--no-early-exit", action="store_true", help="ablation: run the full block in forward pass 1")
    ap.add_argument("--no-defer", action="store_true", help="ablation: no split of forward pass 2 at the last quantised layer")
    ap.add_argument("--no-fused-forward", action="store_true", help="ablation: HF's eager RMSNorm / rotary / SiLU*up kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compare-left", type=int, default=0, metavar="K",
                    help="ablation: K extra steps with the exact LEFT-looking schedule (one launch per layer), same process")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import torch.distributed as dist
    from gptq_gguf_toolkit_b200 import ops
    from gptq_gguf_toolkit_b200.data_utils import synthetic_tokens
    from gptq_gguf_toolkit_b200.quant import build_quant_config
    from gptq_gguf_toolkit_b200.quantizer import PhaseTimer, Quantizer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    model = build_model(w, device)
    quant_config = build_quant_config(args.qtype, args.bit_width_configuration)
    qdesc = f"uniform {args.qtype}" if args.bit_width_configuration is None else \
        "mixed " + ", ".join(f"{k}:{v.name}" for k, v in sorted(quant_config.items()))
    names = [n for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)] + ["model.embed_tokens"]
    mods = {n: model.get_submodule(n) for n in names}
    pristine = {n: m.weight.data.clone() for n, m in mods.items()}            # HBM copy (value runs)
    tokens = synthetic_tokens(w["n_seq"], w["seq_len"], w["vocab_size"], seed=1)
    per = len(tokens) // world
    tokens = [t.to(device) for t in tokens[rank * per:(rank + 1) * per]]      # quant.py:177-179 slicing
    loader = [([], {"input_ids": t}) for t in tokens]
    host_w = None

    def restore_from_hbm():
        for n, m in mods.items():
            m.weight.data = pristine[n].clone()

    main_mode = "fast" if args.mode == "fast" else "exact"
    checks = {}
    copy_stream = [None]

    def result_crcs(results):
        import zlib
        return {n: zlib.crc32(memoryview(o["packed"].numpy()).cast("B")) & 0xFFFFFFFF for n, o in results.items() if "packed" in o}

    def result_checksums(per):
        """CRC-32 of the packed GGUF bytes of every module as they arrived in host memory (outside the timed region).
        embed_tokens / lm_head (RTN, no calibration data) must be identical for every N; the GPTQ modules depend on the
        rounding of the Hessian all-reduce (gptq.py:131-132 averages per-rank sums), so their bytes agree across N only
        statistically -- in the reference as well."""
        import zlib
        allc = 0
        for n in sorted(per):
            allc = zlib.crc32(per[n].to_bytes(4, "little"), allc)
        return {"crc32_embed_tokens": per.get("model.embed_tokens"), "crc32_lm_head": per.get("lm_head"),
                "crc32_block0_q_proj": per.get("model.layers.0.self_attn.q_proj"),
                "crc32_of_all_module_crcs": allc & 0xFFFFFFFF, "modules": len(per)}

    def weights_identical_on_all_ranks():
        """After a step every rank's model holds the dequantised weights it was given by the all-gathers: compare an
        order-independent 64-bit sum of the raw bits of every quantised module across the ranks."""
        tot = torch.zeros(1, dtype=torch.int64, device=device)
        for n, m in mods.items():
            tot += m.weight.data.view(torch.int16).to(torch.int64).sum()
        lo, hi = tot.clone(), tot.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool((lo == hi).item())

    def one_step(e2e: bool, mode: str = None, defer: bool = None):
        timer = PhaseTimer(True)
        q = Quantizer(model, data_loader=loader, quantizable_modules=REGEX,
                      quantizer_kwargs=dict(rel_damp=0.01, block_size=128, act_order=False, quant_scale="absmax",
                                            static_groups=False, rmin=-1.0, rdelta=0.1, nstep=20, verbose=False,
                                            mode=mode or main_mode),
                      pre_block_modules=["model.embed_tokens"], block_modules="model.layers", post_block_modules=["lm_head"],
                      quant_non_block_modules=True, device=device, save_dir=None, keep_results=e2e,
                      calibration_batch_size=args.batch, timer=timer, overlap_prepare=False if args.no_overlap else args.overlap,
                      early_exit_pass1=not args.no_early_exit, defer_last_layer=(not args.no_defer) if defer is None else defer,
                      fused_forward_ops=not args.no_fused_forward)
        if not e2e:
            restore_from_hbm()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = ops.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        h2d = 0
        hooks = []
        if e2e:
            # Host weights stream in on a copy stream, in the order the quantiser needs them; a forward pre-hook per decoder block
            # (and on embed_tokens) makes the compute stream wait for that block's weights only, so the 16 GB host -> device copy
            # runs underneath the first blocks instead of in front of them.
            copy = copy_stream[0] = copy_stream[0] or torch.cuda.Stream()      # ONE stream for all steps: the allocator pools per stream
            main = torch.cuda.current_stream()
            copy.wait_stream(main)
            order = ["model.embed_tokens"] + [n for n in mods if n.startswith("model.layers.")] + [n for n in mods if n == "lm_head"]
            events = {}
            with torch.cuda.stream(copy):
                for n in order:
                    t = host_w[n].to(device, non_blocking=True)
                    t.record_stream(main)
                    mods[n].weight.data = t
                    ev = torch.cuda.Event()
                    ev.record(copy)
                    events[n] = ev
                    h2d += host_w[n].numel() * host_w[n].element_size()
            layers = model.model.layers

            def waiter(names):
                def _h(_m, _a):
                    for nm in names:
                        torch.cuda.current_stream().wait_event(events[nm])
                return _h
            hooks.append(model.model.embed_tokens.register_forward_pre_hook(waiter(["model.embed_tokens"])))
            for i, blk in enumerate(layers):
                names_i = [n for n in mods if n.startswith(f"model.layers.{i}.")]
                if i == len(layers) - 1:
                    names_i = names_i + [n for n in mods if n == "lm_head"]
                hooks.append(blk.register_forward_pre_hook(waiter(names_i)))
        q.quantize(quant_config)
        ev1.record()
        torch.cuda.synchronize()
        for h in hooks:
            h.remove()
        if world > 1:
            dist.barrier()
        secs = ev0.elapsed_time(ev1) / 1e3
        if world > 1:
            t = torch.tensor([secs], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        d2h = 0
        if e2e:
            for obj in q.results.values():
                d2h += sum(v.numel() * v.element_size() for v in obj.values() if isinstance(v, torch.Tensor))
            per = result_crcs(q.results)
            if world > 1:      # the results are dealt out over the ranks (Quantizer.spread_emission)
                t = torch.tensor([d2h], device=device, dtype=torch.int64)
                dist.all_reduce(t)
                d2h = int(t.item())
                parts = [None] * world
                dist.all_gather_object(parts, per)
                per = {k: v for part in parts for k, v in part.items()}
            if rank == 0:
                checks.clear()
                checks.update(result_checksums(per))
            q.results.clear()
        bad = q.non_invertible_modules()
        return secs, timer.totals(), ops.launch_count() - l0, h2d, d2h, bad

    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    times, phases, launches, bad = [], {}, 0, []
    # CUDA events around every launch of the column-loop kernels (gq_profile_*, recorded on the launching stream, read
    # after the timed region): the live duration of the dominant kernel -- exact_update_kernel, the rank-k trailing
    # update of the exact schedule -- for the roofline object.
    prof, prof_on = None, main_mode == "exact"
    if prof_on:
        ops.profile_enable(True)
    for _ in range(args.steps):
        secs, ph, nl, _, _, bad = one_step(False)
        times.append(secs)
        launches += nl
        for k, v in ph.items():
            phases[k] = phases.get(k, 0.0) + v / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = sum(times) / len(times)

    if prof_on:
        prof = ops.profile_read()
        ops.profile_enable(False)
        for k in prof:
            prof[k] = prof[k] / args.steps

    # ---- optional extras.  The headline numbers above are complete at this point; a failure in an extra (they fail the same
    # way on every rank: same code, same shapes) is recorded in "notes" instead of losing the whole JSON line.
    notes = []

    def guarded(what, fn):
        try:
            return fn()
        except Exception as ex:  # noqa: BLE001
            notes.append(f"{what} failed: {type(ex).__name__}: {str(ex)[:200]}")
            return None

    def run_left():
        one_step(False, "exact_left")
        lt = [one_step(False, "exact_left") for _ in range(args.compare_left)]
        return {"value": sum(t[0] for t in lt) / len(lt), "unit": "s", "steps": args.compare_left,
                "phases_s": {k: round(v, 4) for k, v in sorted(lt[-1][1].items())},
                "note": "GQ_MODE_EXACT_LEFT: one left-looking launch per layer; bit-identical outputs"}

    def run_fast():
        # one extra step with the tcgen05 rank-k path; CUDA events around every rank-k GEMM launch (gq_profile_*)
        one_step(False, "fast")
        ops.profile_enable(True)
        try:
            secs_f, ph_f, _, _, _, _ = one_step(False, "fast")
            pr = ops.profile_read()
            # the same step without the deferred tail: down_proj's launches (57 % of the rank-k flops) then run alone on the
            # main stream instead of underneath the pass-2 forwards, i.e. the kernel's duration without SM contention
            one_step(False, "fast", defer=False)
            pr_alone = ops.profile_read()
        finally:
            ops.profile_enable(False)
        return (secs_f, ph_f, pr, pr_alone)

    def run_e2e():
        nonlocal host_w
        # pinned host copies of the pristine weights (16 GB per rank for Llama-3-8B); every rank must succeed, otherwise all
        # ranks skip the end-to-end step together (a rank that bails out alone would leave the others in a collective)
        ok, why = 1, ""
        try:
            host_w = {}
            for n, t in pristine.items():
                buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                buf.copy_(t)
                host_w[n] = buf
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            ok, why, host_w = 0, f"pinned host allocation failed: {type(ex).__name__}", None
        if world > 1:
            t = torch.tensor([ok], device=device, dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if ok and int(t.item()) == 0:
                why = "pinned host allocation failed on another rank"
            ok = int(t.item())
        if not ok:
            return {"value": None, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "unavailable": why}
        one_step(True)      # warm-up: fills torch's pinned-host cache with the result buffers (cudaHostAlloc is slow)
        secs, _, _, h2d, d2h, _ = one_step(True)
        out = {"value": secs, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": 1}
        if world > 1 and w["dtype"] != "float32":
            out["weights_identical_on_all_ranks"] = weights_identical_on_all_ranks()
        return out

    left = guarded("left-looking ablation", run_left) if (args.compare_left > 0 and main_mode == "exact") else None
    # e2e (a contract key) first; the fast-mode extra only on one GPU: it is an extra of the N = 1 line, and a rank-local
    # failure in an extra step at N > 1 would leave the other ranks waiting in a collective
    e2e = guarded("e2e", run_e2e) if not args.no_e2e else None
    fast = guarded("fast mode", run_fast) if (args.mode == "both" and world == 1) else None

    if rank == 0:
        pk = peaks()
        rk_flops, hs_flops = algorithmic_work(w)
        T_tok = w["n_seq"] * w["seq_len"]
        # executed SYRK work: one Hessian per distinct input (q/k/v, o, gate/up, down), 128 x 256 tiles that touch the upper triangle
        hs_exec = sum(2 * T_tok * c * c * (0.5 + 128.0 / c) for c in (w["hidden_size"], w["hidden_size"], w["hidden_size"], w["intermediate_size"])) * w["num_hidden_layers"]
        t_gptq = phases.get("gptq", 0.0)
        simt_peak = 148 * 128 * 2 * 1.965e9 / 1e12       # fp32 FFMA peak of the SIMT pipes at the max SM clock
        # flops of the launches of exact_update_kernel: sum over super-blocks of 2*d_row*256*(d_col - 256*(sb+1))
        upd_flops = sum(r * c * (c - 256) for _, r, c in layer_shapes(w)) * w["num_hidden_layers"] / world
        if prof is not None and prof["rankk_gemm_ms"] > 0:
            t_upd = prof["rankk_gemm_ms"] * 1e-3
            achieved = upd_flops / t_upd / 1e12
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved / pk["tf"],
                "traffic": EXACT_UPDATE_DRAM_BYTES, "traffic_note": EXACT_UPDATE_TRAFFIC_NOTE,
                "kernel": "exact_update_kernel: rank-256 trailing update of the exact right-looking schedule, W[:, c+256:] -= "
                          "E[:, c:c+256] U[c:c+256, c+256:] as two sequentially rounded 128-term fp32 FMA chains per element",
                "launches_per_step": int(round(prof["rankk_gemm_launches"])), "avg_launch_ms": prof["rankk_gemm_ms"] / max(1, prof["rankk_gemm_launches"]),
                "total_ms_per_step": prof["rankk_gemm_ms"], "algorithmic_flops_per_step": upd_flops,
                "peak_source": pk["source"],
                "simt_fp32_peak_tflops": round(simt_peak, 1), "frac_of_simt_fp32_peak": achieved / simt_peak,
                "panel_kernel": {"name": "gptq_layer_kernel<Q4_K> (scale search + 256 column steps + in-super-block update + GGUF pack, "
                                         "one launch per super-block)", "launches_per_step": int(round(prof["panel_launches"])),
                                 "total_ms_per_step": prof["panel_ms"]},
                "note": ("CUDA events around every launch, on the launching stream, inside the timed steps; "
                         "down_proj's launches run on a side stream concurrently with the pass-2 block forwards, which "
                         "lengthens them.  The contract asks for the fraction of the TENSOR peak; exact mode keeps the "
                         "reference's sequentially rounded fp32 chain (bit-identical packed bytes), which no tensor-core "
                         "instruction reproduces, so the kernel runs on the SIMT FFMA pipes and frac_of_simt_fp32_peak is "
                         "the figure that describes its quality; fast_mode below runs the same update on tcgen05"),
            }
        else:
            n_layer_launch = w["num_hidden_layers"] * 4      # q/k/v stacked, o, gate/up stacked, down
            achieved = (rk_flops / world) / t_gptq / 1e12 if t_gptq > 0 else 0.0
            roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved / pk["tf"],
                        "traffic": None, "kernel": "column-loop launches (panel + rank-k update), whole 'gptq' phase",
                        "launches_per_step": n_layer_launch, "avg_launch_ms": 1e3 * t_gptq / max(1, n_layer_launch),
                        "peak_source": pk["source"]}
        roofline["column_loop_phase"] = {"seconds": t_gptq, "rank_k_tflops_over_whole_phase": (rk_flops / world) / t_gptq / 1e12 if t_gptq > 0 else None}
        hot = sum(phases.get(k, 0.0) for k in ("hessian", "prepare_host", "gptq", "rtn"))   # prepare is nested in prepare_host
        line = {
            "metric": METRIC, "value": value, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, w, world, qdesc, main_mode),
            "roofline": roofline,
            "phases_s": {k: round(v, 4) for k, v in sorted(phases.items())},
            "hot_path_s": round(hot, 4),
            # the reference computes 7 full Hessians per block (2 T d_col^2 flops each); this build executes 4 (shared inputs),
            # upper-triangle tiles only
            "hessian_algorithmic_tflops": round(hs_flops / world / phases["hessian"] / 1e12, 1) if phases.get("hessian") else None,
            "hessian_executed_tflops": round(hs_exec / world / phases["hessian"] / 1e12, 1) if phases.get("hessian") else None,
            "hessian_executed_frac_of_peak": round(hs_exec / world / phases["hessian"] / 1e12 / pk["tf"], 3) if phases.get("hessian") else None,
            "gpu_launches": launches,
            "clocks": clocks,
            "non_invertible_modules": bad,
        }
        if left is not None:
            line["left_looking_schedule"] = left
        if fast is not None:
            secs_f, ph_f, pr, pr_alone = fast
            tfs_alone = (sum(r * c * (c - 256) for _, r, c in layer_shapes(w)) * w["num_hidden_layers"] / world) / \
                (pr_alone["rankk_gemm_ms"] * 1e-3) / 1e12 if pr_alone["rankk_gemm_ms"] > 0 else 0.0
            # the GEMMs cover d_row*d_col*(d_col-256) of the rank-k flops (the first 128 columns' update of each
            # super-block's second half stays in the fused kernel)
            gemm_flops = sum(r * c * (c - 256) for _, r, c in layer_shapes(w)) * w["num_hidden_layers"] / world
            tfs = gemm_flops / (pr["rankk_gemm_ms"] * 1e-3) / 1e12 if pr["rankk_gemm_ms"] > 0 else 0.0
            line["fast_mode"] = {
                "value": secs_f, "unit": "s", "phases_s": {k: round(v, 4) for k, v in sorted(ph_f.items())},
                "note": "GQ_MODE_FAST: the rank-k updates between 256-column super-blocks run as tcgen05 split-fp16 GEMMs "
                        "(csrc/gemm_f16x3.cu: TMA-fed, 128x256 tiles, TMEM double-buffered, three kind::f16 MMAs per product on "
                        "hi/lo operand pairs = 22-bit operands, fp32 accumulation); groups of 2 super-blocks share one trailing "
                        "update (K = 512).  fp32-class accuracy, not bit-identical to the reference (tests: objective within 2e-3)",
                "roofline": {"bound": "tensor", "achieved": tfs, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tfs / pk["tf"],
                             "traffic": 265.8e6,
                             "traffic_note": "ncu --set full of ONE launch (profiles/r02/r02b_ncu_details_gemm_f16x3_kernel.csv: down_proj "
                                             "shape, 148 CTAs, 94.7 us): 172.4 MB read + 93.4 MB written; tensor pipe 72.7 % of active cycles",
                             "kernel": "gemm_f16x3_kernel<256> (rank-256/512 trailing update, 3 fp16 MMAs per product)",
                             "launches": pr["rankk_gemm_launches"], "total_ms": pr["rankk_gemm_ms"],
                             "avg_launch_ms": pr["rankk_gemm_ms"] / max(1, pr["rankk_gemm_launches"]),
                             "algorithmic_flops_per_step": gemm_flops,
                             "executed_fp16_tflops": 3 * tfs, "executed_frac_of_peak": 3 * tfs / pk["tf"],
                             "peak_source": pk["source"],
                             "note": "CUDA events around every launch inside the timed fast-mode step; down_proj's launches run on a "
                                     "side stream underneath the pass-2 block forwards (cuBLAS / cuDNN kernels competing for the SMs), "
                                     "which lengthens them -- uncontended below is the same measurement in a step without that overlap",
                             "uncontended": {"achieved": tfs_alone, "frac": tfs_alone / pk["tf"], "total_ms": pr_alone["rankk_gemm_ms"],
                                             "launches": pr_alone["rankk_gemm_launches"], "executed_fp16_tflops": 3 * tfs_alone,
                                             "executed_frac_of_peak": 3 * tfs_alone / pk["tf"],
                                             "how": "one extra fast-mode step with defer_last_layer=False (all column loops on the main stream)"}},
                "operand_split_ms": pr.get("split_ms"), "operand_split_launches": pr.get("split_launches"),
                "panel_kernel_ms": pr["panel_ms"], "panel_kernel_launches": pr["panel_launches"],
            }
        if e2e is not None:
            line["e2e"] = e2e
            if checks:
                line["checksums"] = dict(checks, note="CRC-32 of the packed GGUF bytes per module, from the e2e step's host copies; "
                                         "the RTN modules (embed_tokens, lm_head) are N-independent, the GPTQ modules depend on the "
                                         "rounding of the Hessian all-reduce (B3 class, as in the reference)")

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu = guarded("cpu_baseline", lambda: cpu_reference_sample(w, threads))
        if cpu is not None:
            v, detail = cpu
            line["cpu_baseline"] = {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": CPU_SAMPLE_NOTE,
                                    "extrapolated": True, "sample_seconds": detail.get("sample_seconds"), "detail": detail,
                                    "reference_measured": REFERENCE_MEASURED}
        if notes:
            line["notes"] = notes
        print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


if __name__ == "__main__":
    main()
