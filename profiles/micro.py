#!/usr/bin/env python
"""Micro-timings of the individual libgq ops at Llama-3-8B layer shapes (CUDA events, 1 GPU).
    python profiles/micro.py [prepare] [gptq] [hessian] [rtn]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gptq_gguf_toolkit_b200 import ops  # noqa: E402


def timed(fn, warm=1, it=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


PHASES = ["rank_update", "search", "finalize", "serial0", "mid_update", "serial1", "emit"]


def phase_clocks(fn, n_cta):
    """Per-phase cycle counters of gptq_layer_kernel (gq_debug_phase_clocks), averaged per CTA."""
    import ctypes as C
    from gptq_gguf_toolkit_b200 import _lib
    lib = _lib.load()
    lib.gq_debug_phase_clocks.argtypes = [C.c_void_p]
    lib.gq_debug_phase_clocks.restype = None
    buf = torch.zeros(8, dtype=torch.int64, device="cuda")
    lib.gq_debug_phase_clocks(C.c_void_p(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.gq_debug_phase_clocks(None)
    c = buf.cpu().tolist()
    tot = sum(c[:7]) or 1
    print("   phase cycles per CTA: " + ", ".join(f"{n} {v / n_cta / 1e3:.0f}k ({100 * v / tot:.0f}%)" for n, v in zip(PHASES, c)), flush=True)


def spd(n, dev="cuda"):
    x = torch.randn(2 * n, n, device=dev)
    H = (x.T @ x) / n
    return H


def main():
    what = set(sys.argv[1:]) or {"prepare", "gptq", "hessian", "rtn"}
    if "small" in what:
        what.add("gptq")
    torch.manual_seed(0)
    if "prepare" in what:
        for n in (4096, 14336):
            H0 = spd(n)
            W = torch.randn(256, n, device="cuda")
            l0 = ops.launch_count()
            mn, av = timed(lambda: ops.prepare(H0.clone(), W, 0.01), warm=1, it=2)
            print(f"prepare n={n} defaults (f16 GEMM, diag v4, group 4): {mn:.2f} ms (avg {av:.2f}), launches/call {(ops.launch_count() - l0) // 3}", flush=True)
            for be, dv, grp in (("f16", "4", "1"), ("f16", "4", "2"), ("f16", "4", "8"), ("f16", "3", "4"), ("f16", "1", "4"), ("tf32", "4", "4"),
                                ("tf32", "1", "1")):      # tcgen05 GEMM back-end, diagonal-block kernel, steps per trailing update
                os.environ.update(GQ_PREPARE_GEMM=be, GQ_DIAG_V2=dv, GQ_PREPARE_GROUP=grp)
                mn, av = timed(lambda: ops.prepare(H0.clone(), W, 0.01), warm=1, it=2)
                print(f"prepare n={n} GQ_PREPARE_GEMM={be} GQ_DIAG_V2={dv} GQ_PREPARE_GROUP={grp}: {mn:.2f} ms (avg {av:.2f})", flush=True)
            for k in ("GQ_PREPARE_GEMM", "GQ_DIAG_V2", "GQ_PREPARE_GROUP"):
                os.environ.pop(k, None)
    if "hessian" in what:
        for n, T in ((4096, 16384), (14336, 16384)):
            X = torch.randn(T, n, device="cuda").to(torch.bfloat16)
            H = torch.zeros(n, n, device="cuda")
            mn, av = timed(lambda: ops.hessian_update(H, X, 0.5, 0.5), warm=1, it=3)
            print(f"hessian n={n} T={T}: {mn:.3f} ms  -> {2 * T * n * n / 2 / mn / 1e9:.0f} TFLOP/s (upper-triangle flops)", flush=True)
    if "gptq" in what:
        shapes = ((6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336))
        if "small" in what:      # fewer CTAs streaming the same U: probes L2 hot-line contention
            shapes = ((256, 4096), (1024, 4096), (2048, 4096), (4096, 4096))
        for rows, n in shapes:
            U = torch.triu(torch.randn(n, n, device="cuda") * 0.01) + torch.eye(n, device="cuda")
            W0 = torch.randn(rows, n, device="cuda") * 0.02
            mn, av = timed(lambda: ops.gptq_quantize(W0.clone(), U, 12, wdeq_dtype=torch.bfloat16), warm=1, it=2)
            fl = rows * n * (n - 128)
            print(f"gptq {rows}x{n}: {mn:.2f} ms -> rank-k {fl / mn / 1e9:.1f} TFLOP/s", flush=True)
            phase_clocks(lambda: ops.gptq_quantize(W0.clone(), U, 12, wdeq_dtype=torch.bfloat16), (rows + 31) // 32)
            for grp in (() if "nofast" in what else ("1", "2", "4")):
                os.environ["GQ_FAST_GROUP"] = grp
                mn, av = timed(lambda: ops.gptq_quantize(W0.clone(), U, 12, wdeq_dtype=torch.bfloat16, mode=1), warm=1, it=2)
                ops.profile_enable(True)
                ops.gptq_quantize(W0.clone(), U, 12, wdeq_dtype=torch.bfloat16, mode=1)
                pr = ops.profile_read()
                ops.profile_enable(False)
                print(f"gptq FAST group={grp} {rows}x{n}: {mn:.2f} ms -> rank-k {fl / mn / 1e9:.1f} TFLOP/s; panel {pr['panel_ms']:.2f} ms/{pr['panel_launches']}, "
                      f"split {pr['split_ms']:.2f} ms/{pr['split_launches']}, "
                      f"tcgen05 rank-k GEMMs {pr['rankk_gemm_ms']:.2f} ms/{pr['rankk_gemm_launches']} -> {rows * n * (n - 256) / max(pr['rankk_gemm_ms'], 1e-9) / 1e9:.0f} TFLOP/s", flush=True)
            os.environ.pop("GQ_FAST_GROUP", None)
    if "schedules" in what:
        # the two bit-identical schedules of the exact arithmetic on row slices (what one rank of an N-GPU run launches)
        for rows, n in ((4096, 14336), (2048, 14336), (1024, 14336), (512, 14336), (4096, 4096), (512, 4096), (28672, 4096), (14336, 4096), (3584, 4096), (6144, 4096), (3072, 4096), (768, 4096)):
            U = torch.triu(torch.randn(n, n, device="cuda") * 0.01) + torch.eye(n, device="cuda")
            W0 = torch.randn(rows, n, device="cuda") * 0.02
            res = {}
            for name, mode in (("auto", 0), ("left", 2), ("right", 3)):
                mn, _ = timed(lambda: ops.gptq_quantize(W0.clone(), U, 12, wdeq_dtype=torch.bfloat16, mode=mode), warm=1, it=2)
                res[name] = mn
            print(f"schedules {rows}x{n}: " + ", ".join(f"{k} {v:.2f} ms" for k, v in res.items()), flush=True)
            del U, W0
    if "rtn" in what:
        W = torch.randn(128256, 4096, device="cuda").to(torch.bfloat16)
        mn, av = timed(lambda: ops.rtn_quantize(W, 12, wdeq_dtype=torch.bfloat16), warm=1, it=2)
        print(f"rtn 128256x4096: {mn:.2f} ms", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"wall {time.time() - t0:.1f} s")
