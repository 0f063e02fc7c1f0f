"""TEST-ONLY: runs bench.py's own arm on a machine without a GPU by replacing everything that touches CUDA (events, the
model, the Quantizer, libgq's counters) with stand-ins, so that the control flow and the JSON contract of the bench line
can be checked in the CPU suite (tests/test_bench_contract_cpu.py).  `--fail-fast` makes the fast-mode extra raise."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
real_device = torch.device
import bench

class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, *a): pass
    def elapsed_time(self, other): return 1234.0
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Event = FakeEvent
from gptq_gguf_toolkit_b200 import ops, quantizer as Q
import gptq_gguf_toolkit_b200.data_utils, gptq_gguf_toolkit_b200.quant, torch.distributed
class TinyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(8, 8, bias=False)
        self.model = torch.nn.Module(); self.model.embed_tokens = torch.nn.Embedding(16, 8)
bench.build_model = lambda w, device: TinyModel()
bench.torch.device = lambda s: real_device("cpu")
ops.launch_count = lambda: 7
ops.profile_enable = lambda on: None
ops.profile_read = lambda: {"panel_ms": 100.0, "panel_launches": 10, "rankk_gemm_ms": 200.0, "rankk_gemm_launches": 10}

class FakeTimer:
    def __init__(self, enabled=True): pass
    def totals(self): return {"gptq": 2.0, "hessian": 1.0, "prepare_host": 0.1, "rtn": 0.01, "forward1": 3.0}
class FakeQ:
    def __init__(self, model, **kw):
        self.results = {"m": {"qweight": torch.zeros(4, dtype=torch.uint8), "q_type": 12}}
        self.kw = kw
    def quantize(self, cfg):
        if self.kw["quantizer_kwargs"]["mode"] == "fast" and FAIL_FAST: raise RuntimeError("boom")
    def non_invertible_modules(self): return []
Q.PhaseTimer = FakeTimer
Q.Quantizer = FakeQ
FAIL_FAST = "--fail-fast" in sys.argv
if FAIL_FAST: sys.argv.remove("--fail-fast")
bench.cpu_reference_sample = lambda w, threads: (99.0, {"x": 1})
bench.ClockSampler.start = lambda self: None
bench.ClockSampler.stop = lambda self: {"sm_mhz": 1.0, "sm_max_mhz": 2.0, "reasons": []}
bench.main()
