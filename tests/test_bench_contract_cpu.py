"""bench.py's JSON line (the driver's contract) without a GPU: tests/helpers/bench_harness.py swaps every CUDA-touching piece
for a stand-in and runs bench.main() for real, so a missing key, a NameError in a rarely taken branch or a lost line when an
optional extra fails shows up here and not at the end of a round.  The reference arm runs as it is (bounded CPU sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "helpers", "bench_harness.py")


def _run(*args):
    out = subprocess.run([sys.executable, HARNESS, "--workload", "tiny", *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "rank 0 prints exactly ONE JSON line"
    return json.loads(lines[0])


def test_own_arm_line_has_every_contract_key():
    d = _run("--compare-left", "1")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "llama3_8b_q4k_quantize_wall_clock_s" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["gpu_launches"] == 0 or isinstance(d["gpu_launches"], int)
    assert "left_looking_schedule" in d and "fast_mode" in d


def test_failing_extra_does_not_lose_the_line():
    d = _run("--fail-fast")
    assert "fast_mode" not in d and any("fast mode failed" in n for n in d["notes"])
    assert d["value"] > 0 and "roofline" in d and "cpu_baseline" in d


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "tiny"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["metric"] == "llama3_8b_q4k_quantize_wall_clock_s" and d["unit"] == "s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the line's own step time is the bounded sample's wall time, not the extrapolated metric
    assert d["cpu_baseline"]["extrapolated"] is True and 0 < d["ms_per_step"] < 600e3
    assert set(("hessian", "prepare", "step", "forwards", "rtn")) <= set(d["cpu_baseline"]["sample_seconds"])
    # both arms name the SAME workload: identical `config` objects
    own = _run()
    assert own["config"] == d["config"]
