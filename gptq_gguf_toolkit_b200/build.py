"""Builds gptq_gguf_toolkit_b200/libgq.so (sm_100a only) with nvcc, in-tree.

    python -m gptq_gguf_toolkit_b200.build [--force] [--verbose]

-fmad=false is part of the numerical contract: the library reproduces the reference's CPU arithmetic,
so a*b+c must never be contracted; every FMA in the sources is an explicit __fmaf_rn().
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libgq.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "gq.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, extra_flags=(), so: str = SO, bdir_name: str = "build") -> str:
    """extra_flags / so / bdir_name: development variants (e.g. -DGQ_SERIAL_VARIANT=1 into libgq_v1.so), selected at
    run time with GQ_LIB_PATH (see _lib.py); the product build uses the defaults."""
    if so == SO and not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    bdir = os.path.join(HERE, bdir_name)
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(out, flush=True)
        if p.returncode != 0:
            failed = True
            print(f"nvcc failed on {src}", file=sys.stderr)
    if failed:
        raise RuntimeError("libgq build failed")
    cmd = [nvcc, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    if defs:     # python -m gptq_gguf_toolkit_b200.build -DGQ_SERIAL_VARIANT=1 --tag v1  ->  libgq_v1.so
        tag = sys.argv[sys.argv.index("--tag") + 1]
        print(build(force=True, verbose="--verbose" in sys.argv, extra_flags=defs,
                    so=os.path.join(HERE, f"libgq_{tag}.so"), bdir_name=f"build_{tag}"))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
