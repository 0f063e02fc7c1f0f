"""Randomised comparison of the SHIPPED column-loop kernel (run on the SIMT emulator, tests/helpers/simt_emu) with the CPU oracle:
random types, d_col in {256..1024}, 1..69 rows, both schedules, ill-conditioned correlated Hessians, zero rows / columns and
tie-heavy weights.  Not part of the default suite (minutes); run it by hand:

    python tests/helpers/fuzz_emulated_kernel.py [seconds]

Recorded on 2026-10-17: 222 cases in 420 s, 0 mismatches (codes, four scale tensors, packed bytes, dequantised weights); again after
the panel-kernel rework of round 2 (conversion-free search, shared-memory in-super-block update, serial steps on 4-column groups):
273 cases in 600 s, 0 mismatches."""
import sys, ctypes as C, numpy as np, subprocess, time
sys.path.insert(0,'/root/repo')
from oracle import oracle as orc
ROOT='/root/repo'; EMU=ROOT+'/tests/helpers/simt_emu'
subprocess.run(["g++","-O1","-std=c++17","-ffp-contract=off","-fno-fast-math","-Wno-unknown-pragmas","-fPIC","-shared","-I",EMU,"-I",ROOT+"/tests/helpers/host_shim","-I",ROOT+"/gptq_gguf_toolkit_b200/csrc",EMU+"/gptq_layer_host.cpp","-o","/tmp/gl.so"],check=True,capture_output=True)
lib=C.CDLL('/tmp/gl.so')
TS={10:84,11:110,12:144,13:176,14:210}; GS={10:16,11:16,12:32,13:32,14:16}
rng=np.random.default_rng(2026)
t0=time.time(); n=0; bad=0
BUDGET = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
while time.time()-t0 < BUDGET:
    qt=int(rng.choice([10,11,12,13,14])); d_col=int(rng.choice([256,512,768,1024])); d_row=int(rng.integers(1,70)); right=int(rng.integers(0,2))
    kind=int(rng.integers(0,4))
    W=(rng.standard_normal((d_row,d_col))*0.05*np.exp(rng.standard_normal((d_row,1)))).astype(np.float32)
    if kind==1: W[rng.integers(0,d_row)]=0.0
    if kind==2: W[:, rng.integers(0,d_col,8)]=0.0
    if kind==3: W=np.round(W*64)/64   # many ties
    W=W.astype(np.float32)
    X=(rng.standard_normal((2*d_col,d_col))@(rng.standard_normal((d_col,d_col))/np.sqrt(d_col))*np.exp(rng.standard_normal(d_col))).astype(np.float32)
    H=np.zeros((d_col,d_col),np.float32); orc.hessian_update(H,X,0.0,2.0/4)
    U,_,_,b=orc.prepare(H,W.copy(),0.01)
    if b: continue
    U=np.ascontiguousarray(U,dtype=np.float32)
    ref=orc.gptq_step(W.copy(),U,qt)
    Wk=W.copy(); nsb=d_col//256; gs=GS[qt]
    qw=np.zeros((d_row,d_col),np.uint8); d=np.zeros((d_row,nsb),np.uint16); dm=np.zeros_like(d); sq=np.zeros((d_row,d_col//gs),np.uint8); zq=np.zeros_like(sq); pk=np.zeros((d_row,nsb*TS[qt]),np.uint8); wd=np.zeros((d_row,d_col),np.float32)
    p=lambda a,t:a.ctypes.data_as(C.POINTER(t))
    lib.run_gptq_layer(C.c_int(qt),C.c_int(right),p(Wk,C.c_float),p(U,C.c_float),C.c_int(d_row),C.c_int(d_col),C.c_double(-1.0),C.c_double(0.1),C.c_int(20),p(qw,C.c_uint8),p(d,C.c_uint16),p(sq,C.c_uint8),p(dm,C.c_uint16),p(zq,C.c_uint8),p(pk,C.c_uint8),p(wd,C.c_float))
    ok=all(np.array_equal(a.view(np.uint8), (r.view(np.uint16) if r.dtype==np.float16 else r).view(np.uint8).reshape(a.shape[0],-1)) for a,r in zip((qw,d,sq,dm,zq),ref[:5]))
    ok=ok and np.array_equal(pk, orc.pack(qt,*ref[:5])) and np.array_equal(wd, ref[5]) and np.array_equal(Wk, ref[6] if len(ref)>6 else Wk)
    n+=1
    if not ok:
        bad+=1; print("MISMATCH", qt, d_row, d_col, right, kind)
print("cases", n, "mismatches", bad, "len(ref)", len(ref))
