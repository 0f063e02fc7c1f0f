"""Out-of-bounds check of the native-arithmetic RTN kernel body (csrc/rtn_native.cuh) on the SIMT emulator: every global buffer is
placed so that its end touches an inaccessible page (an overrun faults) and the slack before it carries a canary (an underrun is
seen), all five types, ragged row counts.  Run by hand:  python tests/helpers/guard_page_check.py
Recorded on 2026-10-17: clean."""
import sys, ctypes as C, numpy as np, mmap, subprocess, torch
sys.path.insert(0,'/root/repo')
ROOT='/root/repo'; EMU=ROOT+'/tests/helpers/simt_emu'
subprocess.run(["g++","-O1","-std=c++17","-ffp-contract=off","-fno-fast-math","-Wno-unknown-pragmas","-fPIC","-shared","-I",EMU,"-I",ROOT+"/tests/helpers/host_shim","-I",ROOT+"/gptq_gguf_toolkit_b200/csrc",EMU+"/rtn_native_host.cpp","-o","/tmp/rn.so"],check=True,capture_output=True)
lib=C.CDLL('/tmp/rn.so'); libc=C.CDLL(None)
PAGE=mmap.PAGESIZE
def guarded(nbytes):
    """buffer of nbytes whose END touches an inaccessible page and whose START is preceded by one (OOB reads/writes fault)"""
    n=(nbytes+PAGE-1)//PAGE*PAGE
    libc.mmap.restype=C.c_void_p; libc.mmap.argtypes=[C.c_void_p,C.c_size_t,C.c_int,C.c_int,C.c_int,C.c_long]
    base=libc.mmap(None,n+2*PAGE,3,0x22,-1,0)
    libc.mprotect.argtypes=[C.c_void_p,C.c_size_t,C.c_int]
    assert libc.mprotect(base,PAGE,0)==0 and libc.mprotect(base+PAGE+n,PAGE,0)==0
    start=base+PAGE+n-nbytes          # end-aligned: overruns fault immediately; underruns land in slack (checked by canary)
    return start, base+PAGE, n-nbytes
TS={10:84,11:110,12:144,13:176,14:210}; GS={10:16,11:16,12:32,13:32,14:16}
rng=np.random.default_rng(3)
for qt in (10,11,12,13,14):
  for d_row,d_col in ((33,512),(64,256),(1,256)):
    gs=GS[qt]; nsb=d_col//256
    sizes={"W":d_row*d_col*2,"qw":d_row*d_col,"d":d_row*nsb*2,"dm":d_row*nsb*2,"sq":d_row*(d_col//gs),"zq":d_row*(d_col//gs),"pk":d_row*nsb*TS[qt],"wd":d_row*d_col*2}
    bufs={}
    for k,nb in sizes.items():
        start,slack0,slack=guarded(nb); C.memset(slack0,0xAB,slack); bufs[k]=(start,slack0,slack)
    Wt=(torch.randn(d_row,d_col)*0.03).to(torch.bfloat16)
    C.memmove(bufs["W"][0], Wt.view(torch.int16).numpy().ctypes.data, sizes["W"])
    P=lambda k,t: C.cast(bufs[k][0], C.POINTER(t))
    rc=lib.run_rtn_native(C.c_int(2),C.c_int(qt),P("W",C.c_uint16),C.c_int(d_row),C.c_int(d_col),C.c_double(-1.0),C.c_double(0.1),C.c_int(20),P("qw",C.c_uint8),P("d",C.c_uint16),P("dm",C.c_uint16),P("sq",C.c_uint8),P("zq",C.c_uint8),P("pk",C.c_uint8),P("wd",C.c_uint16))
    for k,(start,slack0,slack) in bufs.items():
        if slack: assert bytes((C.c_ubyte*slack).from_address(slack0))==b"\xAB"*slack, (k,"underrun")
print("rtn_native: no out-of-bounds access (guard pages + canaries), all types, ragged rows")
