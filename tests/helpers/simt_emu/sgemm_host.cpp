// TEST-ONLY: gptq_gguf_toolkit_b200/csrc/sgemm.cuh (sg::sgemm_kernel: the fp32 SIMT tile GEMM behind gq_hessian_update for fp32
// activations and the GQ_PREPARE_SIMT chain) on the SIMT emulator.
#define SIMT_EMU 1
#define GQ_HOST_SHIM 1
#include "simt_emu.h"
#undef __shared__
#define __shared__ static          // this kernel only has statically sized shared arrays; blocks run one after the other
#include "host_shim_intrinsics.h"
#include "sgemm.cuh"

// H <- beta * H + alpha * X^T X  exactly as gq_hessian_update sets it up (csrc/linalg.cu): upper tiles + mirrored store
extern "C" int run_hessian_simt(float *H, const float *X, long n_tok, int d_col, float beta, float alpha) {
    sg::Args a;
    a.A = X; a.B = X; a.C = H;
    a.a_rs = 1; a.a_cs = d_col;
    a.b_rs = d_col; a.b_cs = 1;
    a.ldc = d_col; a.a_batch = a.b_batch = a.c_batch = 0;
    a.M = d_col; a.N = d_col; a.K = (int)n_tok;
    a.alpha = alpha; a.beta = beta; a.tile_mode = sg::TM_UPPER_MIRROR; a.k_mode = sg::KM_FULL; a.in_dtype = GQ_F32;
    simt::launch(dim3(a.N / sg::BN, a.M / sg::BM, 1), dim3(sg::NT), [&]() { sg::sgemm_kernel(a); });
    return 0;
}
