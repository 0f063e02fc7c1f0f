// TEST-ONLY: runs the DEVICE source gptq_gguf_toolkit_b200/csrc/kquant_bf16.cuh on the host (through host_shim/kquant.cuh) on
// the (rows x 256) super-block slabs of a weight: the four scale tensors of get_scale_and_zero in bf16 arithmetic.
//   g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -I tests/helpers/host_shim -I gptq_gguf_toolkit_b200/csrc -shared -fPIC ...
#define GQ_HOST_SHIM 1
#include "host_shim_intrinsics.h"   // stands in for common.cuh
#include "kquant.cuh"               // product header (kq_sum8 ...)
#include "kquant_bf16.cuh"          // the product header under test

template <int QT, int RND>
static void run(const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep, uint16_t *d, uint16_t *dmin,
                uint8_t *sq, uint8_t *zq) {
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS, MAXQ = (1 << Fmt<QT>::BITS) - 1;
    SearchParams sp;
    sp.nstep = nstep;
    for (int i = 0; i <= nstep && i < 64; ++i) sp.num[i] = (float)(rmin + rdelta * (double)i + (double)MAXQ);
    const int nsb = d_col / GQ_QK_K, ng = d_col / GS;
    for (int r = 0; r < d_row; ++r)
        for (int sb = 0; sb < nsb; ++sb) {
            float gs[16], gz[16];
            for (int g = 0; g < GPR; ++g) {
                float x[GS];
                for (int k = 0; k < GS; ++k) x[k] = W[(long)r * d_col + sb * GQ_QK_K + g * GS + k];
                if constexpr (Fmt<QT>::ASYM) kqb_search_asym<GS, MAXQ, RND>(x, sp, gs[g], gz[g]);
                else kqb_search_sym<GS, MAXQ, RND>(x, gs[g], gz[g]);
            }
            kqb_row_finalize<QT, RND>(gs, gz, d[(long)r * nsb + sb], dmin[(long)r * nsb + sb], sq + (long)r * ng + sb * GPR,
                                 zq + (long)r * ng + sb * GPR);
        }
}

template <int RND>
static int dispatch(int qtype, const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep, uint16_t *d, uint16_t *dmin,
                    uint8_t *sq, uint8_t *zq) {
    switch (qtype) {
    case GQ_Q2_K: run<GQ_Q2_K, RND>(W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq); return 0;
    case GQ_Q3_K: run<GQ_Q3_K, RND>(W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq); return 0;
    case GQ_Q4_K: run<GQ_Q4_K, RND>(W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq); return 0;
    case GQ_Q5_K: run<GQ_Q5_K, RND>(W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq); return 0;
    case GQ_Q6_K: run<GQ_Q6_K, RND>(W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq); return 0;
    }
    return -1;
}

// rnd: 2 = bf16, 1 = fp16 (the gq_dtype codes)
extern "C" int host_scales_native(int rnd, int qtype, const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep,
                                  uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq) {
    return rnd == 2 ? dispatch<GQ_RND_BF16>(qtype, W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq)
                    : dispatch<GQ_RND_F16>(qtype, W, d_row, d_col, rmin, rdelta, nstep, d, dmin, sq, zq);
}
