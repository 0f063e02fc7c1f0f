// exact_update64.cuh -- the trailing update of the exact right-looking schedule with an 8 x 8 register tile per thread.
//
// Same arithmetic as exact_update_body (rank_update.cuh) -- per element of W[:, c+256:], for each of the finished super-block's two
// 128-column blocks a fresh single-accumulator fp32 FMA chain over its 128 k's in ascending order, then ONE subtraction
// (gptq.py:270 == addmm_ on CPU) -- on a different shape of work per thread: a CTA of 256 threads owns 64 rows x 256 columns, a
// warp 8 rows, a lane the columns 4l..4l+3 and 128+4l..128+4l+3, i.e. 64 accumulators and the 64 weights they are subtracted from
// stay in registers (one CTA per SM).  Per k a warp issues 64 FFMAs for 2 LDS.128 of U (both conflict-free, 512 contiguous bytes
// per warp) and 2 LDS.128 of E (8 broadcast loads per 4 k's): 0.06 shared-memory loads per FFMA against 0.09 with the 8 x 4 tile,
// half the shared-memory wavefronts per FFMA, half the cp.async issues and half the L2 -> shared-memory traffic for U per row
// (the slab is shared by 64 rows), and one block-wide barrier per 32 k's.
#pragma once
#ifndef SIMT_EMU
#include "common.cuh"
#endif

namespace rk64 {
constexpr int R = 64;        // rows per CTA
constexpr int NT = 256;      // threads per CTA
constexpr int KP = 32;       // k's per pipeline piece (one block-wide barrier per piece)
constexpr int S = 3;         // pipeline stages
constexpr int US_FLOATS = KP * 256;
constexpr int ES_FLOATS = R * KP;
constexpr size_t SMEM_BYTES = (size_t)S * (US_FLOATS + ES_FLOATS) * sizeof(float);      // 120 KB
}  // namespace rk64

// Params needs W, U, d_row, d_col.  grid = (windows of 256 later columns, groups of 64 rows).
template <class Params>
__device__ __forceinline__ void exact_update64_body(const Params &p, const int c, uint8_t *smem_raw) {
    using namespace rk64;
    float *Us_base = reinterpret_cast<float *>(smem_raw);
    float *Es_base = Us_base + S * US_FLOATS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.y * R;
    const int cw = c + 256 * (1 + blockIdx.x);       // this CTA's window of 256 later columns
    const size_t ld = (size_t)p.d_col;
    const float *__restrict__ U = p.U;
    float *__restrict__ Wg = p.W;

    constexpr int P = 256 / KP;                      // pieces of the finished super-block's 256 k's
    auto issue = [&](int pc) {
        if (pc < P) {
            const int k0 = c + KP * pc, st = pc % S;
            float *us = Us_base + st * US_FLOATS;
            float *es = Es_base + st * ES_FLOATS;
#pragma unroll
            for (int m = 0; m < KP / 4; ++m) {       // U[k0 + row, cw : cw + 256]: 64 float4 per row
                const int id = tid + NT * m, row = id >> 6, c16 = id & 63;
                cp_async16(us + row * 256 + 4 * c16, U + (size_t)(k0 + row) * ld + cw + 4 * c16);
            }
#pragma unroll
            for (int m = 0; m < (R * KP / 4) / NT; ++m) {      // E[row, k0 : k0 + KP] = W[r0 + row, k0 ...]: KP/4 float4 per row
                const int id = tid + NT * m, row = id / (KP / 4), part = id % (KP / 4);
                const int gr = min(r0 + row, p.d_row - 1);
                cp_async16(es + row * KP + 4 * part, Wg + (size_t)gr * ld + k0 + 4 * part);
            }
        }
        cp_async_commit();
    };

    for (int s = 0; s < S - 1; ++s) issue(s);

    // the 8 x 8 tile of W: rows 8*warp + i, columns cw + 4*lane + (0..3) and cw + 128 + 4*lane + (0..3)
    float w[8][8], acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gr = min(r0 + 8 * warp + i, p.d_row - 1);
        const float4 a = *reinterpret_cast<const float4 *>(Wg + (size_t)gr * ld + cw + 4 * lane);
        const float4 b = *reinterpret_cast<const float4 *>(Wg + (size_t)gr * ld + cw + 128 + 4 * lane);
        w[i][0] = a.x; w[i][1] = a.y; w[i][2] = a.z; w[i][3] = a.w;
        w[i][4] = b.x; w[i][5] = b.y; w[i][6] = b.z; w[i][7] = b.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    }

    for (int pc = 0; pc < P; ++pc) {
        cp_async_wait<S - 2>();
        __syncthreads();
        issue(pc + S - 1);
        const float *us = Us_base + (pc % S) * US_FLOATS + 4 * lane;
        const float *es = Es_base + (pc % S) * ES_FLOATS + (8 * warp) * KP;
#pragma unroll 2
        for (int kk = 0; kk < KP; kk += 4) {
            float4 e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = *reinterpret_cast<const float4 *>(es + i * KP + kk);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
                const float4 u0 = *reinterpret_cast<const float4 *>(us + (kk + k2) * 256);
                const float4 u1 = *reinterpret_cast<const float4 *>(us + (kk + k2) * 256 + 128);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float ev = k2 == 0 ? e[i].x : k2 == 1 ? e[i].y : k2 == 2 ? e[i].z : e[i].w;
                    acc[i][0] = __fmaf_rn(ev, u0.x, acc[i][0]);
                    acc[i][1] = __fmaf_rn(ev, u0.y, acc[i][1]);
                    acc[i][2] = __fmaf_rn(ev, u0.z, acc[i][2]);
                    acc[i][3] = __fmaf_rn(ev, u0.w, acc[i][3]);
                    acc[i][4] = __fmaf_rn(ev, u1.x, acc[i][4]);
                    acc[i][5] = __fmaf_rn(ev, u1.y, acc[i][5]);
                    acc[i][6] = __fmaf_rn(ev, u1.z, acc[i][6]);
                    acc[i][7] = __fmaf_rn(ev, u1.w, acc[i][7]);
                }
            }
        }
        if ((pc + 1) % (128 / KP) == 0) {   // end of one 128-column block: w <- w - acc  (gptq.py:270, alpha = -1)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    w[i][j] = __fsub_rn(w[i][j], acc[i][j]);
                    acc[i][j] = 0.0f;
                }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gr = r0 + 8 * warp + i;
        if (gr < p.d_row) {
            *reinterpret_cast<float4 *>(Wg + (size_t)gr * ld + cw + 4 * lane) = make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
            *reinterpret_cast<float4 *>(Wg + (size_t)gr * ld + cw + 128 + 4 * lane) = make_float4(w[i][4], w[i][5], w[i][6], w[i][7]);
        }
    }
}
