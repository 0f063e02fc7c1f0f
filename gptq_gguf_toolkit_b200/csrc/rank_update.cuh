// rank_update.cuh -- the exact rank-k arithmetic of the column loop (csrc/gptq_layer.cu): rank_update<> (shared by the fused
// layer kernel and the trailing-update kernel) and the body of exact_update_kernel, in a header of their own so that the CPU
// suite can run the shipped source on the SIMT emulator (tests/test_simt_emu_cpu.py) next to the `-m gpu` parity tests.
#pragma once
#ifndef SIMT_EMU
#include "common.cuh"
#endif

namespace rk {
constexpr int R = 32;        // rows per CTA
constexpr int NT = 256;      // threads per CTA
#ifndef GQ_RK_KP
#define GQ_RK_KP 16
#define GQ_RK_S 4
#endif
constexpr int KP = GQ_RK_KP; // k's per pipeline piece (one block-wide barrier per piece)
constexpr int S = GQ_RK_S;   // pipeline stages
constexpr int US_FLOATS = KP * 256;
constexpr int ES_FLOATS = R * KP;
}  // namespace rk

// Variants of this loop that were built and measured on B200 and did NOT beat it (profiles/r01_gptq_kernel_notes.md,
// profiles/microbench/): packed FFMA2 (same FMA rate, 5-6 register reads per issue), a column-per-warp mapping with a
// k-major E (fewer shared-memory wavefronts, but 4-byte cp.async transposes cost more than they save), a
// producer/consumer warp split with 8 x 8 register tiles (one warp per SM sub-partition issues an FFMA only every
// ~1.65 cycles), and a TMA + mbarrier ring.  In isolation the FFMA pattern itself reaches 70 cycles per k (0.91 FFMA
// per cycle per sub-partition); the pipelined loop runs at ~111.
// tile(8 rows x 4 cols per thread) -= E[:, kbeg:kend] * U[kbeg:kend, window]; the reference's addmm_ arithmetic.
// HALF: only the window's columns 128..255 are updated (warps with ch == 1 compute, all warps load).
template <bool HALF, class Params>
__device__ __forceinline__ void rank_update(float (&w)[8][4], const Params &p, float *__restrict__ Us_base,
                                            float *__restrict__ Es_base, int r0, int c,
                                            int kbeg, int kend, int tid, int rg, int ch, int lane) {
    using namespace rk;
    const int P = (kend - kbeg) / KP;
    const float *__restrict__ U = p.U;
    const float *__restrict__ Wg = p.W;
    const size_t ld = (size_t)p.d_col;
    auto issue = [&](int pc) {
        if (pc < P) {
            const int k0 = kbeg + KP * pc, st = pc % S;
            float *us = Us_base + st * US_FLOATS;
            float *es = Es_base + st * ES_FLOATS;
#pragma unroll
            for (int m = 0; m < KP / 4; ++m) {
                const int id = tid + NT * m, row = id >> 6, c16 = id & 63;
                if (!HALF || c16 >= 32) cp_async16(us + row * 256 + 4 * c16, U + (size_t)(k0 + row) * ld + c + 4 * c16);
            }
            if (tid < R * (KP / 4)) {
                const int row = tid / (KP / 4), part = tid % (KP / 4);
                const int gr = min(r0 + row, p.d_row - 1);
                cp_async16(es + row * KP + 4 * part, Wg + (size_t)gr * ld + k0 + 4 * part);
            }
        }
        cp_async_commit();
    };
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int s = 0; s < S - 1; ++s) issue(s);
    for (int pc = 0; pc < P; ++pc) {
        cp_async_wait<S - 2>();
        __syncthreads();
        issue(pc + S - 1);
        if (!HALF || ch == 1) {
            const float *us = Us_base + (pc % S) * US_FLOATS + ch * 128 + 4 * lane;
            const float *es = Es_base + (pc % S) * ES_FLOATS + (8 * rg) * KP;
#pragma unroll
            for (int kk = 0; kk < KP; kk += 4) {
                float4 e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = *reinterpret_cast<const float4 *>(es + i * KP + kk);
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const float4 u = *reinterpret_cast<const float4 *>(us + (kk + k2) * 256);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float ev = k2 == 0 ? e[i].x : k2 == 1 ? e[i].y : k2 == 2 ? e[i].z : e[i].w;
                        acc[i][0] = __fmaf_rn(ev, u.x, acc[i][0]);
                        acc[i][1] = __fmaf_rn(ev, u.y, acc[i][1]);
                        acc[i][2] = __fmaf_rn(ev, u.z, acc[i][2]);
                        acc[i][3] = __fmaf_rn(ev, u.w, acc[i][3]);
                    }
                }
            }
        }
        if ((pc + 1) % (128 / KP) == 0) {   // end of one earlier 128-column block: w <- w - acc  (gptq.py:270, alpha = -1)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    w[i][j] = __fsub_rn(w[i][j], acc[i][j]);
                    acc[i][j] = 0.0f;
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}


// Body of exact_update_kernel (csrc/gptq_layer.cu): Params needs W, U, d_row, d_col.
template <class Params>
__device__ __forceinline__ void exact_update_body(const Params &p, const int c, uint8_t *smem_raw) {
    using namespace rk;
    float *Us = reinterpret_cast<float *>(smem_raw);
    float *Es = Us + S * US_FLOATS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rg = warp >> 1, ch = warp & 1;
    const int r0 = blockIdx.y * R;
    const int cw = c + 256 * (1 + blockIdx.x);       // this CTA's window of 256 later columns
    const size_t ld = (size_t)p.d_col;
    float w[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gr = min(r0 + 8 * rg + i, p.d_row - 1);
        const float4 v = *reinterpret_cast<const float4 *>(p.W + (size_t)gr * ld + cw + ch * 128 + 4 * lane);
        w[i][0] = v.x; w[i][1] = v.y; w[i][2] = v.z; w[i][3] = v.w;
    }
    rank_update<false>(w, p, Us, Es, r0, cw, c, c + 256, tid, rg, ch, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gr = r0 + 8 * rg + i;
        if (gr < p.d_row)
            *reinterpret_cast<float4 *>(p.W + (size_t)gr * ld + cw + ch * 128 + 4 * lane) =
                make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
    }
}

