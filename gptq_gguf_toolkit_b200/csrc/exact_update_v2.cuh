// exact_update_v2.cuh -- EXPERIMENTAL variant of exact_update_kernel (csrc/gptq_layer.cu), kept in its own header so that the
// CPU suite can run it on the SIMT emulator (tests/helpers/simt_emu, tests/test_simt_emu_cpu.py) before it ever runs on a GPU.
#pragma once
#ifndef SIMT_EMU
#include "common.cuh"
#endif

namespace upd2 {
constexpr int R = 32;        // rows per CTA
constexpr int KP = 16;       // k's per pipeline piece
constexpr int S = 4;         // pipeline stages
constexpr int US_FLOATS = KP * 256;
constexpr int ES_FLOATS = R * KP;
constexpr int NT2 = 128;
constexpr size_t SMEM_BYTES = (size_t)S * (US_FLOATS + ES_FLOATS) * sizeof(float);
struct Params { float *W; const float *U; int d_row, d_col; };
}  // namespace upd2

// EXPERIMENTAL (GQ_UPDATE_V2=1; off by default, not yet run on hardware -- written at the end of round 1 after the GPU
// budget was spent; tests/test_gpu_schedules.py::test_update_v2_bit_identical is skipped unless GQ_TEST_EXPERIMENTAL=1).
// Same arithmetic as exact_update_kernel, different work split.  ncu on exact_update_kernel: FMA pipe 60 % busy, and the
// shared-memory pipe is the co-limiter -- per k a warp issues 32 FFMAs and 6 shared-memory wavefronts (one 512-byte row
// segment of U = 4 wavefronts, 8 broadcast E values = 2), 16 warps per SM => 96 wavefronts against 128 FFMA issue cycles.
// Here a warp owns 16 rows x 128 columns (16 x 4 accumulators per thread): 64 FFMAs per 8 wavefronts per k, i.e. 0.125
// instead of 0.1875 wavefronts per FFMA; four warps per CTA, three CTAs per SM (12 warps).  `w` is not held in registers:
// at each of the two 128-k boundaries the tile is read from global memory (L2), updated and written back -- per element still
// w <- (w - chain1) - chain2 with two roundings, bit-identical to the other schedules.
__global__ void __launch_bounds__(upd2::NT2, 3) exact_update_v2_kernel(const upd2::Params p, const int c) {
    using namespace upd2;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *Us = reinterpret_cast<float *>(smem_raw);
    float *Es = Us + S * US_FLOATS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rh = warp >> 1, ch = warp & 1;             // rows 16*rh .. +15, columns ch*128 + 4*lane .. +3
    const int r0 = blockIdx.y * R;
    const int cw = c + 256 * (1 + blockIdx.x);
    const size_t ld = (size_t)p.d_col;
    const float *__restrict__ U = p.U;
    float *__restrict__ Wg = p.W;
    constexpr int P = 256 / KP;                      // 16 pieces of 16 k's
    auto issue = [&](int pc) {
        if (pc < P) {
            const int k0 = c + KP * pc, st = pc % S;
            float *us = Us + st * US_FLOATS;
            float *es = Es + st * ES_FLOATS;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int id = tid + NT2 * m, row = id >> 6, c16 = id & 63;
                cp_async16(us + row * 256 + 4 * c16, U + (size_t)(k0 + row) * ld + cw + 4 * c16);
            }
            {
                const int row = tid >> 2, part = tid & 3;
                const int gr = min(r0 + row, p.d_row - 1);
                cp_async16(es + row * KP + 4 * part, Wg + (size_t)gr * ld + k0 + 4 * part);
            }
        }
        cp_async_commit();
    };
    float acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int s = 0; s < S - 1; ++s) issue(s);
    for (int pc = 0; pc < P; ++pc) {
        cp_async_wait<S - 2>();
        __syncthreads();
        issue(pc + S - 1);
        const float *us = Us + (pc % S) * US_FLOATS + ch * 128 + 4 * lane;
        const float *es = Es + (pc % S) * ES_FLOATS + (16 * rh) * KP;
#pragma unroll
        for (int kk = 0; kk < KP; kk += 2) {
            float2 e[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = *reinterpret_cast<const float2 *>(es + i * KP + kk);
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
                const float4 u = *reinterpret_cast<const float4 *>(us + (kk + k2) * 256);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float ev = k2 == 0 ? e[i].x : e[i].y;
                    acc[i][0] = __fmaf_rn(ev, u.x, acc[i][0]);
                    acc[i][1] = __fmaf_rn(ev, u.y, acc[i][1]);
                    acc[i][2] = __fmaf_rn(ev, u.z, acc[i][2]);
                    acc[i][3] = __fmaf_rn(ev, u.w, acc[i][3]);
                }
            }
        }
        if ((pc & 7) == 7) {   // end of a 128-column block: w <- w - acc (gptq.py:270, alpha = -1), through global memory
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int gr = r0 + 16 * rh + i;
                if (gr < p.d_row) {
                    float4 *wp = reinterpret_cast<float4 *>(Wg + (size_t)gr * ld + cw + ch * 128 + 4 * lane);
                    float4 v = *wp;
                    v.x = __fsub_rn(v.x, acc[i][0]); v.y = __fsub_rn(v.y, acc[i][1]);
                    v.z = __fsub_rn(v.z, acc[i][2]); v.w = __fsub_rn(v.w, acc[i][3]);
                    *wp = v;
                }
                acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
            }
        }
    }
    cp_async_wait<0>();
}

