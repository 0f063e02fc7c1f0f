// hessian_tc.cu -- tensor-core Hessian accumulation  H <- beta*H + alpha * X^T X  for 16-bit activations
// (replaces GPTQ.update's fp32 addmm_, quant/gptq/src/gptq.py:108-112 -- the FLOP giant of the whole run).
//
// bf16 x bf16 (and fp16 x fp16) products are exact in fp32, so feeding the 16-bit activations straight to
// tcgen05.mma kind::f16 with fp32 TMEM accumulators gives the same class of result as the reference's fp32
// GEMM of the widened inputs (only the summation order differs, which is unspecified in the reference too).
//
// Pipeline (one persistent CTA per SM, 256 threads, warp-specialised):
//   operands         : X (T x n, channels contiguous) is fed to the tensor cores AS IT IS, as MN-major operands (round 2): TMA boxes
//                      of 64 tokens x 64 channels (128-byte rows, 128B swizzle) land as 1024-byte swizzle atoms of 64 channels x
//                      8 tokens, SBO = 1024 B between token groups, LBO = 8192 B between 64-channel columns; tokens past T are
//                      zero-filled by the TMA unit.  (GQ_HESSIAN_MN=0: round 1's path -- a transpose kernel writes Xt (n x Tp)
//                      first so that both operands are K-major; 7-22 % slower and one more pass over the activations.)
//   warp 0           : TMA producer, cp.async.bulk.tensor 2D, 128B swizzle, 4-stage mbarrier ring
//                      stage = A tile 128 x 64 (16 KB) + B tile 256 x 64 (32 KB) of Xt
//   warp 1           : one elected thread issues tcgen05.mma.cta_group::1.kind::f16, M128 N256 K16, accumulators
//                      double-buffered in TMEM (2 x 256 columns) so the epilogue overlaps the next tile's MMAs
//   warps 4-7        : epilogue: tcgen05.ld -> H = beta*H + alpha*acc, direct + mirrored store (H stays symmetric)
// Only tiles that touch the upper triangle are computed (SYRK); strictly-lower elements come from the mirror.
#include <cstdlib>

#include "tc_common.cuh"

#ifndef GQ_HESSIAN_MN_DEFAULT
#define GQ_HESSIAN_MN_DEFAULT true
#endif

namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NTHREADS = 256;
constexpr uint32_t TMEM_COLS = 512;

struct Barriers {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(Barriers);

// Upper-triangle tile list: tile (tm, tn) is needed iff its last column >= its first row, i.e. tn >= tm/2.
__device__ __forceinline__ void tile_coords(int idx, int ntn, int &tm, int &tn) {
    int m = 0;
    while (true) {
        const int cnt = ntn - (m >> 1);
        if (idx < cnt) break;
        idx -= cnt;
        ++m;
    }
    tm = m;
    tn = (m >> 1) + idx;
}

// ---------------------------------------------------------------------------------------------
// X (T x n) -> Xt (n x Tp), zero padded in t
// ---------------------------------------------------------------------------------------------
// (64 tokens x 64 channels) tiles; 128-bit global loads and stores on both sides, the transposition itself moves
// 32-bit words holding two consecutive tokens of one channel through a conflict-free (stride 33) shared-memory tile.
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t *__restrict__ X, uint16_t *__restrict__ Xt,
                                                          int T, int n, int Tp) {
    __shared__ uint32_t tile[64 * 33];
    const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    {
        const int tp = threadIdx.x >> 3, v = threadIdx.x & 7;          // token pair, group of 8 channels
        const int ta = t0 + 2 * tp, tb = ta + 1;
        union { uint4 u; uint16_t h[8]; } r0, r1;
        r0.u = make_uint4(0, 0, 0, 0);
        r1.u = make_uint4(0, 0, 0, 0);
        if (ta < T) r0.u = *reinterpret_cast<const uint4 *>(X + (size_t)ta * n + c0 + 8 * v);
        if (tb < T) r1.u = *reinterpret_cast<const uint4 *>(X + (size_t)tb * n + c0 + 8 * v);
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[(8 * v + e) * 33 + tp] = (uint32_t)r0.h[e] | ((uint32_t)r1.h[e] << 16);
    }
    __syncthreads();
    {
        const int c = threadIdx.x >> 2, q = threadIdx.x & 3;           // channel, group of 16 tokens
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = tile[c * 33 + 8 * q + j];
        uint4 *dst = reinterpret_cast<uint4 *>(Xt + (size_t)(c0 + c) * Tp + t0 + 16 * q);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// persistent SYRK kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
hessian_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float *H,
                  int n, int nkb, int ntiles, float alpha, float beta, uint32_t idesc, int mn_major) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers &bar = *reinterpret_cast<Barriers *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntn = n / BN;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar.tmem_full[b], 1); mbar_init(&bar.tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int tm, tn;
                tile_coords(t, ntn, tm, tn);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar.empty[stage], phase ^ 1);
                    uint8_t *a = smem + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&bar.full[stage], STAGE_BYTES);
                    if (mn_major) {
                        // X itself (tokens x channels): boxes of 64 tokens x 64 channels (128 B rows) -- two of them side by side
                        // for the A tile's 128 channels, four for B's 256; tokens past T are zero-filled by the TMA unit
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c) tma_load_2d(a + c * 8192, &map_a, &bar.full[stage], tm * BM + 64 * c, kb * BK);
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c) tma_load_2d(a + A_BYTES + c * 8192, &map_a, &bar.full[stage], tn * BN + 64 * c, kb * BK);
                    } else {
                        tma_load_2d(a, &map_a, &bar.full[stage], kb * BK, tm * BM);
                        tma_load_2d(a + A_BYTES, &map_b, &bar.full[stage], kb * BK, tn * BN);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0, phase = 0, it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    if (mn_major) {
                        // MN-major operands: 64 channels x 8 tokens per 1024-byte swizzle atom, token groups 1024 B apart (SBO),
                        // the next 64 channels 8192 B apart (LBO); a K step of 16 tokens = two token groups = 2048 B
                        const uint64_t adesc = make_mnmajor_sw128_desc(a_addr), bdesc = make_mnmajor_sw128_desc(a_addr + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            tc_mma_f16(tmem_d, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                    } else {
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                        const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)   // +32 B per K step inside the 128 B swizzle span
                            tc_mma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    tc_commit(&bar.empty[stage]);            // frees the smem stage when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bar.tmem_full[buf]);              // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> H =====
        const int q = warp & 3;     // TMEM lane quadrant of this warp
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            int tm, tn;
            tile_coords(t, ntn, tm, tn);
            const int buf = it & 1;
            mbar_wait(&bar.tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const int i = tm * BM + q * 32 + lane;            // global row of this thread
            const int i_lo = tm * BM + q * 32, i_hi = i_lo + 31;
            float *hrow = H + (size_t)i * n;
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                const int j0 = tn * BN + ch * 32;
                if (j0 + 31 < i_lo) continue;                 // chunk entirely below the diagonal: mirrored from elsewhere
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + ch * 32), v);
                const bool all_upper = (j0 >= i_hi);          // every (i, j) of this warp's chunk has j >= i
                if (all_upper) {
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (beta != 0.0f) h = *reinterpret_cast<const float4 *>(hrow + j0 + 4 * c4);
                        float4 o;
                        o.x = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 0]), __fmul_rn(beta, h.x));
                        o.y = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 1]), __fmul_rn(beta, h.y));
                        o.z = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 2]), __fmul_rn(beta, h.z));
                        o.w = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 3]), __fmul_rn(beta, h.w));
                        *reinterpret_cast<float4 *>(hrow + j0 + 4 * c4) = o;
                        v[4 * c4 + 0] = __float_as_uint(o.x); v[4 * c4 + 1] = __float_as_uint(o.y);
                        v[4 * c4 + 2] = __float_as_uint(o.z); v[4 * c4 + 3] = __float_as_uint(o.w);
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c)             // mirror: lanes write consecutive addresses
                        if (j0 + c != i) H[(size_t)(j0 + c) * n + i] = __uint_as_float(v[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int j = j0 + c;
                        if (j >= i) {
                            const float h = (beta != 0.0f) ? hrow[j] : 0.0f;
                            const float o = __fmaf_rn(alpha, __uint_as_float(v[c]), __fmul_rn(beta, h));
                            hrow[j] = o;
                            if (j != i) H[(size_t)j * n + i] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.tmem_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// cta_group::2 variant (the default; GQ_HESSIAN_2CTA=0 selects the one-CTA kernel above): a CTA PAIR owns a (256 x 256) tile.  Each CTA stages its own 128 channels of A and
// HALF of B (128 of the 256 channels): 32 KB per stage instead of 48 KB, i.e. a third less L2 -> shared-memory traffic per
// flop, and a 6-stage ring.  The leader CTA's one thread issues tcgen05.mma.cta_group::2 (M = 256: each CTA's tensor core
// computes its 128 rows into its own TMEM, reading the other half of B from the peer's shared memory); TMA loads of both CTAs
// complete on the LEADER's full barrier, the MMA commits are multicast to both CTAs' empty / accumulator-full barriers, and the
// epilogue warps of both CTAs release the accumulator on the leader's barrier.  Measured on B200 per 16384 tokens: n = 4096
// 0.251 -> 0.226 ms, n = 14336 2.98 -> 2.67 ms (1.26 PFLOP/s of upper-triangle flops).
// ---------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6;
constexpr int STAGE2_BYTES = 4 * 8192;            // 2 boxes (64 tokens x 64 channels) of A + 2 of B
struct Barriers2 {
    uint64_t full[STAGES2];
    uint64_t empty[STAGES2];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};
constexpr size_t SMEM2_BYTES = 1024 + (size_t)STAGES2 * STAGE2_BYTES + sizeof(Barriers2);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s counterpart in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_2cta(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint64_t *bar) {      // arrives on `bar` (same offset) in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// pair tiles: (tm2, tn) over 256-row / 256-column blocks with tn >= tm2 (upper triangle)
__device__ __forceinline__ void tile_coords2(int idx, int ntn, int &tm2, int &tn) {
    int m = 0;
    while (idx >= ntn - m) { idx -= ntn - m; ++m; }
    tm2 = m;
    tn = m + idx;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
hessian_tc2_kernel(const __grid_constant__ CUtensorMap map_x, float *H, int n, int nkb, int ntiles, float alpha, float beta,
                   uint32_t idesc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers2 &bar = *reinterpret_cast<Barriers2 *>(smem + (size_t)STAGES2 * STAGE2_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntn = n / 256;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar.tmem_full[b], 1); mbar_init(&bar.tmem_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer (both CTAs; transactions complete on the leader's full barrier) =====
            int stage = 0, phase = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                int tm2, tn;
                tile_coords2(t, ntn, tm2, tn);
                const int a_ch = tm2 * 256 + (int)rank * 128, b_ch = tn * 256 + (int)rank * 128;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar.empty[stage], phase ^ 1);
                    uint8_t *a = smem + (size_t)stage * STAGE2_BYTES;
                    const uint32_t full0 = mapa_u32(&bar.full[stage], 0);
                    if (rank == 0) mbar_expect_tx(&bar.full[stage], 2 * STAGE2_BYTES);
                    tma_load_2d_2cta(a, &map_x, full0, a_ch, kb * BK);
                    tma_load_2d_2cta(a + 8192, &map_x, full0, a_ch + 64, kb * BK);
                    tma_load_2d_2cta(a + 16384, &map_x, full0, b_ch, kb * BK);
                    tma_load_2d_2cta(a + 24576, &map_x, full0, b_ch + 64, kb * BK);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {   // ===== MMA issuer (leader CTA only) =====
            int stage = 0, phase = 0, it = 0;
            for (int t = pair; t < ntiles; t += npairs, ++it) {
                const int buf = it & 1;
                mbar_wait(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 256;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE2_BYTES);
                    const uint64_t adesc = make_mnmajor_sw128_desc(a_addr), bdesc = make_mnmajor_sw128_desc(a_addr + 16384);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        tc_mma_f16_2cta(tmem_d, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                    tc_commit_2cta(&bar.empty[stage]);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                tc_commit_2cta(&bar.tmem_full[buf]);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own TMEM (128 rows of the pair's tile) -> H, direct + mirrored =====
        const int q = warp & 3;
        int it = 0;
        for (int t = pair; t < ntiles; t += npairs, ++it) {
            int tm2, tn;
            tile_coords2(t, ntn, tm2, tn);
            const int buf = it & 1;
            mbar_wait(&bar.tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const int i_lo = tm2 * 256 + (int)rank * 128 + q * 32, i = i_lo + lane, i_hi = i_lo + 31;
            float *hrow = H + (size_t)i * n;
#pragma unroll 1
            for (int ch = 0; ch < 256 / 32; ++ch) {
                const int j0 = tn * 256 + ch * 32;
                if (j0 + 31 < i_lo) continue;
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + ch * 32), v);
                const bool all_upper = (j0 >= i_hi);
                if (all_upper) {
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (beta != 0.0f) h = *reinterpret_cast<const float4 *>(hrow + j0 + 4 * c4);
                        float4 o;
                        o.x = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 0]), __fmul_rn(beta, h.x));
                        o.y = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 1]), __fmul_rn(beta, h.y));
                        o.z = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 2]), __fmul_rn(beta, h.z));
                        o.w = __fmaf_rn(alpha, __uint_as_float(v[4 * c4 + 3]), __fmul_rn(beta, h.w));
                        *reinterpret_cast<float4 *>(hrow + j0 + 4 * c4) = o;
                        v[4 * c4 + 0] = __float_as_uint(o.x); v[4 * c4 + 1] = __float_as_uint(o.y);
                        v[4 * c4 + 2] = __float_as_uint(o.z); v[4 * c4 + 3] = __float_as_uint(o.w);
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (j0 + c != i) H[(size_t)(j0 + c) * n + i] = __uint_as_float(v[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int j = j0 + c;
                        if (j >= i) {
                            const float h = (beta != 0.0f) ? hrow[j] : 0.0f;
                            const float o = __fmaf_rn(alpha, __uint_as_float(v[c]), __fmul_rn(beta, h));
                            hrow[j] = o;
                            if (j != i) H[(size_t)j * n + i] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&bar.tmem_empty[buf], 0));
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// GQ_HESSIAN_MN=1: feed X (tokens x channels) to the tensor cores as MN-major operands, without the transposed copy
// (read on every call: tests flip it).
bool hessian_mn_major() {
    const char *e = getenv("GQ_HESSIAN_MN");
    return e ? e[0] == '1' : GQ_HESSIAN_MN_DEFAULT;
}

bool hessian_2cta() {      // GQ_HESSIAN_2CTA=0: one CTA per tile (round 2's first MN-major kernel); default: the cta_group::2 kernel
    const char *e = getenv("GQ_HESSIAN_2CTA");
    return e ? e[0] == '1' : true;
}

bool make_map(CUtensorMap *m, void *base, int dtype, uint64_t rows, uint64_t cols_padded, uint32_t box_rows) {
    const CUtensorMapDataType dt = dtype == GQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    return make_map_2d(m, base, dt, 2, rows, cols_padded, BK, box_rows);
}

}  // namespace

size_t gq_hessian_tc_workspace_bytes(long n_tok, int d_col) {
    if (hessian_mn_major()) return 0;      // the tensor map covers X itself
    const long Tp = (n_tok + BK - 1) / BK * BK;
    return (size_t)d_col * (size_t)Tp * 2 + 1024;
}

bool gq_hessian_tc_supported(long n_tok, int d_col, int x_dtype) {
    return (x_dtype == GQ_BF16 || x_dtype == GQ_F16) && d_col % BN == 0 && n_tok > 0;
}

int gq_hessian_tc(float *H, const void *X, long n_tok, int d_col, int x_dtype, float beta, float alpha, void *workspace,
                  size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < gq_hessian_tc_workspace_bytes(n_tok, d_col) || (workspace == nullptr && !hessian_mn_major())) {
        gq_set_error("gq_hessian_update: workspace %zu < %zu bytes", ws_bytes, gq_hessian_tc_workspace_bytes(n_tok, d_col));
        return GQ_ERR_WORKSPACE;
    }
    const int n = d_col, T = (int)n_tok;
    const int Tp = (T + BK - 1) / BK * BK;
    const bool mn = hessian_mn_major();
    CUtensorMap map_a, map_b;
    if (mn) {
        // no transposed copy: the tensor map covers X itself, (T rows of n channels), boxes of 64 channels x 64 tokens
        const CUtensorMapDataType dt = x_dtype == GQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        if (!make_map_2d(&map_a, const_cast<void *>(X), dt, 2, (uint64_t)T, (uint64_t)n, 64, 64)) {
            gq_set_error("gq_hessian_update: cuTensorMapEncodeTiled failed");
            return GQ_ERR_CUDA;
        }
        map_b = map_a;
    } else {
        uint16_t *Xt = reinterpret_cast<uint16_t *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
        dim3 grid(Tp / 64, n / 64);
        transpose16_kernel<<<grid, 256, 0, st>>>((const uint16_t *)X, Xt, T, n, Tp);
        gq_count_launches(1);
        if (!make_map(&map_a, Xt, x_dtype, (uint64_t)n, (uint64_t)Tp, BM) || !make_map(&map_b, Xt, x_dtype, (uint64_t)n, (uint64_t)Tp, BN)) {
            gq_set_error("gq_hessian_update: cuTensorMapEncodeTiled failed");
            return GQ_ERR_CUDA;
        }
    }
    if (mn && hessian_2cta()) {
        const int ntn2 = n / 256;
        const int npair_tiles = ntn2 * (ntn2 + 1) / 2;
        const uint32_t fmt2 = x_dtype == GQ_BF16 ? 1u : 0u;
        // kind::f16, D = F32, A / B = fmt (MN-major), N = 256, M = 256 (the CTA pair's tile)
        const uint32_t idesc2 = (1u << 4) | (fmt2 << 7) | (fmt2 << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) |
                                ((uint32_t)(256 >> 4) << 24);
        GQ_CHECK_CUDA(cudaFuncSetAttribute(hessian_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES));
        const int max_pairs = num_sms() / 2;
        const int pairs = npair_tiles < max_pairs ? npair_tiles : max_pairs;
        hessian_tc2_kernel<<<2 * pairs, NTHREADS, SMEM2_BYTES, st>>>(map_a, H, n, Tp / BK, npair_tiles, alpha, beta, idesc2);
        gq_count_launches(1);
        GQ_CHECK_CUDA(cudaGetLastError());
        return GQ_OK;
    }
    const int ntm = n / BM, ntn = n / BN;
    int ntiles = 0;
    for (int m = 0; m < ntm; ++m) ntiles += ntn - (m >> 1);
    const uint32_t fmt = x_dtype == GQ_BF16 ? 1u : 0u;
    // kind::f16 instruction descriptor: D = F32, A/B = fmt, both K-major, N = 256, M = 128
    uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    if (mn) idesc |= (1u << 15) | (1u << 16);      // A and B are MN-major (the channel index is the contiguous one)
    GQ_CHECK_CUDA(cudaFuncSetAttribute(hessian_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    hessian_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(map_a, map_b, H, n, Tp / BK, ntiles, alpha, beta, idesc, mn ? 1 : 0);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
