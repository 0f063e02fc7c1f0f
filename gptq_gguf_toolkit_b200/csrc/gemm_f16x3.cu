// gemm_f16x3.cu -- fp32-class GEMM on tcgen05 tensor cores from fp16 operand pairs ("3xFP16"):
//
//     C[m][n] = alpha * sum_k A[m][k] * B[n][k] + beta * C[m][n]        (NT: both operands K-contiguous)
//
// Every fp32 operand row is scaled by a power of two 2^e so that its largest magnitude lands in [2^14, 2^15), and each
// element is split into hi = fp16(x 2^e) and lo = fp16(x 2^e - hi): 22 significant bits, as many as the hi/lo pair of the
// 3xTF32 scheme keeps.  The kernel accumulates hi*hi + hi*lo + lo*hi with kind::f16 MMAs (fp16 products are exact in the fp32
// TMEM accumulators) and undoes the row scalings of A (per accumulator row) and B (per accumulator column) in the epilogue.
// Against 3xTF32 (gemm_tf32.cu) the operands take half the bytes (4 instead of 8 per element) and the MMAs run at twice the
// rate -- that GEMM is bound by the L2 -> shared-memory operand traffic of its 128 x 128 tiles (ncu, profiles/), which is
// what this kernel is built around:
//   * operand rows interleave hi and lo per 32-wide k block (32 hi | 32 lo = one 128-byte swizzle row), so ONE TMA box per
//     operand and stage brings both parts and the three MMAs of a k step only differ in the descriptors' k offsets;
//   * 128 x 256 output tiles (0.0117 operand rows per output against 0.0156 for 128 x 128), 4-stage ring of 48 KB stages;
//   * persistent CTA per SM, warp-specialised: TMA producer / one-thread MMA issuer / eight epilogue warps; two TMEM
//     accumulators (2 x 256 columns) so that a tile's epilogue -- C is read and written in full 128-byte lines through a
//     shared-memory transpose, its reads issued before the accumulator is waited for -- runs under the next tile's MMAs.
// Users: the rank-k update of GQ_MODE_FAST (gptq_layer.cu) and, with 128-wide tiles, the Cholesky chain (linalg.cu).
#include <cstdlib>

#include "gemm_f16x3.cuh"
#include "tc_common.cuh"


namespace {
using namespace tc;

constexpr int BM = 128, BK = 32, UMMA_K = 16;
constexpr int ROW_BYTES = 128;                     // 32 hi + 32 lo halves of one operand row and k block
constexpr int NTHREADS = 384;                      // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int EPI_WARPS = 8;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;       // one 32 x 32 fp32 transpose tile per epilogue warp
constexpr int MAX_STAGES = 6;

struct Barriers {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};
constexpr size_t BAR_BYTES = 256;
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block");

template <int BN> struct Cfg {
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int A_BYTES = BM * ROW_BYTES, B_BYTES = BN * ROW_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN;
    static constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + BAR_BYTES + (size_t)EPI_WARPS * EPI_STAGE_BYTES;
    static constexpr int CH = BN / 64;             // 32-column chunks per epilogue warp (the two warps of a lane quadrant split BN)
};
static_assert(Cfg<256>::SMEM_BYTES <= 232448, "shared memory budget");

struct KParams {
    float *C;
    long ldc, c_batch;
    int M, N, nkb, batch, tiles_per_batch, ntn;
    int a_row0, b_row0, a_kb0, b_kb0;      // offsets of the operand blocks inside the arrays the tensor maps cover
    int m_valid;                           // rows m >= m_valid of C are neither read nor written (ragged M)
    float alpha, beta;
    int tile_mode, k_mode;
    uint32_t idesc;
    const float *sa, *sb;                  // per-row descaling factors of the A / B arrays (indexed like their rows)
};

template <int BN>
__device__ __forceinline__ void decode_tile(const KParams &p, int idx, int &b, int &tm, int &tn, int &kb0, int &kb1) {
    b = idx / p.tiles_per_batch;
    int r = idx - b * p.tiles_per_batch;
    if (p.tile_mode == tg::TM_LOWER) {         // tile (tm, tn) holds an element with n <= m  <=>  tn * BN <= tm * BM + BM - 1
        int m = 0;
        while (true) {
            const int cnt = (m * BM + BM - 1) / BN + 1;
            if (r < cnt) break;
            r -= cnt;
            ++m;
        }
        tm = m; tn = r;
    } else {
        tm = r / p.ntn; tn = r - tm * p.ntn;
    }
    kb0 = 0; kb1 = p.nkb;
    if (p.k_mode == tg::KM_FROM_M) kb0 = tm * (BM / BK);
    else if (p.k_mode == tg::KM_TO_M) kb1 = min(p.nkb, (tm + 1) * (BM / BK));
    else if (p.k_mode == tg::KM_FROM_N) kb0 = tn * (BN / BK);
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const KParams p) {
    using cfg = Cfg<BN>;
    constexpr int STAGES = cfg::STAGES, STAGE_BYTES = cfg::STAGE_BYTES, A_BYTES = cfg::A_BYTES, CH = cfg::CH;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers &bar = *reinterpret_cast<Barriers *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = p.tiles_per_batch * p.batch;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar.tmem_full[b], 1); mbar_init(&bar.tmem_empty[b], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)), "r"(cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            int stage = 0, phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int b, tm, tn, kb0, kb1;
                decode_tile<BN>(p, t, b, tm, tn, kb0, kb1);
                const int arow = p.a_row0 + b * p.M + tm * BM, brow = p.b_row0 + b * p.N + tn * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bar.empty[stage], phase ^ 1);
                    uint8_t *s = smem + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&bar.full[stage], STAGE_BYTES);
                    tma_load_2d(s, &map_a, &bar.full[stage], (p.a_kb0 + kb) * 64, arow);            // 64 halves = hi | lo of 32 k's
                    tma_load_2d(s + A_BYTES, &map_b, &bar.full[stage], (p.b_kb0 + kb) * 64, brow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer =====
            int stage = 0, phase = 0, it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                int b, tm, tn, kb0, kb1;
                decode_tile<BN>(p, t, b, tm, tn, kb0, kb1);
                const int buf = it & 1;
                mbar_wait(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bar.full[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint64_t ad = make_kmajor_sw128_desc(s), bd = make_kmajor_sw128_desc(s + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {     // 16 halves = 32 B per K step; hi at +0 / +32 B, lo at +64 / +96 B
                        const uint64_t h = (uint64_t)(2 * k), l = (uint64_t)(2 * k + 4);
                        tc_mma_f16(tmem_d, ad + l, bd + h, p.idesc, (kb > kb0) || (k > 0));   // small terms first
                        tc_mma_f16(tmem_d, ad + h, bd + l, p.idesc, 1);
                        tc_mma_f16(tmem_d, ad + h, bd + h, p.idesc, 1);
                    }
                    tc_commit(&bar.empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bar.tmem_full[buf]);
            }
        }
    } else if (warp >= 4) {   // ===== epilogue =====
        // Warp w may read TMEM lanes 32*(w%4)..+31 (= 32 rows of the tile); the two warps of a lane quadrant take half of the
        // tile's columns each, in 32-column chunks.  tcgen05.ld hands every lane one ROW of a chunk (scaled there by the
        // row's 2^-e); a swizzled 32 x 32 staging tile in shared memory turns that into row-contiguous float4 so that C is
        // read and written in full 128-byte lines (4 rows per warp instruction).  Two chunks of C are always in flight, the
        // first two requested BEFORE the accumulator is waited for, i.e. they stream in underneath the tile's MMAs.
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        float *stage = reinterpret_cast<float *>(smem + (size_t)STAGES * STAGE_BYTES + BAR_BYTES) + ew * (EPI_STAGE_BYTES / 4);
        const int rr = lane >> 3, jj = lane & 7;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            int b, tm, tn, kb0, kb1;
            decode_tile<BN>(p, t, b, tm, tn, kb0, kb1);
            const int buf = it & 1;
            const int m0 = tm * BM + q * 32;
            const int n0 = tn * BN + half * (BN / 2);
            float *cbase = p.C + (long)b * p.c_batch + (long)m0 * p.ldc + n0 + 4 * jj;
            const float sa = p.sa ? p.sa[p.a_row0 + b * p.M + m0 + lane] : 1.0f;
            const float *sbp = p.sb ? p.sb + p.b_row0 + b * p.N + n0 + 4 * jj : nullptr;
            float4 cbuf[2][8];
            auto load_c = [&](float4 (&dst)[8], int cc) {
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = 4 * i8 + rr;
                    dst[i8] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.beta != 0.0f && m0 + r < p.m_valid)
                        dst[i8] = *reinterpret_cast<const float4 *>(cbase + (long)r * p.ldc + cc * 32);
                }
            };
            load_c(cbuf[0], 0);
            if (CH > 1) load_c(cbuf[1], 1);
            mbar_wait(&bar.tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const bool empty = kb1 <= kb0;
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2) + cc * 32), v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(__fmul_rn(sa, __uint_as_float(v[4 * j])), __fmul_rn(sa, __uint_as_float(v[4 * j + 1])),
                                    __fmul_rn(sa, __uint_as_float(v[4 * j + 2])), __fmul_rn(sa, __uint_as_float(v[4 * j + 3])));
                __syncwarp();
                float4 s4 = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
                if (sbp) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(sbp + cc * 32);
                    s4 = make_float4(__fmul_rn(p.alpha, t4.x), __fmul_rn(p.alpha, t4.y), __fmul_rn(p.alpha, t4.z), __fmul_rn(p.alpha, t4.w));
                }
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = 4 * i8 + rr;
                    float4 a = *reinterpret_cast<const float4 *>(stage + r * 32 + ((jj ^ (r & 7)) << 2));
                    if (empty) a = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 h = cbuf[cc & 1][i8];
                    float4 o;
                    o.x = __fmaf_rn(s4.x, a.x, __fmul_rn(p.beta, h.x));
                    o.y = __fmaf_rn(s4.y, a.y, __fmul_rn(p.beta, h.y));
                    o.z = __fmaf_rn(s4.z, a.z, __fmul_rn(p.beta, h.z));
                    o.w = __fmaf_rn(s4.w, a.w, __fmul_rn(p.beta, h.w));
                    if (m0 + r < p.m_valid) *reinterpret_cast<float4 *>(cbase + (long)r * p.ldc + cc * 32) = o;
                }
                __syncwarp();
                if (cc + 2 < CH) load_c(cbuf[cc & 1], cc + 2);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cfg::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// cta_group::2 variant for the big rank-k updates (TM_FULL / KM_FULL, M and N multiples of 256): a CTA PAIR owns a 256 x 256
// tile.  Each CTA stages its own 128 rows of A and HALF of B (128 of the 256 rows): 32 KB per stage instead of 48 KB -- a third
// less L2 -> shared-memory operand traffic per flop -- in a 6-stage ring; the leader's one thread issues
// tcgen05.mma.cta_group::2 (M = 256), TMA transactions of both CTAs complete on the leader's full barrier, MMA commits are
// multicast to both CTAs, and the epilogue warps of both CTAs (each on its own 128 rows in its own TMEM) release the
// accumulator on the leader's barrier.  Same protocol as hessian_tc2_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6, STAGE2_BYTES = 2 * BM * ROW_BYTES;       // A (128 rows) + half of B (128 rows)
constexpr size_t SMEM2_BYTES = 1024 + (size_t)STAGES2 * STAGE2_BYTES + BAR_BYTES + (size_t)EPI_WARPS * EPI_STAGE_BYTES;
static_assert(SMEM2_BYTES <= 232448, "shared memory budget (2-CTA kernel)");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
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
__device__ __forceinline__ void tc_commit_2cta(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gemm_f16x3_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const KParams p) {
    constexpr int BN = 256, CH = 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers &bar = *reinterpret_cast<Barriers *>(smem + (size_t)STAGES2 * STAGE2_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int ntn = p.N / 256, ntiles = (p.M / 256) * ntn;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar.tmem_full[b], 1); mbar_init(&bar.tmem_empty[b], 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer (both CTAs) =====
            int stage = 0, phase = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                const int tm2 = t / ntn, tn = t - tm2 * ntn;
                const int arow = p.a_row0 + tm2 * 256 + (int)rank * 128, brow = p.b_row0 + tn * 256 + (int)rank * 128;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&bar.empty[stage], phase ^ 1);
                    uint8_t *s = smem + (size_t)stage * STAGE2_BYTES;
                    const uint32_t full0 = mapa_u32(&bar.full[stage], 0);
                    if (rank == 0) mbar_expect_tx(&bar.full[stage], 2 * STAGE2_BYTES);
                    tma_load_2d_2cta(s, &map_a, full0, (p.a_kb0 + kb) * 64, arow);
                    tma_load_2d_2cta(s + BM * ROW_BYTES, &map_b, full0, (p.b_kb0 + kb) * 64, brow);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {   // ===== MMA issuer (leader CTA) =====
            int stage = 0, phase = 0, it = 0;
            for (int t = pair; t < ntiles; t += npairs, ++it) {
                const int buf = it & 1;
                mbar_wait(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&bar.full[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + (size_t)stage * STAGE2_BYTES);
                    const uint64_t ad = make_kmajor_sw128_desc(s), bd = make_kmajor_sw128_desc(s + BM * ROW_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t h = (uint64_t)(2 * k), l = (uint64_t)(2 * k + 4);
                        tc_mma_f16_2cta(tmem_d, ad + l, bd + h, p.idesc, (kb > 0) || (k > 0));
                        tc_mma_f16_2cta(tmem_d, ad + h, bd + l, p.idesc, 1);
                        tc_mma_f16_2cta(tmem_d, ad + h, bd + h, p.idesc, 1);
                    }
                    tc_commit_2cta(&bar.empty[stage]);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                tc_commit_2cta(&bar.tmem_full[buf]);
            }
        }
    } else if (warp >= 4) {   // ===== epilogue (both CTAs, each on its 128 rows of the pair's tile) =====
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        float *stage = reinterpret_cast<float *>(smem + (size_t)STAGES2 * STAGE2_BYTES + BAR_BYTES) + ew * (EPI_STAGE_BYTES / 4);
        const int rr = lane >> 3, jj = lane & 7;
        int it = 0;
        for (int t = pair; t < ntiles; t += npairs, ++it) {
            const int tm2 = t / ntn, tn = t - tm2 * ntn;
            const int buf = it & 1;
            const int m0 = tm2 * 256 + (int)rank * 128 + q * 32;
            const int n0 = tn * BN + half * (BN / 2);
            float *cbase = p.C + (long)m0 * p.ldc + n0 + 4 * jj;
            const float sa = p.sa ? p.sa[p.a_row0 + m0 + lane] : 1.0f;
            const float *sbp = p.sb ? p.sb + p.b_row0 + n0 + 4 * jj : nullptr;
            float4 cbuf[2][8];
            auto load_c = [&](float4 (&dst)[8], int cc) {
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = 4 * i8 + rr;
                    dst[i8] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.beta != 0.0f && m0 + r < p.m_valid)
                        dst[i8] = *reinterpret_cast<const float4 *>(cbase + (long)r * p.ldc + cc * 32);
                }
            };
            load_c(cbuf[0], 0);
            load_c(cbuf[1], 1);
            mbar_wait(&bar.tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2) + cc * 32), v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(__fmul_rn(sa, __uint_as_float(v[4 * j])), __fmul_rn(sa, __uint_as_float(v[4 * j + 1])),
                                    __fmul_rn(sa, __uint_as_float(v[4 * j + 2])), __fmul_rn(sa, __uint_as_float(v[4 * j + 3])));
                __syncwarp();
                float4 s4 = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
                if (sbp) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(sbp + cc * 32);
                    s4 = make_float4(__fmul_rn(p.alpha, t4.x), __fmul_rn(p.alpha, t4.y), __fmul_rn(p.alpha, t4.z), __fmul_rn(p.alpha, t4.w));
                }
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = 4 * i8 + rr;
                    const float4 a = *reinterpret_cast<const float4 *>(stage + r * 32 + ((jj ^ (r & 7)) << 2));
                    const float4 h = cbuf[cc & 1][i8];
                    float4 o;
                    o.x = __fmaf_rn(s4.x, a.x, __fmul_rn(p.beta, h.x));
                    o.y = __fmaf_rn(s4.y, a.y, __fmul_rn(p.beta, h.y));
                    o.z = __fmaf_rn(s4.z, a.z, __fmul_rn(p.beta, h.z));
                    o.w = __fmaf_rn(s4.w, a.w, __fmul_rn(p.beta, h.w));
                    if (m0 + r < p.m_valid) *reinterpret_cast<float4 *>(cbase + (long)r * p.ldc + cc * 32) = o;
                }
                __syncwarp();
                if (cc + 2 < CH) load_c(cbuf[cc & 1], cc + 2);
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
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
// Power-of-two scaling that puts mx (>= 0) into [2^14, 2^15): returns s = 2^e, inv = 2^-e.
__device__ __forceinline__ void row_scale(float mx, float &s, float &inv) {
    s = 1.0f; inv = 1.0f;
    if (mx > 0.0f && mx < 3.0e38f) {
        int ex;
        frexpf(mx, &ex);                        // mx = f * 2^ex, f in [0.5, 1)
        int e = 15 - ex;
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        s = ldexpf(1.0f, e);
        inv = ldexpf(1.0f, -e);
    }
}
__device__ __forceinline__ void split_f16(float xs, __half &hi, __half &lo) {
    hi = __float2half_rn(xs);
    lo = __float2half_rn(__fsub_rn(xs, __half2float(hi)));
}

// One warp per (batch, row): max |x| over the row, then hi / lo of x * 2^e in the interleaved layout.
// VEC: the row (K <= 1024, K % 4 == 0, 16-byte aligned) is read ONCE with 16-byte loads and stays in registers between the
// two passes; the outputs go out as 8-byte stores (4 hi resp. 4 lo halves per lane).
template <bool VEC>
__global__ void __launch_bounds__(256) split_rows_f16_kernel(const float *__restrict__ src, long ld, long batch_stride, int rows,
                                                            int rows_pad, int K, int Kp, int batch, __half *__restrict__ dst,
                                                            float *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const long nrows = (long)batch * rows_pad;
    for (long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nrows; w += (long)gridDim.x * (blockDim.x >> 5)) {
        const int r = (int)(w % rows_pad);
        const long b = w / rows_pad;
        __half *d = dst + w * (2L * Kp);
        if (r >= rows) {
            for (int k = lane; k < 2 * Kp; k += 32) d[k] = __float2half_rn(0.0f);
            if (lane == 0) scale[w] = 1.0f;
            continue;
        }
        const float *s = src + b * batch_stride + (long)r * ld;
        if (VEC) {
            float4 v[8];
            float mx = 0.0f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int k = 4 * (lane + 32 * t);
                v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < K) v[t] = *reinterpret_cast<const float4 *>(s + k);
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[t].x), fabsf(v[t].y))), fmaxf(fabsf(v[t].z), fabsf(v[t].w)));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sc, inv;
            row_scale(mx, sc, inv);
            if (lane == 0) scale[w] = inv;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int k = 4 * (lane + 32 * t);
                if (k < Kp) {
                    __half h[4], l[4];
                    split_f16(__fmul_rn(v[t].x, sc), h[0], l[0]);
                    split_f16(__fmul_rn(v[t].y, sc), h[1], l[1]);
                    split_f16(__fmul_rn(v[t].z, sc), h[2], l[2]);
                    split_f16(__fmul_rn(v[t].w, sc), h[3], l[3]);
                    __half *o = d + 64 * (k >> 5) + (k & 31);
                    *reinterpret_cast<uint2 *>(o) = *reinterpret_cast<const uint2 *>(h);
                    *reinterpret_cast<uint2 *>(o + 32) = *reinterpret_cast<const uint2 *>(l);
                }
            }
            continue;
        }
        float mx = 0.0f;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, fabsf(s[k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sc, inv;
        row_scale(mx, sc, inv);
        if (lane == 0) scale[w] = inv;
        for (int k0 = 0; k0 < Kp; k0 += 32) {
            const int k = k0 + lane;
            const float x = k < K ? __fmul_rn(s[k], sc) : 0.0f;
            __half hi, lo;
            split_f16(x, hi, lo);
            d[2 * k0 + lane] = hi;
            d[2 * k0 + 32 + lane] = lo;
        }
    }
}

// cmax[j] = max_k |U[k][j]| as uint bits (non-negative floats order like their bit patterns); cmax zeroed by the caller.
__global__ void __launch_bounds__(256) colmax_upper_kernel(const float *__restrict__ U, int n, unsigned int *cmax) {
    __shared__ float red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * 256;
    if (r0 > blockIdx.x * 32 + 31) return;               // rows below the diagonal are zero
    const int r1 = min(n, r0 + 256);
    float m = 0.0f;
    for (int r = r0 + ty; r < r1; r += 8) m = fmaxf(m, fabsf(U[(size_t)r * n + j]));
    red[ty][tx] = m;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i][tx]);
        atomicMax(cmax + j, __float_as_uint(m));
    }
}
// (32 k x 32 j) tiles of U -> rows j of the Split16 transpose; tiles entirely below the diagonal are skipped.
__global__ void __launch_bounds__(256) transpose_split_upper_f16_kernel(const float *__restrict__ U, int n,
                                                                       const unsigned int *__restrict__ cmax,
                                                                       __half *__restrict__ dst, float *__restrict__ scale) {
    __shared__ float t[32][33];
    const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (k0 > j0 + 31) return;
    for (int r = ty; r < 32; r += 8) t[r][tx] = U[(size_t)(k0 + r) * n + j0 + tx];
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        float sc, inv;
        row_scale(__uint_as_float(cmax[j0 + c]), sc, inv);
        if (blockIdx.x == 0 && tx == 0) scale[j0 + c] = inv;
        __half hi, lo;
        split_f16(__fmul_rn(t[tx][c], sc), hi, lo);
        __half *d = dst + (size_t)(j0 + c) * (2 * (size_t)n) + 2 * k0;
        d[tx] = hi;
        d[32 + tx] = lo;
    }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// CTA pairs (cta_group::2) for the pre-split rank-k updates?  GQ_GEMM_2CTA=0|1 forces it (read on every call); by default
// only when there are at least two waves of pair tiles: measured on B200 (profiles/r02/notes.md) the pair kernel gains 2-4 % on
// the big updates (down_proj, gate/up: 347 / 362 TFLOP/s) and loses as much on the small ones, whose few 256 x 256 tiles leave
// SM pairs idle in the last wave.
bool gemm_2cta(int pair_tiles) {
    const char *e = getenv("GQ_GEMM_2CTA");
    if (e) return e[0] == '1';
    return pair_tiles >= 2 * (tc::num_sms() / 2);
}

template <int BN> int launch(const CUtensorMap &ma, const CUtensorMap &mb, KParams &p, cudaStream_t st) {
    using cfg = Cfg<BN>;
    const int ntm = p.M / BM;
    p.ntn = p.N / BN;
    if (p.tile_mode == tg::TM_LOWER) {
        p.tiles_per_batch = 0;
        for (int m = 0; m < ntm; ++m) p.tiles_per_batch += (m * BM + BM - 1) / BN + 1;
    } else {
        p.tiles_per_batch = ntm * p.ntn;
    }
    // kind::f16 instruction descriptor: D = F32, A/B = F16, both K-major, N = BN, M = 128
    p.idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    GQ_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg::SMEM_BYTES));
    const int ntiles = p.tiles_per_batch * p.batch;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    gemm_f16x3_kernel<BN><<<grid, NTHREADS, cfg::SMEM_BYTES, st>>>(ma, mb, p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

int ew_grid_rows(long nrows) {          // 8 warps (rows) per CTA
    long g = (nrows + 7) / 8;
    const long cap = 148L * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

namespace th {

int split_rows_f16(const float *src, long ld, long batch_stride, int rows, int rows_pad, int K, int Kp, int batch, __half *dst,
                   float *scale, cudaStream_t st) {
    const bool vec = K <= 1024 && K % 4 == 0 && ld % 4 == 0 && batch_stride % 4 == 0 && ((uintptr_t)src % 16) == 0;
    if (vec)
        split_rows_f16_kernel<true><<<ew_grid_rows((long)batch * rows_pad), 256, 0, st>>>(src, ld, batch_stride, rows, rows_pad, K, Kp, batch, dst, scale);
    else
        split_rows_f16_kernel<false><<<ew_grid_rows((long)batch * rows_pad), 256, 0, st>>>(src, ld, batch_stride, rows, rows_pad, K, Kp, batch, dst, scale);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

int transpose_split_upper_f16(const float *U, int n, __half *dst, float *scale, unsigned int *cmax, cudaStream_t st) {
    if (n % 32) {
        gq_set_error("transpose_split_upper_f16: n=%d must be a multiple of 32", n);
        return GQ_ERR_INVALID;
    }
    GQ_CHECK_CUDA(cudaMemsetAsync(cmax, 0, (size_t)n * sizeof(unsigned int), st));
    colmax_upper_kernel<<<dim3(n / 32, (n + 255) / 256), 256, 0, st>>>(U, n, cmax);
    transpose_split_upper_f16_kernel<<<dim3(n / 32, n / 32), 256, 0, st>>>(U, n, cmax, dst, scale);
    gq_count_launches(2);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

int gemm_f16x3_nt_presplit(const Split16 &A, const Split16 &B, float *C, long ldc, int M, int m_valid, int N, int K, float alpha,
                           float beta, cudaStream_t st) {
    if (M % BM || N % 128 || K % BK || A.k0 % BK || B.k0 % BK || M <= 0 || N <= 0 || K <= 0) {
        gq_set_error("gemm_f16x3_nt_presplit: bad shape M=%d N=%d K=%d", M, N, K);
        return GQ_ERR_INVALID;
    }
    const bool wide = N % 256 == 0;
    const bool pair = wide && M % 256 == 0 && gemm_2cta((M / 256) * (N / 256));      // CTA pairs on 256 x 256 tiles (cta_group::2)
    CUtensorMap ma, mb;
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    bool ok = make_map_2d(&ma, (void *)A.data, dt, 2, (uint64_t)A.rows, (uint64_t)A.pitch, 64, BM) &&
              make_map_2d(&mb, (void *)B.data, dt, 2, (uint64_t)B.rows, (uint64_t)B.pitch, 64, pair ? 128 : (wide ? 256 : 128));
    if (!ok) {
        gq_set_error("gemm_f16x3_nt_presplit: cuTensorMapEncodeTiled failed");
        return GQ_ERR_CUDA;
    }
    KParams p;
    p.a_row0 = A.row0; p.b_row0 = B.row0; p.a_kb0 = A.k0 / BK; p.b_kb0 = B.k0 / BK;
    p.m_valid = m_valid;
    p.C = C; p.ldc = ldc; p.c_batch = 0; p.M = M; p.N = N; p.nkb = K / BK; p.batch = 1;
    p.alpha = alpha; p.beta = beta; p.tile_mode = tg::TM_FULL; p.k_mode = tg::KM_FULL;
    p.sa = A.scale; p.sb = B.scale;
    if (pair) {
        // kind::f16 instruction descriptor: D = F32, A/B = F16, both K-major, N = 256, M = 256 (the pair's tile)
        p.idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        p.ntn = N / 256; p.tiles_per_batch = (M / 256) * p.ntn;
        GQ_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16x3_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES));
        const int max_pairs = num_sms() / 2;
        const int pairs = p.tiles_per_batch < max_pairs ? p.tiles_per_batch : max_pairs;
        gemm_f16x3_2cta_kernel<<<2 * pairs, NTHREADS, SMEM2_BYTES, st>>>(ma, mb, p);
        gq_count_launches(1);
        GQ_CHECK_CUDA(cudaGetLastError());
        return GQ_OK;
    }
    return wide ? launch<256>(ma, mb, p, st) : launch<128>(ma, mb, p, st);
}

size_t workspace_bytes(int M, int N, int K, int batch, bool same_ab) {
    const size_t Kp = align_up((size_t)K, BK);
    const size_t a = align_up((size_t)batch * M * Kp * 4, 1024), b = align_up((size_t)batch * N * Kp * 4, 1024);
    const size_t sa = align_up((size_t)batch * M * 4, 1024), sb = align_up((size_t)batch * N * 4, 1024);
    return 1024 + a + sa + (same_ab ? 0 : b + sb);
}

int gemm_f16x3_nt(const tg::GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (g.M % BM || g.N % 128 || g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) {
        gq_set_error("gemm_f16x3_nt: M=%d N=%d must be positive multiples of 128 (K=%d, batch=%d)", g.M, g.N, g.K, g.batch);
        return GQ_ERR_INVALID;
    }
    if (g.tile_mode == tg::TM_LOWER && g.M != g.N) {
        gq_set_error("gemm_f16x3_nt: TM_LOWER needs M == N");
        return GQ_ERR_INVALID;
    }
    if (g.same_ab && (g.N > g.M || g.batch != 1) && g.N != g.M) {      // same_ab: B = the first N rows of A (N == M: SYRK)
        gq_set_error("gemm_f16x3_nt: same_ab needs N <= M (B is a row prefix of A) and, for N < M, batch == 1");
        return GQ_ERR_INVALID;
    }
    if (ws == nullptr || ws_bytes < workspace_bytes(g.M, g.N, g.K, g.batch, g.same_ab)) {
        gq_set_error("gemm_f16x3_nt: workspace too small");
        return GQ_ERR_WORKSPACE;
    }
    const int Kp = (int)align_up((size_t)g.K, BK);
    uint8_t *base = reinterpret_cast<uint8_t *>(align_up((size_t)(uintptr_t)ws, 1024));
    const size_t abytes = align_up((size_t)g.batch * g.M * Kp * 4, 1024), sabytes = align_up((size_t)g.batch * g.M * 4, 1024);
    const size_t bbytes = align_up((size_t)g.batch * g.N * Kp * 4, 1024);
    __half *a16 = (__half *)base;
    float *sa = (float *)(base + abytes);
    __half *b16 = g.same_ab ? a16 : (__half *)(base + abytes + sabytes);
    float *sb = g.same_ab ? sa : (float *)(base + abytes + sabytes + bbytes);
    int rc = split_rows_f16(g.A, g.lda, g.a_batch, g.M, g.M, g.K, Kp, g.batch, a16, sa, st);
    if (rc) return rc;
    if (!g.same_ab) {
        rc = split_rows_f16(g.B, g.ldb, g.b_batch, g.N, g.N, g.K, Kp, g.batch, b16, sb, st);
        if (rc) return rc;
    }
    // 256-wide tiles when N allows it and the problem is not a triangle of 128-blocks (whose k ranges are per 128 rows)
    const bool wide = g.N % 256 == 0 && g.tile_mode == tg::TM_FULL && g.k_mode != tg::KM_FROM_N;
    CUtensorMap ma, mb;
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    bool ok = make_map_2d(&ma, a16, dt, 2, (uint64_t)g.batch * g.M, (uint64_t)2 * Kp, 64, BM) &&
              make_map_2d(&mb, b16, dt, 2, (uint64_t)g.batch * (g.same_ab ? g.M : g.N), (uint64_t)2 * Kp, 64, wide ? 256 : 128);
    if (!ok) {
        gq_set_error("gemm_f16x3_nt: cuTensorMapEncodeTiled failed");
        return GQ_ERR_CUDA;
    }
    KParams p;
    p.a_row0 = p.b_row0 = p.a_kb0 = p.b_kb0 = 0;
    p.m_valid = g.M;
    p.C = g.C; p.ldc = g.ldc; p.c_batch = g.c_batch; p.M = g.M; p.N = g.N; p.nkb = Kp / BK; p.batch = g.batch;
    p.alpha = g.alpha; p.beta = g.beta; p.tile_mode = g.tile_mode; p.k_mode = g.k_mode;
    p.sa = sa; p.sb = sb;
    return wide ? launch<256>(ma, mb, p, st) : launch<128>(ma, mb, p, st);
}

}  // namespace th

// test hook (exported; not part of the reference-facing API): the same plain NT GEMM interface as gq_debug_gemm_tf32x3_nt
extern "C" GQ_API int gq_debug_gemm_f16x3_nt(const float *A, long lda, const float *B, long ldb, float *C, long ldc, int M, int N,
                                             int K, int batch, long a_batch, long b_batch, long c_batch, float alpha, float beta,
                                             int tile_mode, int k_mode, void *ws, size_t ws_bytes, gq_stream_t stream) {
    tg::GemmArgs g;
    g.A = A; g.lda = lda; g.a_batch = a_batch; g.B = B; g.ldb = ldb; g.b_batch = b_batch; g.C = C; g.ldc = ldc; g.c_batch = c_batch;
    g.M = M; g.N = N; g.K = K; g.batch = batch; g.alpha = alpha; g.beta = beta; g.tile_mode = tile_mode; g.k_mode = k_mode;
    g.same_ab = (A == B && lda == ldb && a_batch == b_batch && M == N);
    return th::gemm_f16x3_nt(g, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" GQ_API size_t gq_debug_gemm_f16x3_workspace(int M, int N, int K, int batch) { return th::workspace_bytes(M, N, K, batch, false); }
