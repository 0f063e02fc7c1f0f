// gemm_tf32.cu -- fp32-accurate GEMM on tcgen05 tensor cores by error-compensated TF32 ("3xTF32"):
//
//     C[m][n] = alpha * sum_k A[m][k] * B[n][k] + beta * C[m][n]        (NT: both operands K-contiguous)
//
// Each fp32 operand x is split once into hi = x with the low 13 mantissa bits cleared (exactly a TF32 value) and
// lo = x - hi (exact in fp32); the kernel accumulates hi*hi + hi*lo + lo*hi in fp32 TMEM accumulators, which
// recovers ~21 mantissa bits per product (the dropped lo*lo term is < 2^-22 relative).
//
// Used by the Cholesky chain (csrc/linalg.cu: panel solves, trailing SYRK updates, triangular-inverse merges);
// the zero tiles of triangular operands are skipped through k-ranges per tile.
//
// Structure = hessian_tc.cu: persistent CTA per SM, warp-specialised (TMA producer / one-thread MMA issuer /
// eight epilogue warps with a shared-memory transpose so that C moves in full 128-byte lines and is prefetched
// underneath the MMAs), 3-stage mbarrier ring of 64 KB stages (A_hi, A_lo, B_hi, B_lo tiles of 128 x 32 fp32,
// 128B-swizzled), M128 x N128 x K8 kind::tf32 MMAs, two 128-column TMEM accumulators.
#include "gemm_tf32.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, UMMA_K = 8;
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // A_hi, A_lo, B_hi, B_lo
constexpr int NTHREADS = 384;                      // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int EPI_WARPS = 8;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;       // one 32 x 32 fp32 transpose tile per epilogue warp
constexpr uint32_t TMEM_COLS = 256;

struct Barriers {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};
constexpr size_t BAR_BYTES = 128;
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block");
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + BAR_BYTES + (size_t)EPI_WARPS * EPI_STAGE_BYTES;

struct KParams {
    float *C;
    long ldc, c_batch;
    int M, N, nkb, batch, tiles_per_batch, ntn;
    int a_row0, b_row0, a_kb0, b_kb0;      // offsets of the operands inside the arrays the tensor maps cover
    int m_valid;                           // rows m >= m_valid of C are neither read nor written (ragged M)
    float alpha, beta;
    int tile_mode, k_mode;
    uint32_t idesc;
};

__device__ __forceinline__ void decode_tile(const KParams &p, int idx, int &b, int &tm, int &tn, int &kb0, int &kb1) {
    b = idx / p.tiles_per_batch;
    int r = idx - b * p.tiles_per_batch;
    if (p.tile_mode == tg::TM_LOWER) {
        int m = 0;
        while (r > m) { r -= m + 1; ++m; }
        tm = m; tn = r;
    } else {
        tm = r / p.ntn; tn = r - tm * p.ntn;
    }
    kb0 = 0; kb1 = p.nkb;
    if (p.k_mode == tg::KM_FROM_M) kb0 = tm * (BM / BK);
    else if (p.k_mode == tg::KM_TO_M) kb1 = min(p.nkb, (tm + 1) * (BM / BK));
    else if (p.k_mode == tg::KM_FROM_N) kb0 = tn * (BN / BK);
}

// hi/lo split of a strided fp32 operand into dense (rows*batch, Kp) arrays, zero padded in k
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ src, long ld, long batch_stride, int rows,
                                                         int K, int Kp, int batch, float *__restrict__ hi, float *__restrict__ lo) {
    const long total4 = (long)batch * rows * (Kp / 4);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const int k4 = (int)(i % (Kp / 4));
        const long rb = i / (Kp / 4);
        const int r = (int)(rb % rows);
        const long b = rb / rows;
        const float *s = src + b * batch_stride + (long)r * ld + 4 * k4;
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = (4 * k4 + u < K) ? s[u] : 0.0f;
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x[0]) & 0xFFFFE000u); l.x = __fsub_rn(x[0], h.x);
        h.y = __uint_as_float(__float_as_uint(x[1]) & 0xFFFFE000u); l.y = __fsub_rn(x[1], h.y);
        h.z = __uint_as_float(__float_as_uint(x[2]) & 0xFFFFE000u); l.z = __fsub_rn(x[2], h.z);
        h.w = __uint_as_float(__float_as_uint(x[3]) & 0xFFFFE000u); l.w = __fsub_rn(x[3], h.w);
        *reinterpret_cast<float4 *>(hi + i * 4) = h;
        *reinterpret_cast<float4 *>(lo + i * 4) = l;
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                   const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, const KParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers &bar = *reinterpret_cast<Barriers *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = p.tiles_per_batch * p.batch;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bl) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar.tmem_full[b], 1); mbar_init(&bar.tmem_empty[b], EPI_WARPS); }
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
        if (lane == 0) {   // ===== TMA producer =====
            int stage = 0, phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int b, tm, tn, kb0, kb1;
                decode_tile(p, t, b, tm, tn, kb0, kb1);
                const int arow = p.a_row0 + b * p.M + tm * BM, brow = p.b_row0 + b * p.N + tn * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bar.empty[stage], phase ^ 1);
                    uint8_t *s = smem + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&bar.full[stage], STAGE_BYTES);
                    tma_load_2d(s, &map_ah, &bar.full[stage], (p.a_kb0 + kb) * BK, arow);
                    tma_load_2d(s + TILE_BYTES, &map_al, &bar.full[stage], (p.a_kb0 + kb) * BK, arow);
                    tma_load_2d(s + 2 * TILE_BYTES, &map_bh, &bar.full[stage], (p.b_kb0 + kb) * BK, brow);
                    tma_load_2d(s + 3 * TILE_BYTES, &map_bl, &bar.full[stage], (p.b_kb0 + kb) * BK, brow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer =====
            int stage = 0, phase = 0, it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                int b, tm, tn, kb0, kb1;
                decode_tile(p, t, b, tm, tn, kb0, kb1);
                const int buf = it & 1;
                mbar_wait(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bar.full[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint64_t ah = make_kmajor_sw128_desc(s), al = make_kmajor_sw128_desc(s + TILE_BYTES);
                    const uint64_t bh = make_kmajor_sw128_desc(s + 2 * TILE_BYTES), bl = make_kmajor_sw128_desc(s + 3 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {     // 8 fp32 = 32 B per K step inside the 128 B swizzle span
                        const uint64_t o = (uint64_t)(2 * k);
                        tc_mma_tf32(tmem_d, al + o, bh + o, p.idesc, (kb > kb0) || (k > 0));   // small terms first
                        tc_mma_tf32(tmem_d, ah + o, bl + o, p.idesc, 1);
                        tc_mma_tf32(tmem_d, ah + o, bh + o, p.idesc, 1);
                    }
                    tc_commit(&bar.empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bar.tmem_full[buf]);
            }
        }
    } else if (warp >= 4) {   // ===== epilogue =====
        // Warp w may read TMEM lanes 32*(w%4)..+31 (= 32 rows of the tile); the eight warps split the tile's four
        // 32-column chunks two each.  tcgen05.ld hands every lane one ROW of a chunk; a swizzled 32 x 32 staging
        // tile in shared memory turns that into row-contiguous float4 so that C is read and written in full
        // 128-byte lines (4 rows per warp instruction).  The C values are requested BEFORE waiting for the
        // accumulator, i.e. they stream in underneath the tile's MMAs.
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        float *stage = reinterpret_cast<float *>(smem + (size_t)STAGES * STAGE_BYTES + BAR_BYTES) + ew * (EPI_STAGE_BYTES / 4);
        const int rr = lane >> 3, jj = lane & 7;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            int b, tm, tn, kb0, kb1;
            decode_tile(p, t, b, tm, tn, kb0, kb1);
            const int buf = it & 1;
            const int m0 = tm * BM + q * 32;
            float *cbase = p.C + (long)b * p.c_batch + (long)m0 * p.ldc + tn * BN + half * 64 + 4 * jj;
            float4 cpre[2][8];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = 4 * i8 + rr;
                    cpre[cc][i8] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.beta != 0.0f && m0 + r < p.m_valid)
                        cpre[cc][i8] = *reinterpret_cast<const float4 *>(cbase + (long)r * p.ldc + cc * 32);
                }
            mbar_wait(&bar.tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const bool empty = kb1 <= kb0;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 64 + cc * 32), v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                    __uint_as_float(v[4 * j + 3]));
                __syncwarp();
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = 4 * i8 + rr;
                    float4 a = *reinterpret_cast<const float4 *>(stage + r * 32 + ((jj ^ (r & 7)) << 2));
                    if (empty) a = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 h = cpre[cc][i8];
                    float4 o;
                    o.x = __fmaf_rn(p.alpha, a.x, __fmul_rn(p.beta, h.x));
                    o.y = __fmaf_rn(p.alpha, a.y, __fmul_rn(p.beta, h.y));
                    o.z = __fmaf_rn(p.alpha, a.z, __fmul_rn(p.beta, h.z));
                    o.w = __fmaf_rn(p.alpha, a.w, __fmul_rn(p.beta, h.w));
                    if (m0 + r < p.m_valid) *reinterpret_cast<float4 *>(cbase + (long)r * p.ldc + cc * 32) = o;
                }
                __syncwarp();
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

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

namespace tg {

size_t workspace_bytes(int M, int N, int K, int batch, bool same_ab) {
    const size_t Kp = align_up((size_t)K, BK);
    const size_t a = align_up((size_t)batch * M * Kp * 4, 1024), b = align_up((size_t)batch * N * Kp * 4, 1024);
    return 1024 + 2 * a + (same_ab ? 0 : 2 * b);
}

int gemm_tf32x3_nt(const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (g.M % BM || g.N % BN || g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) {
        gq_set_error("gemm_tf32x3_nt: M=%d N=%d must be positive multiples of 128 (K=%d, batch=%d)", g.M, g.N, g.K, g.batch);
        return GQ_ERR_INVALID;
    }
    if (g.tile_mode == TM_LOWER && g.M != g.N) {
        gq_set_error("gemm_tf32x3_nt: TM_LOWER needs M == N");
        return GQ_ERR_INVALID;
    }
    if (ws == nullptr || ws_bytes < workspace_bytes(g.M, g.N, g.K, g.batch, g.same_ab)) {
        gq_set_error("gemm_tf32x3_nt: workspace too small");
        return GQ_ERR_WORKSPACE;
    }
    const int Kp = (int)align_up((size_t)g.K, BK);
    uint8_t *base = reinterpret_cast<uint8_t *>(align_up((size_t)(uintptr_t)ws, 1024));
    const size_t abytes = align_up((size_t)g.batch * g.M * Kp * 4, 1024), bbytes = align_up((size_t)g.batch * g.N * Kp * 4, 1024);
    float *a_hi = (float *)base, *a_lo = (float *)(base + abytes);
    float *b_hi = g.same_ab ? a_hi : (float *)(base + 2 * abytes);
    float *b_lo = g.same_ab ? a_lo : (float *)(base + 2 * abytes + bbytes);
    auto grid_for = [](long n4) { long gsz = (n4 + 255) / 256; const long cap = 148L * 16; return (int)(gsz < 1 ? 1 : (gsz > cap ? cap : gsz)); };
    split_tf32_kernel<<<grid_for((long)g.batch * g.M * (Kp / 4)), 256, 0, st>>>(g.A, g.lda, g.a_batch, g.M, g.K, Kp, g.batch, a_hi, a_lo);
    gq_count_launches(1);
    if (!g.same_ab) {
        split_tf32_kernel<<<grid_for((long)g.batch * g.N * (Kp / 4)), 256, 0, st>>>(g.B, g.ldb, g.b_batch, g.N, g.K, Kp, g.batch, b_hi, b_lo);
        gq_count_launches(1);
    }
    CUtensorMap mah, mal, mbh, mbl;
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    bool ok = make_map_2d(&mah, a_hi, dt, 4, (uint64_t)g.batch * g.M, (uint64_t)Kp, BK, BM) &&
              make_map_2d(&mal, a_lo, dt, 4, (uint64_t)g.batch * g.M, (uint64_t)Kp, BK, BM) &&
              make_map_2d(&mbh, b_hi, dt, 4, (uint64_t)g.batch * g.N, (uint64_t)Kp, BK, BN) &&
              make_map_2d(&mbl, b_lo, dt, 4, (uint64_t)g.batch * g.N, (uint64_t)Kp, BK, BN);
    if (!ok) {
        gq_set_error("gemm_tf32x3_nt: cuTensorMapEncodeTiled failed");
        return GQ_ERR_CUDA;
    }
    KParams p;
    p.a_row0 = p.b_row0 = p.a_kb0 = p.b_kb0 = 0;
    p.m_valid = g.M;
    p.C = g.C; p.ldc = g.ldc; p.c_batch = g.c_batch; p.M = g.M; p.N = g.N; p.nkb = Kp / BK; p.batch = g.batch;
    const int ntm = g.M / BM;
    p.ntn = g.N / BN;
    p.tiles_per_batch = g.tile_mode == TM_LOWER ? ntm * (ntm + 1) / 2 : ntm * p.ntn;
    p.alpha = g.alpha; p.beta = g.beta; p.tile_mode = g.tile_mode; p.k_mode = g.k_mode;
    // kind::tf32 instruction descriptor: D = F32, A/B = TF32, both K-major, N = 128, M = 128
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    GQ_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int ntiles = p.tiles_per_batch * g.batch;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    gemm_tf32x3_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(mah, mal, mbh, mbl, p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

// Same GEMM on operands the caller has already split (hi/lo arrays with their own pitch); the operand blocks start at
// (row0, k0) inside those arrays.  K and k0 must be multiples of 32, pitches multiples of 4 floats, arrays 16-byte aligned.
int gemm_tf32x3_nt_presplit(const PreSplit &A, const PreSplit &B, float *C, long ldc, int M, int m_valid, int N, int K, float alpha,
                            float beta, cudaStream_t st) {
    if (M % BM || N % BN || K % BK || A.k0 % BK || B.k0 % BK || M <= 0 || N <= 0 || K <= 0) {
        gq_set_error("gemm_tf32x3_nt_presplit: bad shape M=%d N=%d K=%d", M, N, K);
        return GQ_ERR_INVALID;
    }
    CUtensorMap mah, mal, mbh, mbl;
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    bool ok = make_map_2d(&mah, (void *)A.hi, dt, 4, (uint64_t)A.rows, (uint64_t)A.pitch, BK, BM) &&
              make_map_2d(&mal, (void *)A.lo, dt, 4, (uint64_t)A.rows, (uint64_t)A.pitch, BK, BM) &&
              make_map_2d(&mbh, (void *)B.hi, dt, 4, (uint64_t)B.rows, (uint64_t)B.pitch, BK, BN) &&
              make_map_2d(&mbl, (void *)B.lo, dt, 4, (uint64_t)B.rows, (uint64_t)B.pitch, BK, BN);
    if (!ok) {
        gq_set_error("gemm_tf32x3_nt_presplit: cuTensorMapEncodeTiled failed");
        return GQ_ERR_CUDA;
    }
    KParams p;
    p.a_row0 = A.row0; p.b_row0 = B.row0; p.a_kb0 = A.k0 / BK; p.b_kb0 = B.k0 / BK;
    p.m_valid = m_valid;
    p.C = C; p.ldc = ldc; p.c_batch = 0; p.M = M; p.N = N; p.nkb = K / BK; p.batch = 1;
    p.ntn = N / BN;
    p.tiles_per_batch = (M / BM) * p.ntn;
    p.alpha = alpha; p.beta = beta; p.tile_mode = TM_FULL; p.k_mode = KM_FULL;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    GQ_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int grid = p.tiles_per_batch < num_sms() ? p.tiles_per_batch : num_sms();
    gemm_tf32x3_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(mah, mal, mbh, mbl, p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

}  // namespace tg

// test hook (exported; not part of the reference-facing API): plain fp32-accurate NT GEMM on the tensor cores
extern "C" GQ_API int gq_debug_gemm_tf32x3_nt(const float *A, long lda, const float *B, long ldb, float *C, long ldc, int M, int N,
                                              int K, int batch, long a_batch, long b_batch, long c_batch, float alpha, float beta,
                                              int tile_mode, int k_mode, void *ws, size_t ws_bytes, gq_stream_t stream) {
    tg::GemmArgs g;
    g.A = A; g.lda = lda; g.a_batch = a_batch; g.B = B; g.ldb = ldb; g.b_batch = b_batch; g.C = C; g.ldc = ldc; g.c_batch = c_batch;
    g.M = M; g.N = N; g.K = K; g.batch = batch; g.alpha = alpha; g.beta = beta; g.tile_mode = tile_mode; g.k_mode = k_mode;
    g.same_ab = (A == B && lda == ldb && a_batch == b_batch && M == N);
    return tg::gemm_tf32x3_nt(g, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" GQ_API size_t gq_debug_gemm_tf32x3_workspace(int M, int N, int K, int batch) { return tg::workspace_bytes(M, N, K, batch, false); }
