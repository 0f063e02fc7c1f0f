// sgemm.cuh -- fp32 SIMT tile GEMM used by the Cholesky chain (prepare.cu) and by the any-dtype Hessian
// path (hessian.cu).  128x128x16 CTA tile, 256 threads, 8x8 accumulators per thread, explicit FMA.
//
//   C[m][n] = alpha * sum_k A(m,k) * B(k,n) + beta * C[m][n]
//   A(m,k) = A[m*a_rs + k*a_cs],  B(k,n) = B[k*b_rs + n*b_cs]   (one of the two strides of each must be 1)
//
// M and N must be multiples of 128; K arbitrary (zero padded).  Batched through gridDim.z.
#pragma once
#ifndef SIMT_EMU      // the CPU suite runs sgemm_kernel on an emulator (tests/helpers/simt_emu)
#include "common.cuh"
#endif

namespace sg {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

enum TileMode {
    TM_FULL = 0,
    TM_LOWER = 1,        // only tiles with tile_m >= tile_n are computed (SYRK on the lower triangle)
    TM_UPPER_MIRROR = 2  // only tiles with tile_n >= tile_m; the result is also stored transposed (symmetric C)
};
enum KMode {
    KM_FULL = 0,
    KM_FROM_N = 1,   // B is lower triangular: k starts at the tile's first n
    KM_TO_M = 2      // A is lower triangular: k ends after the tile's last m
};

struct Args {
    const void *A;
    const void *B;
    float *C;
    long a_rs, a_cs, b_rs, b_cs, ldc;
    long a_batch, b_batch, c_batch;   // element strides between batches
    int M, N, K;
    float alpha, beta;
    int tile_mode, k_mode;
    int in_dtype;   // dtype of A and B (GQ_F32 for the Cholesky chain)
};

__device__ __forceinline__ float ld_elem(const void *p, long idx, int dtype) { return load_as_f32(p, idx, dtype); }

// Fetch a (128 x 16) operand tile into registers (8 per thread), then commit it to smem as T[k][x]
// (x = m or n).  sx/sk: element strides along x and k.  Split so the global loads overlap the FMA loop.
__device__ __forceinline__ void fetch_tile(float (&r)[8], const void *P, int dtype, long sx, long sk, int x0, int k0,
                                           int K, int tid) {
    if (sx == 1) {   // contiguous along x: lanes walk x
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
            const int id = tid + NT * pass, k = id >> 7, x = id & 127;
            r[pass] = (k0 + k < K) ? ld_elem(P, (long)(x0 + x) + (long)(k0 + k) * sk, dtype) : 0.0f;
        }
    } else {         // contiguous along k: lanes walk k (16 consecutive k per x)
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
            const int id = tid + NT * pass, x = id >> 4, k = id & 15;
            r[pass] = (k0 + k < K) ? ld_elem(P, (long)(x0 + x) * sx + (long)(k0 + k), dtype) : 0.0f;
        }
    }
}
__device__ __forceinline__ void commit_tile(float (*T)[BM + PAD], const float (&r)[8], long sx, int tid) {
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
        const int id = tid + NT * pass;
        if (sx == 1) T[id >> 7][id & 127] = r[pass];
        else T[id & 15][id >> 4] = r[pass];
    }
}

__global__ void __launch_bounds__(NT) sgemm_kernel(const Args a) {
    __shared__ float As[2][BK][BM + PAD];
    __shared__ float Bs[2][BK][BN + PAD];
    const int tm = blockIdx.y, tn = blockIdx.x, bz = blockIdx.z;
    if (a.tile_mode == TM_LOWER && tm < tn) return;
    if (a.tile_mode == TM_UPPER_MIRROR && tn < tm) return;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = tm * BM, n0 = tn * BN;
    const char *Ab = (const char *)a.A, *Bb = (const char *)a.B;
    const int esz = a.in_dtype == GQ_F32 ? 4 : 2;
    const void *A = Ab + (size_t)bz * a.a_batch * esz;
    const void *B = Bb + (size_t)bz * a.b_batch * esz;
    float *C = a.C + (size_t)bz * a.c_batch;

    int kb = 0, ke = a.K;
    if (a.k_mode == KM_FROM_N) kb = n0;
    if (a.k_mode == KM_TO_M) ke = min(a.K, m0 + BM);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int nk = (ke - kb + BK - 1) / BK;
    float ra[8], rb[8];
    if (nk > 0) {
        fetch_tile(ra, A, a.in_dtype, a.a_rs, a.a_cs, m0, kb, ke, tid);
        fetch_tile(rb, B, a.in_dtype, a.b_cs, a.b_rs, n0, kb, ke, tid);
        commit_tile(As[0], ra, a.a_rs, tid);
        commit_tile(Bs[0], rb, a.b_cs, tid);
    }
    __syncthreads();
    for (int it = 0; it < nk; ++it) {
        const int cur = it & 1;
        if (it + 1 < nk) {
            fetch_tile(ra, A, a.in_dtype, a.a_rs, a.a_cs, m0, kb + (it + 1) * BK, ke, tid);
            fetch_tile(rb, B, a.in_dtype, a.b_cs, a.b_rs, n0, kb + (it + 1) * BK, ke, tid);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < nk) {   // the other buffer was last read in iteration it-1, which every thread has left
            commit_tile(As[cur ^ 1], ra, a.a_rs, tid);
            commit_tile(Bs[cur ^ 1], rb, a.b_cs, tid);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            float *c = C + (size_t)m * a.ldc + n;
            float v = __fmul_rn(a.alpha, acc[i][j]);
            if (a.beta != 0.0f) v = __fmaf_rn(a.beta, *c, v);
            *c = v;
            if (a.tile_mode == TM_UPPER_MIRROR && tm != tn) C[(size_t)n * a.ldc + m] = v;
        }
    }
}

#ifndef SIMT_EMU
inline int launch(const Args &a, int batch, cudaStream_t st) {
    dim3 grid(a.N / BN, a.M / BM, batch);
    sgemm_kernel<<<grid, NT, 0, st>>>(a);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
#endif  // SIMT_EMU

}  // namespace sg
