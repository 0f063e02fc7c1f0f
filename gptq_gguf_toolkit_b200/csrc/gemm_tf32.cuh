// gemm_tf32.cuh -- internal interface of the 3xTF32 tcgen05 GEMM (gemm_tf32.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace tg {

enum TileMode { TM_FULL = 0, TM_LOWER = 1 };           // TM_LOWER: only tiles with tile_m >= tile_n (needs M == N)
enum KMode {
    KM_FULL = 0,
    KM_FROM_M = 1,   // A[m][k] == 0 for k < m  (A upper triangular): k starts at the tile's first m
    KM_TO_M = 2,     // A[m][k] == 0 for k > m  (A lower triangular): k ends after the tile's last m
    KM_FROM_N = 3    // B[n][k] == 0 for k < n  (B upper triangular): k starts at the tile's first n
};

// C[b][m][n] = alpha * sum_k A[b][m][k] * B[b][n][k] + beta * C[b][m][n];  M, N multiples of 128.
struct GemmArgs {
    const float *A; long lda; long a_batch;   // (M x K) row-major, element stride between batches
    const float *B; long ldb; long b_batch;   // (N x K) row-major
    float *C; long ldc; long c_batch;
    int M, N, K, batch;
    float alpha, beta;
    int tile_mode, k_mode;
    bool same_ab;                             // B is A (SYRK): the operand is split once; gemm_f16x3_nt also takes N < M (B = A's first N rows)
};

// An operand whose hi / lo parts already exist: (rows x pitch) fp32 arrays; the block used starts at (row0, k0).
struct PreSplit { const float *hi; const float *lo; long pitch; long rows; int row0; int k0; };

size_t workspace_bytes(int M, int N, int K, int batch, bool same_ab);
// M: multiple of 128 (rows of the A arrays); only rows < m_valid of C are touched.
int gemm_tf32x3_nt_presplit(const PreSplit &A, const PreSplit &B, float *C, long ldc, int M, int m_valid, int N, int K, float alpha,
                            float beta, cudaStream_t st);
int gemm_tf32x3_nt(const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace tg
