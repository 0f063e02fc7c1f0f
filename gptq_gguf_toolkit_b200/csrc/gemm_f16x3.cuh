// gemm_f16x3.cuh -- internal interface of the split-fp16 ("3xFP16") tcgen05 GEMM (gemm_f16x3.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>

#include "gemm_tf32.cuh"      // tg::GemmArgs, TileMode, KMode (same problem description as the 3xTF32 GEMM)

namespace th {

// A pre-split operand: `rows` rows of `pitch` halves.  Row r holds, per 32-wide block j of the operand's K axis,
// 32 fp16 "hi" values at [64 j, 64 j + 32) and 32 fp16 "lo" values at [64 j + 32, 64 j + 64), both of x * 2^e(r);
// scale[r] = 2^-e(r) undoes the row's power-of-two scaling in the epilogue.  The block used starts at (row0, k0).
struct Split16 { const __half *data; const float *scale; long pitch; long rows; int row0; int k0; };

// fp32 (rows x K, row stride ld, `batch` matrices batch_stride apart) -> Split16 rows (rows_pad >= rows per batch; the padding
// rows and the k's in [K, Kp) are written as zeros).  dst: batch * rows_pad rows of 2 * Kp halves; scale: batch * rows_pad floats.
int split_rows_f16(const float *src, long ld, long batch_stride, int rows, int rows_pad, int K, int Kp, int batch, __half *dst,
                   float *scale, cudaStream_t st);
// The transpose of an upper triangular (n x n) row-major U as a Split16 operand: row j of dst = column j of U (K axis = U's
// row index, pitch 2 n halves).  Only the part the rank-k updates read (k <= j, in whole 32 x 32 tiles) is written.
// cmax: n uint32 of scratch.
int transpose_split_upper_f16(const float *U, int n, __half *dst, float *scale, unsigned int *cmax, cudaStream_t st);

// C[m][n] = alpha * sA[m] * sB[n] * sum_k (Ah Bh + Ah Bl + Al Bh)[m][n] + beta * C[m][n]
// M: multiple of 128 (rows of the A block); only rows < m_valid of C are touched.  N: multiple of 128; K, k0: multiples of 32.
int gemm_f16x3_nt_presplit(const Split16 &A, const Split16 &B, float *C, long ldc, int M, int m_valid, int N, int K, float alpha,
                           float beta, cudaStream_t st);

// Same problem description as tg::gemm_tf32x3_nt (operands split here, into `ws`).
size_t workspace_bytes(int M, int N, int K, int batch, bool same_ab);
int gemm_f16x3_nt(const tg::GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace th
