// gptq_blocksize.cu -- GPTQ.step (quant/gptq/src/gptq.py:146-295 of the reference) for --block_size 32, 64 and 256.
//
// The reference takes any block size (gptq.py:55, CLI quant.py --block_size); run_quant.sh uses 128, which is what the fused
// kernels of gptq_layer.cu are built around (two 128-column blocks per 256-column super-block).  The other sizes that divide a
// super-block, or equal it, run here on a plain right-looking schedule with the SAME arithmetic, bit for bit
// (tests/golden/blocksize_a.npz holds the reference's own outputs for them):
//   per block of B columns  [c1, c2):
//     * at a super-block boundary (c1 % 256 == 0): scale / min search on W[:, c1:c1+256] as it is now (all earlier blocks
//       applied) -- gq_get_scale_and_zero's kernel, gptq.py:240-245;
//     * block_serial_kernel: one thread per row walks the B columns: quantise, dequantise, err = (w - w_q) / U[i,i], then
//       w[j] -= fl(err * U[i,j]) for the later columns OF THE BLOCK (two roundings, like addr_ on the CPU), gptq.py:247-268;
//       the errors replace the consumed columns of W;
//     * block_trailing_kernel: W[:, c2:] -= E U[c1:c2, c2:], per element ONE fresh fp32 FMA chain over the block's B k's in
//       ascending order and one subtraction (== addmm_ on the CPU, gptq.py:270).
//   GGUF bytes and dequantised weights come from the stand-alone kernels (gq_pack, gq_dequantize) at the end.
// Correctness path, not a tuned one: a 4096 x 4096 layer takes a few milliseconds, block size 32 launches 3 kernels per 32 columns.
#include "kquant.cuh"

namespace {

constexpr int RS = 32;       // rows per CTA of the serial kernel (one warp, one thread per row)

template <int QT>
__global__ void __launch_bounds__(RS) block_serial_kernel(float *W, const float *__restrict__ U, int d_row, int d_col, int c1, int B,
                                                          const uint16_t *__restrict__ d, const uint16_t *__restrict__ dmin,
                                                          const uint8_t *__restrict__ sq, const uint8_t *__restrict__ zq,
                                                          uint8_t *qweight) {
    extern __shared__ float tile[];      // [RS][B + 1]
    constexpr int GS = Fmt<QT>::GS;
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    const int t = threadIdx.x, r0 = blockIdx.x * RS, ts = B + 1;
    const size_t ld = (size_t)d_col;
    const int nsb = d_col / GQ_QK_K, ng = d_col / GS;
    for (int r = 0; r < RS; ++r) {
        const size_t gr = (size_t)min(r0 + r, d_row - 1);
        for (int j = t; j < B; j += RS) tile[r * ts + j] = W[gr * ld + c1 + j];
    }
    __syncwarp();
    const int row = r0 + t;
    if (row < d_row) {
        float *w = tile + t * ts;
        float sc = 0.0f, zz = 0.0f;
        for (int i = 0; i < B; ++i) {
            const int col = c1 + i;
            if (i == 0 || col % GS == 0) {
                sc = __fmul_rn(__half2float(__ushort_as_half(d[(size_t)row * nsb + col / GQ_QK_K])),
                               kq_code_to_f<QT>(sq[(size_t)row * ng + col / GS]));
                zz = __fmul_rn(__half2float(__ushort_as_half(dmin[(size_t)row * nsb + col / GQ_QK_K])),
                               kq_code_to_f<QT>(zq[(size_t)row * ng + col / GS]));
            }
            const float x = w[i];
            const float qv = kq_quant(x, sc, zz, lo, hi);                                // gptq.py:247-254
            const float wq = kq_dequant(qv, sc, zz);                                     // :255-261
            const float *urow = U + (size_t)col * ld + c1;
            const float err = __fdiv_rn(__fsub_rn(x, wq), urow[i]);                      // :264
            qweight[(size_t)row * ld + col] = (uint8_t)(int8_t)(int)qv;                  // :263
            w[i] = err;                                                                  // :268 (E replaces the column)
            for (int j = i + 1; j < B; ++j) w[j] = __fsub_rn(w[j], __fmul_rn(err, urow[j]));   // :267
        }
    }
    __syncwarp();
    for (int r = 0; r < RS; ++r) {
        if (r0 + r >= d_row) break;
        for (int j = t; j < B; j += RS) W[(size_t)(r0 + r) * ld + c1 + j] = tile[r * ts + j];
    }
}

// W[r, col] <- W[r, col] - chain_k(E[r, k] U[c1 + k, col]),  col >= c2; one thread per column, 32 rows per CTA.
__global__ void __launch_bounds__(256) block_trailing_kernel(float *W, const float *__restrict__ U, int d_row, int d_col, int c1, int B) {
    extern __shared__ float E[];         // [32][B]
    const int r0 = blockIdx.y * 32, c2 = c1 + B;
    const size_t ld = (size_t)d_col;
    for (int id = threadIdx.x; id < 32 * B; id += 256) {
        const int r = id / B, k = id - r * B;
        E[id] = W[(size_t)min(r0 + r, d_row - 1) * ld + c1 + k];
    }
    __syncthreads();
    const int col = c2 + blockIdx.x * 256 + threadIdx.x;
    if (col >= d_col) return;
    float acc[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) acc[r] = 0.0f;
    for (int k = 0; k < B; ++k) {
        const float u = U[(size_t)(c1 + k) * ld + col];
#pragma unroll
        for (int r = 0; r < 32; ++r) acc[r] = __fmaf_rn(E[r * B + k], u, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r)
        if (r0 + r < d_row) {
            float *p = W + (size_t)(r0 + r) * ld + col;
            *p = __fsub_rn(*p, acc[r]);
        }
}

template <int QT>
int run_blocks(float *W, const float *U, int d_row, int d_col, int B, double rmin, double rdelta, int nstep, bool searched,
               uint8_t *qweight, uint16_t *d, uint8_t *sq, uint16_t *dmin, uint8_t *zq, uint32_t *flags, cudaStream_t st) {
    constexpr int GPR = GQ_QK_K / Fmt<QT>::GS;
    const int nsb = d_col / GQ_QK_K, ng = d_col / Fmt<QT>::GS;
    const size_t serial_smem = (size_t)RS * (B + 1) * sizeof(float), trail_smem = (size_t)32 * B * sizeof(float);
    GQ_CHECK_CUDA(cudaFuncSetAttribute(block_serial_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)serial_smem));
    GQ_CHECK_CUDA(cudaFuncSetAttribute(block_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trail_smem));
    for (int c1 = 0; c1 < d_col; c1 += B) {
        if (!searched && c1 % GQ_QK_K == 0) {
            const int sb = c1 / GQ_QK_K;
            const int rc = gq_get_scale_and_zero(W + c1, d_col, d_row, QT, rmin, rdelta, nstep, d + sb, dmin + sb, nsb, sq + sb * GPR,
                                                 zq + sb * GPR, ng, flags ? flags + 2 * sb : nullptr, (gq_stream_t)st);
            if (rc) return rc;
        }
        block_serial_kernel<QT><<<(d_row + RS - 1) / RS, RS, serial_smem, st>>>(W, U, d_row, d_col, c1, B, d, dmin, sq, zq, qweight);
        gq_count_launches(1);
        const int rest = d_col - c1 - B;
        if (rest > 0) {
            block_trailing_kernel<<<dim3((rest + 255) / 256, (d_row + 31) / 32), 256, trail_smem, st>>>(W, U, d_row, d_col, c1, B);
            gq_count_launches(1);
        }
    }
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

}  // namespace

// Internal (gq_gptq_quantize_ex dispatches here for block_size 32 / 64 / 256).  `searched`: d / dmin / sq / zq already hold the
// scales of every super-block (static_groups).
int gq_gptq_blocksize(float *W, const float *U, int d_row, int d_col, int qtype, int B, double rmin, double rdelta, int nstep,
                      bool searched, void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq, uint8_t *packed, void *wdeq,
                      int wdeq_dtype, uint32_t *flags, cudaStream_t st) {
    uint8_t *q8 = (uint8_t *)qweight, *s8 = (uint8_t *)sq, *z8 = (uint8_t *)zq;
    int rc;
    switch (qtype) {
    case GQ_Q2_K: rc = run_blocks<GQ_Q2_K>(W, U, d_row, d_col, B, rmin, rdelta, nstep, searched, q8, d, s8, dmin, z8, flags, st); break;
    case GQ_Q3_K: rc = run_blocks<GQ_Q3_K>(W, U, d_row, d_col, B, rmin, rdelta, nstep, searched, q8, d, s8, dmin, z8, flags, st); break;
    case GQ_Q4_K: rc = run_blocks<GQ_Q4_K>(W, U, d_row, d_col, B, rmin, rdelta, nstep, searched, q8, d, s8, dmin, z8, flags, st); break;
    case GQ_Q5_K: rc = run_blocks<GQ_Q5_K>(W, U, d_row, d_col, B, rmin, rdelta, nstep, searched, q8, d, s8, dmin, z8, flags, st); break;
    default: rc = run_blocks<GQ_Q6_K>(W, U, d_row, d_col, B, rmin, rdelta, nstep, searched, q8, d, s8, dmin, z8, flags, st); break;
    }
    if (rc) return rc;
    if (packed != nullptr) {
        rc = gq_pack(qtype, qweight, d, sq, dmin, zq, d_row, d_col, packed, (gq_stream_t)st);
        if (rc) return rc;
        gq_count_launches(1);
    }
    if (wdeq != nullptr) {
        rc = gq_dequantize(qtype, qweight, d, sq, dmin, zq, d_row, d_col, wdeq, wdeq_dtype, (gq_stream_t)st);
        if (rc) return rc;
        gq_count_launches(1);
    }
    return GQ_OK;
}
