// rtn.cu -- Hessian-free K-quant kernels:
//   gq_rtn_quantize        (replaces Quantizer._quant_non_block_module, quant/gptq/src/quantizer.py:278-330)
//   gq_get_scale_and_zero  (replaces quant_utils.Quantizer.get_scale_and_zero, quant_utils.py:90-145)
//   gq_dequantize          (replaces dequantize_linear_weight, quant_utils.py:277-310)
//   gq_pack                (replaces pack_Q2K..pack_Q6K, packing_utils.py:33-326)
// One CTA handles a (32 rows x 256 columns) super-block tile; the grid covers (row tiles, super-blocks), so
// embed_tokens / lm_head (128256 x 4096) launch 4008 x 16 independent CTAs -- purely HBM/ALU work, no GEMM.
#include "rtn_native.cuh"
#include "tile.cuh"

namespace {

constexpr int R = 32;
constexpr int NT = 256;

#include "rtn_structs.cuh"      // RtnParams, RtnSmem (shared with the CPU suite's emulator harness)

template <int QT>
__global__ void __launch_bounds__(NT) rtn_kernel(const RtnParams p) {
    __shared__ RtnSmem sm;
    rtn_body<QT, R, NT, RtnParams, RtnSmem>(p, sm);
}

template <int QT> int launch_rtn(const RtnParams &p, cudaStream_t st) {
    dim3 grid((p.d_row + R - 1) / R, p.nsb);
    rtn_kernel<QT><<<grid, NT, 0, st>>>(p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

// Twin of rtn_kernel for BF16 / FP16 weights with the reference's scale search in that dtype's arithmetic; the
// body lives in rtn_native.cuh.  Reached only through gq_rtn_quantize_native; gq_rtn_quantize is untouched.
template <int QT, int RND>
__global__ void __launch_bounds__(NT) rtn_bf16_kernel(const RtnParams p) {
    __shared__ RtnSmem sm;
    rtn_native_body<QT, RND, R, NT, RtnParams, RtnSmem>(p, sm);
}

template <int QT> int launch_rtn_bf16(const RtnParams &p, cudaStream_t st) {
    dim3 grid((p.d_row + R - 1) / R, p.nsb);
    static_assert(GQ_RND_BF16 == GQ_BF16 && GQ_RND_F16 == GQ_F16, "rounding codes are the gq_dtype codes");
    if (p.w_dtype == GQ_BF16) rtn_bf16_kernel<QT, GQ_RND_BF16><<<grid, NT, 0, st>>>(p);
    else rtn_bf16_kernel<QT, GQ_RND_F16><<<grid, NT, 0, st>>>(p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

int dispatch_rtn(int qtype, const RtnParams &p, cudaStream_t st) {
    switch (qtype) {
    case GQ_Q2_K: return launch_rtn<GQ_Q2_K>(p, st);
    case GQ_Q3_K: return launch_rtn<GQ_Q3_K>(p, st);
    case GQ_Q4_K: return launch_rtn<GQ_Q4_K>(p, st);
    case GQ_Q5_K: return launch_rtn<GQ_Q5_K>(p, st);
    default: return launch_rtn<GQ_Q6_K>(p, st);
    }
}

// ---- dequantize_linear_weight ----------------------------------------------------------------
template <int QT>
__global__ void dequant_kernel(const uint8_t *__restrict__ qw, const uint16_t *__restrict__ d,
                               const uint8_t *__restrict__ sq, const uint16_t *__restrict__ dmin,
                               const uint8_t *__restrict__ zq, int d_row, int d_col, void *out, int out_dtype) {
    constexpr int GS = Fmt<QT>::GS;
    const long n4 = (long)d_row * d_col / 4;
    const int nsb = d_col / GQ_QK_K, ng = d_col / GS;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const long e = i * 4;
        const long row = e / d_col;
        const int col = (int)(e % d_col);
        const float s = __fmul_rn(__half2float(__ushort_as_half(d[row * nsb + col / GQ_QK_K])),
                                  kq_code_to_f<QT>(sq[row * ng + col / GS]));
        const float z = __fmul_rn(__half2float(__ushort_as_half(dmin[row * nsb + col / GQ_QK_K])),
                                  kq_code_to_f<QT>(zq[row * ng + col / GS]));
        const uchar4 c4 = *reinterpret_cast<const uchar4 *>(qw + e);
        const float v0 = kq_dequant(kq_code_to_f<QT>(c4.x), s, z), v1 = kq_dequant(kq_code_to_f<QT>(c4.y), s, z);
        const float v2 = kq_dequant(kq_code_to_f<QT>(c4.z), s, z), v3 = kq_dequant(kq_code_to_f<QT>(c4.w), s, z);
        store_from_f32(out, e + 0, out_dtype, v0); store_from_f32(out, e + 1, out_dtype, v1);
        store_from_f32(out, e + 2, out_dtype, v2); store_from_f32(out, e + 3, out_dtype, v3);
    }
}

// ---- pack_Q*K --------------------------------------------------------------------------------
template <int QT>
__global__ void pack_kernel(const uint8_t *__restrict__ qw, const uint16_t *__restrict__ d,
                            const uint8_t *__restrict__ sq, const uint16_t *__restrict__ dmin,
                            const uint8_t *__restrict__ zq, long nblk, uint8_t *out) {
    constexpr int TS = Fmt<QT>::TS, GPR = GQ_QK_K / Fmt<QT>::GS;
    const long total = nblk * TS;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long blk = i / TS;
        const int b = (int)(i % TS);
        // Q3_K / Q6_K blocks carry no min: dmin / zq may be NULL for them (pack_Q3K / pack_Q6K take 3 tensors)
        out[i] = kq_pack_byte<QT>(b, qw + blk * GQ_QK_K, sq + blk * GPR, zq ? zq + blk * GPR : nullptr, d[blk],
                                  dmin ? dmin[blk] : (uint16_t)0);
    }
}

int grid_for(long n, int bs) {
    long g = (n + bs - 1) / bs;
    const long cap = 148L * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

void gq_fill_search_params(SearchParams &sp, int maxq, double rmin, double rdelta, int nstep);

extern "C" int gq_rtn_quantize(const void *W, int w_dtype, int d_row, int d_col, int qtype, double rmin,
                               double rdelta, int nstep, void *qweight, uint16_t *d, void *sq, uint16_t *dmin,
                               void *zq, uint8_t *packed, void *wdeq, int wdeq_dtype, gq_stream_t stream) {
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "gq_rtn_quantize: unknown q_type %d", qtype);
    GQ_REQUIRE(W && qweight && d && sq && dmin && zq, "gq_rtn_quantize: null pointer");
    GQ_REQUIRE(d_row > 0 && d_col > 0 && d_col % GQ_QK_K == 0, "gq_rtn_quantize: d_col=%d must be a positive multiple of 256", d_col);
    GQ_REQUIRE(w_dtype >= GQ_F32 && w_dtype <= GQ_BF16, "gq_rtn_quantize: bad w_dtype %d", w_dtype);
    GQ_REQUIRE(nstep >= 0 && nstep < 64, "gq_rtn_quantize: nstep=%d out of range [0,63]", nstep);
    GQ_REQUIRE(((uintptr_t)W | (uintptr_t)qweight | (uintptr_t)wdeq) % 16 == 0, "gq_rtn_quantize: W, qweight, wdeq must be 16-byte aligned");
    RtnParams p;
    p.W = W; p.w_dtype = w_dtype; p.ld_in = d_col; p.d_row = d_row; p.nsb = d_col / GQ_QK_K;
    gq_fill_search_params(p.sp, (1 << f.bits) - 1, rmin, rdelta, nstep);
    p.d = d; p.dmin = dmin; p.d_stride = p.nsb; p.sq = (uint8_t *)sq; p.zq = (uint8_t *)zq; p.sq_stride = d_col / f.gs;
    p.qweight = (uint8_t *)qweight; p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = wdeq_dtype; p.flags = nullptr;
    return dispatch_rtn(qtype, p, (cudaStream_t)stream);
}

// (see kquant_bf16.cuh) gq_rtn_quantize for a BF16 / FP16 weight with the scale search in the weight's own
// arithmetic -- what quantizer.py:278-330 computes for embed_tokens / lm_head of a 16-bit model.  Same outputs and conventions.
extern "C" int gq_rtn_quantize_native(const void *W, int w_dtype, int d_row, int d_col, int qtype, double rmin,
                                      double rdelta, int nstep, void *qweight, uint16_t *d, void *sq, uint16_t *dmin,
                                      void *zq, uint8_t *packed, void *wdeq, int wdeq_dtype, gq_stream_t stream) {
    if (w_dtype == GQ_F32)       // fp32 weights: the native arithmetic IS the fp32 search
        return gq_rtn_quantize(W, w_dtype, d_row, d_col, qtype, rmin, rdelta, nstep, qweight, d, sq, dmin, zq, packed, wdeq,
                               wdeq_dtype, stream);
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "gq_rtn_quantize_native: unknown q_type %d", qtype);
    GQ_REQUIRE(W && qweight && d && sq && dmin && zq, "gq_rtn_quantize_native: null pointer");
    GQ_REQUIRE(d_row > 0 && d_col > 0 && d_col % GQ_QK_K == 0, "gq_rtn_quantize_native: d_col=%d must be a positive multiple of 256", d_col);
    GQ_REQUIRE(nstep >= 0 && nstep < 64, "gq_rtn_quantize_native: nstep=%d out of range [0,63]", nstep);
    GQ_REQUIRE(((uintptr_t)W | (uintptr_t)qweight | (uintptr_t)wdeq) % 16 == 0, "gq_rtn_quantize_native: W, qweight, wdeq must be 16-byte aligned");
    RtnParams p;
    p.W = W; p.w_dtype = w_dtype; p.ld_in = d_col; p.d_row = d_row; p.nsb = d_col / GQ_QK_K;
    gq_fill_search_params(p.sp, (1 << f.bits) - 1, rmin, rdelta, nstep);
    p.d = d; p.dmin = dmin; p.d_stride = p.nsb; p.sq = (uint8_t *)sq; p.zq = (uint8_t *)zq; p.sq_stride = d_col / f.gs;
    p.qweight = (uint8_t *)qweight; p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = wdeq_dtype; p.flags = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    switch (qtype) {
    case GQ_Q2_K: return launch_rtn_bf16<GQ_Q2_K>(p, st);
    case GQ_Q3_K: return launch_rtn_bf16<GQ_Q3_K>(p, st);
    case GQ_Q4_K: return launch_rtn_bf16<GQ_Q4_K>(p, st);
    case GQ_Q5_K: return launch_rtn_bf16<GQ_Q5_K>(p, st);
    default: return launch_rtn_bf16<GQ_Q6_K>(p, st);
    }
}

extern "C" int gq_get_scale_and_zero(const float *x, long x_stride, int rows, int qtype, double rmin, double rdelta,
                                     int nstep, uint16_t *d, uint16_t *dmin, long d_stride, void *sq, void *zq,
                                     long sq_stride, uint32_t *search_flags, gq_stream_t stream) {
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "gq_get_scale_and_zero: unknown q_type %d", qtype);
    GQ_REQUIRE(x && d && dmin && sq && zq, "gq_get_scale_and_zero: null pointer");
    GQ_REQUIRE(rows > 0 && x_stride >= GQ_QK_K && x_stride % 4 == 0 && (uintptr_t)x % 16 == 0,
               "gq_get_scale_and_zero: x must be 16-byte aligned with a row stride that is a multiple of 4");
    GQ_REQUIRE(nstep >= 0 && nstep < 64, "gq_get_scale_and_zero: nstep=%d out of range [0,63]", nstep);
    RtnParams p;
    p.W = x; p.w_dtype = GQ_F32; p.ld_in = x_stride; p.d_row = rows; p.nsb = 1;
    gq_fill_search_params(p.sp, (1 << f.bits) - 1, rmin, rdelta, nstep);
    p.d = d; p.dmin = dmin; p.d_stride = d_stride; p.sq = (uint8_t *)sq; p.zq = (uint8_t *)zq; p.sq_stride = sq_stride;
    p.qweight = nullptr; p.packed = nullptr; p.wdeq = nullptr; p.wdeq_dtype = GQ_F32; p.flags = search_flags;
    return dispatch_rtn(qtype, p, (cudaStream_t)stream);
}

// Scales / zeros of EVERY super-block of a (d_row, d_col) fp32 matrix in one launch -- GPTQ.step's static_groups
// initialisation (gptq.py:184-196).  Internal (used by gq_gptq_quantize_ex); same outputs as d_col/256 calls of
// gq_get_scale_and_zero on the 256-column slabs.
int gq_search_all_superblocks(const float *W, int d_row, int d_col, int qtype, double rmin, double rdelta, int nstep,
                              uint16_t *d, uint16_t *dmin, void *sq, void *zq, uint32_t *search_flags, cudaStream_t st) {
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "static_groups: unknown q_type %d", qtype);
    RtnParams p;
    p.W = W; p.w_dtype = GQ_F32; p.ld_in = d_col; p.d_row = d_row; p.nsb = d_col / GQ_QK_K;
    gq_fill_search_params(p.sp, (1 << f.bits) - 1, rmin, rdelta, nstep);
    p.d = d; p.dmin = dmin; p.d_stride = p.nsb; p.sq = (uint8_t *)sq; p.zq = (uint8_t *)zq; p.sq_stride = d_col / f.gs;
    p.qweight = nullptr; p.packed = nullptr; p.wdeq = nullptr; p.wdeq_dtype = GQ_F32; p.flags = search_flags;
    return dispatch_rtn(qtype, p, st);
}

extern "C" int gq_dequantize(int qtype, const void *qweight, const uint16_t *d, const void *sq, const uint16_t *dmin,
                             const void *zq, int d_row, int d_col, void *out, int out_dtype, gq_stream_t stream) {
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "gq_dequantize: unknown q_type %d", qtype);
    GQ_REQUIRE(qweight && d && sq && dmin && zq && out, "gq_dequantize: null pointer");
    GQ_REQUIRE(d_row > 0 && d_col > 0 && d_col % GQ_QK_K == 0, "gq_dequantize: d_col=%d must be a positive multiple of 256", d_col);
    GQ_REQUIRE(out_dtype >= GQ_F32 && out_dtype <= GQ_BF16, "gq_dequantize: bad out_dtype %d", out_dtype);
    GQ_REQUIRE((uintptr_t)qweight % 4 == 0, "gq_dequantize: qweight must be 4-byte aligned");
    const int g = grid_for((long)d_row * d_col / 4, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t *q = (const uint8_t *)qweight, *s = (const uint8_t *)sq, *z = (const uint8_t *)zq;
    switch (qtype) {
    case GQ_Q2_K: dequant_kernel<GQ_Q2_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, d_row, d_col, out, out_dtype); break;
    case GQ_Q3_K: dequant_kernel<GQ_Q3_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, d_row, d_col, out, out_dtype); break;
    case GQ_Q4_K: dequant_kernel<GQ_Q4_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, d_row, d_col, out, out_dtype); break;
    case GQ_Q5_K: dequant_kernel<GQ_Q5_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, d_row, d_col, out, out_dtype); break;
    default: dequant_kernel<GQ_Q6_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, d_row, d_col, out, out_dtype); break;
    }
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" int gq_pack(int qtype, const void *qweight, const uint16_t *d, const void *sq, const uint16_t *dmin,
                       const void *zq, int d_row, int d_col, uint8_t *out, gq_stream_t stream) {
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "gq_pack: unknown q_type %d", qtype);
    GQ_REQUIRE(qweight && d && sq && out && (f.asym == 0 || (dmin && zq)), "gq_pack: null pointer");
    GQ_REQUIRE(d_row > 0 && d_col > 0 && d_col % GQ_QK_K == 0, "gq_pack: d_col=%d must be a positive multiple of 256", d_col);
    const long nblk = (long)d_row * (d_col / GQ_QK_K);
    const int g = grid_for(nblk * f.ts, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t *q = (const uint8_t *)qweight, *s = (const uint8_t *)sq, *z = (const uint8_t *)zq;
    switch (qtype) {
    case GQ_Q2_K: pack_kernel<GQ_Q2_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, nblk, out); break;
    case GQ_Q3_K: pack_kernel<GQ_Q3_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, nblk, out); break;
    case GQ_Q4_K: pack_kernel<GQ_Q4_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, nblk, out); break;
    case GQ_Q5_K: pack_kernel<GQ_Q5_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, nblk, out); break;
    default: pack_kernel<GQ_Q6_K><<<g, 256, 0, st>>>(q, d, s, dmin, z, nblk, out); break;
    }
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
