// rtn_native.cuh -- bodies of the RTN kernels of csrc/rtn.cu -- the shipped rtn_kernel (fp32 arithmetic) and its
// native-arithmetic twin (gq_rtn_quantize_native) -- in a header of their own so that the CPU suite can run them on the SIMT
// emulator (tests/test_simt_emu_cpu.py).
// Params / Smem are rtn.cu's RtnParams / RtnSmem (template parameters here only because those live in rtn.cu).
#pragma once
#include "tile.cuh"
#include "kquant_bf16.cuh"

// Twin of rtn_kernel for BF16 / FP16 weights with the reference's scale search in that dtype's arithmetic (kquant_bf16.cuh);
// everything after the search -- quantize() in fp32, codes, GGUF bytes, dequantised weights -- is the same code.
// Reached only through gq_rtn_quantize_native; gq_rtn_quantize (fp32 search on widened weights) is untouched.
template <int QT, int RND, int R, int NT, class Params, class Smem>
__device__ __forceinline__ void rtn_native_body(const Params &p, Smem &sm) {
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS, V = GS / 4;
    constexpr int MAXQ = (1 << Fmt<QT>::BITS) - 1;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * R, sb = blockIdx.y, c = sb * GQ_QK_K;
    for (int id = tid; id < R * 64; id += NT) {
        const int row = id >> 6, c4 = id & 63;
        const long base = (long)min(r0 + row, p.d_row - 1) * p.ld_in + c + 4 * c4;
        float4 v;
        v.x = load_as_f32(p.W, base + 0, RND); v.y = load_as_f32(p.W, base + 1, RND);      // RND == the gq_dtype code
        v.z = load_as_f32(p.W, base + 2, RND); v.w = load_as_f32(p.W, base + 3, RND);
        *reinterpret_cast<float4 *>(sm.Wt + wt_idx4(row, c4)) = v;
    }
    __syncthreads();
    for (int task = tid; task < R * GPR; task += NT) {
        const int row = task / GPR, g = task % GPR;
        float x[GS];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 t = *reinterpret_cast<const float4 *>(sm.Wt + wt_idx4(row, g * V + v));
            x[4 * v + 0] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
        }
        float s, z;
        if constexpr (Fmt<QT>::ASYM) kqb_search_asym<GS, MAXQ, RND>(x, p.sp, s, z);
        else kqb_search_sym<GS, MAXQ, RND>(x, s, z);
        sm.gsc[row * 16 + g] = s;
        sm.gzr[row * 16 + g] = z;
    }
    __syncthreads();
    if (tid < R) {
        uint16_t db, dmb;
        kqb_row_finalize<QT, RND>(sm.gsc + tid * 16, sm.gzr + tid * 16, db, dmb, sm.rs.sq[tid], sm.rs.zq[tid]);
        sm.rs.dbits[tid] = db;
        sm.rs.dmbits[tid] = dmb;
        sm.rs.d[tid] = __half2float(__ushort_as_half(db));
        sm.rs.dm[tid] = __half2float(__ushort_as_half(dmb));
        if (r0 + tid < p.d_row) {
            const long gr = r0 + tid;
            p.d[gr * p.d_stride + sb] = db;
            p.dmin[gr * p.d_stride + sb] = dmb;
#pragma unroll
            for (int g = 0; g < GPR; ++g) {
                p.sq[gr * p.sq_stride + sb * GPR + g] = sm.rs.sq[tid][g];
                p.zq[gr * p.sq_stride + sb * GPR + g] = sm.rs.zq[tid][g];
            }
        }
    }
    __syncthreads();
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    for (int id = tid; id < R * 256; id += NT) {          // quantize() in fp32: quant_utils.py:34-40 promotes bf16 + fp32
        const int row = id >> 8, col = id & 255, g = col / GS;
        const float s = __fmul_rn(sm.rs.d[row], kq_code_to_f<QT>(sm.rs.sq[row][g]));
        const float z = __fmul_rn(sm.rs.dm[row], kq_code_to_f<QT>(sm.rs.zq[row][g]));
        const int wi = wt_idx(row, col);
        const float q = kq_quant(sm.Wt[wi], s, z, lo, hi);
        sm.codes[row * 256 + col] = (uint8_t)(int8_t)(int)q;
        sm.Wt[wi] = kq_dequant(q, s, z);
    }
    __syncthreads();
    tile_emit<QT, R, NT>(sm.Wt, sm.codes, sm.rs, r0, p.d_row, (size_t)p.nsb * GQ_QK_K, c, sb, p.nsb, p.qweight,
                         p.packed, p.wdeq, p.wdeq_dtype);
}

// Body of rtn_kernel (csrc/rtn.cu): the shipped RTN K-quant of one (R rows x 256 columns) tile, fp32 arithmetic.
template <int QT, int R, int NT, class Params, class Smem>
__device__ __forceinline__ void rtn_body(const Params &p, Smem &sm) {
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * R, sb = blockIdx.y, c = sb * GQ_QK_K;

    for (int id = tid; id < R * 64; id += NT) {
        const int row = id >> 6, c4 = id & 63;
        const long base = (long)min(r0 + row, p.d_row - 1) * p.ld_in + c + 4 * c4;
        float4 v;
        if (p.w_dtype == GQ_F32) {
            v = *reinterpret_cast<const float4 *>((const float *)p.W + base);
        } else {
            v.x = load_as_f32(p.W, base + 0, p.w_dtype); v.y = load_as_f32(p.W, base + 1, p.w_dtype);
            v.z = load_as_f32(p.W, base + 2, p.w_dtype); v.w = load_as_f32(p.W, base + 3, p.w_dtype);
        }
        *reinterpret_cast<float4 *>(sm.Wt + wt_idx4(row, c4)) = v;
    }
    __syncthreads();
    uint32_t vmask = 0, amask = 0;
    tile_search<QT, R, NT>(sm.Wt, sm.gsc, sm.gzr, p.sp, vmask, amask);
    publish_flags(p.flags ? p.flags + 2 * sb : nullptr, vmask, amask);
    __syncthreads();
    if (tid < R) {
        tile_finalize_row<QT, R>(tid, sm.gsc, sm.gzr, sm.rs);
        if (r0 + tid < p.d_row) {
            const long gr = r0 + tid;
            p.d[gr * p.d_stride + sb] = sm.rs.dbits[tid];
            p.dmin[gr * p.d_stride + sb] = sm.rs.dmbits[tid];
#pragma unroll
            for (int g = 0; g < GPR; ++g) {
                p.sq[gr * p.sq_stride + sb * GPR + g] = sm.rs.sq[tid][g];
                p.zq[gr * p.sq_stride + sb * GPR + g] = sm.rs.zq[tid][g];
            }
        }
    }
    if (p.qweight == nullptr && p.packed == nullptr && p.wdeq == nullptr) return;
    __syncthreads();
    // quantize every weight of the tile (quantizer.py:323 -> quant_utils.py:34-40)
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    for (int id = tid; id < R * 256; id += NT) {
        const int row = id >> 8, col = id & 255, g = col / GS;
        const float s = __fmul_rn(sm.rs.d[row], kq_code_to_f<QT>(sm.rs.sq[row][g]));
        const float z = __fmul_rn(sm.rs.dm[row], kq_code_to_f<QT>(sm.rs.zq[row][g]));
        const int wi = wt_idx(row, col);
        const float q = kq_quant(sm.Wt[wi], s, z, lo, hi);
        sm.codes[row * 256 + col] = (uint8_t)(int8_t)(int)q;
        sm.Wt[wi] = kq_dequant(q, s, z);
    }
    __syncthreads();
    tile_emit<QT, R, NT>(sm.Wt, sm.codes, sm.rs, r0, p.d_row, (size_t)p.nsb * GQ_QK_K, c, sb, p.nsb, p.qweight,
                         p.packed, p.wdeq, p.wdeq_dtype);
}

