// tile.cuh -- CTA-level building blocks shared by the GPTQ layer kernel and the RTN kernel:
// a (R x 256) fp32 super-block tile in shared memory, the per-group scale search over it,
// and the per-row double quantisation of the group scales.
#pragma once
#include "kquant.cuh"

// Shared-memory layout of a (rows x 256) fp32 tile: row stride 256 floats, 16-byte units XOR-swizzled
// inside each 128-byte span by the span index, so that (a) one thread per group reading its group with
// LDS.128 and (b) a warp writing 32 consecutive float4 are both bank-conflict free.
__device__ __forceinline__ int wt_idx4(int row, int c4) { return row * 256 + ((c4 ^ ((c4 >> 3) & 7)) << 2); }
__device__ __forceinline__ int wt_idx(int row, int col) { return wt_idx4(row, col >> 2) + (col & 3); }

// Per-row K-quant metadata of the current super-block, kept in shared memory.
template <int R> struct RowScales {
    float d[R];            // fp32(fp16 super scale)
    float dm[R];           // fp32(fp16 super min)
    uint16_t dbits[R];
    uint16_t dmbits[R];
    uint8_t sq[R][16];
    uint8_t zq[R][16];
};

// One thread per (row, group): search the group's scale / zero.  gsc/gzr: [R][16] scratch.
template <int QT, int R, int NT>
__device__ __forceinline__ void tile_search(const float *Wt, float *gsc, float *gzr, const SearchParams &sp,
                                            uint32_t &vmask, uint32_t &amask) {
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS, V = GS / 4;
    for (int task = threadIdx.x; task < R * GPR; task += NT) {
        const int row = task / GPR, g = task % GPR;
        float x[GS];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 t = *reinterpret_cast<const float4 *>(Wt + wt_idx4(row, g * V + v));
            x[4 * v + 0] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
        }
        float s, z;
        kq_group_search<QT>(x, sp, s, z, vmask, amask);
        gsc[row * 16 + g] = s;
        gzr[row * 16 + g] = z;
    }
}

// One thread per row: super-block double quantisation; fills RowScales.
template <int QT, int R>
__device__ __forceinline__ void tile_finalize_row(int row, const float *gsc, const float *gzr, RowScales<R> &rs) {
    uint16_t db, dmb;
    kq_row_finalize<QT>(gsc + row * 16, gzr + row * 16, db, dmb, rs.sq[row], rs.zq[row]);
    rs.dbits[row] = db;
    rs.dmbits[row] = dmb;
    rs.d[row] = __half2float(__ushort_as_half(db));
    rs.dm[row] = __half2float(__ushort_as_half(dmb));
}

// OR-reduce the per-thread search masks over the CTA and publish them (2 u32 per super-block).
__device__ __forceinline__ void publish_flags(uint32_t *flags, uint32_t vmask, uint32_t amask) {
    if (flags == nullptr) return;
    vmask = __reduce_or_sync(0xffffffffu, vmask);
    amask = __reduce_or_sync(0xffffffffu, amask);
    if ((threadIdx.x & 31) == 0) {
        if (vmask) atomicOr(flags, vmask);
        if (amask) atomicOr(flags + 1, amask);
    }
}

// Outputs of one finished (R x 256) super-block tile: codes, GGUF block bytes, dequantised weights.
// Wt holds the dequantised values, codes the (R x 256) code bytes, rs the row metadata.
template <int QT, int R, int NT>
__device__ __forceinline__ void tile_emit(const float *Wt, const uint8_t *codes, const RowScales<R> &rs, int r0,
                                          int d_row, size_t ld, int c, int sb, int nsb, uint8_t *qweight,
                                          uint8_t *packed, void *wdeq, int wdeq_dtype) {
    constexpr int TS = Fmt<QT>::TS;
    const int tid = threadIdx.x;
    if (qweight != nullptr) {
        for (int id = tid; id < R * 16; id += NT) {
            const int row = id >> 4, c16 = id & 15;
            if (r0 + row < d_row)
                *reinterpret_cast<uint4 *>(qweight + (size_t)(r0 + row) * ld + c + 16 * c16) =
                    *reinterpret_cast<const uint4 *>(codes + row * 256 + 16 * c16);
        }
    }
    if (packed != nullptr) {
        for (int id = tid; id < R * 8; id += NT) {
            const int row = id >> 3, q8 = id & 7;
            if (r0 + row < d_row) {
                uint8_t *o = packed + ((size_t)(r0 + row) * nsb + sb) * TS;
                for (int b = q8; b < TS; b += 8)
                    o[b] = kq_pack_byte<QT>(b, codes + row * 256, rs.sq[row], rs.zq[row], rs.dbits[row], rs.dmbits[row]);
            }
        }
    }
    if (wdeq != nullptr) {
        for (int id = tid; id < R * 64; id += NT) {
            const int row = id >> 6, c4 = id & 63;
            if (r0 + row < d_row) {
                const float4 v = *reinterpret_cast<const float4 *>(Wt + wt_idx4(row, c4));
                const size_t o = (size_t)(r0 + row) * ld + c + 4 * c4;
                if (wdeq_dtype == GQ_F32) {
                    *reinterpret_cast<float4 *>((float *)wdeq + o) = v;
                } else if (wdeq_dtype == GQ_BF16) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
                    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)wdeq + o) =
                        make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
                } else {
                    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
                    *reinterpret_cast<uint2 *>((__half *)wdeq + o) =
                        make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
                }
            }
        }
    }
}
