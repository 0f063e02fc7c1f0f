// gptq_layer.cu -- the column-blocked quantise -> error -> rank-k-update loop of one linear layer
// as ONE persistent kernel launch (replaces GPTQ.step, quant/gptq/src/gptq.py:146-295 of the reference).
//
// Design (B200-first, not the reference's right-looking Python loop):
//   * Rows of W are independent given U, so a CTA owns R=32 rows for the whole layer and walks the
//     d_col/256 super-blocks serially; there is no inter-CTA communication and no grid sync.
//   * The rank-k update is done LEFT-LOOKING: when a CTA reaches super-block c it applies, to its
//     (32 x 256) tile, the contributions of all earlier 128-column blocks b':  tile -= E[:,b'] * U[b', c:c+256].
//     W's trailing part is therefore read exactly once (no read-modify-write per block as in the
//     reference's addmm_), the propagated errors E are stored in place of the consumed columns of W, and
//     U streams through a 4-stage cp.async pipeline in (16 x 256) pieces.
//     Exact mode keeps the reference's arithmetic: per earlier block a fresh single-accumulator FMA chain
//     over its 128 k's in ascending order, then ONE subtraction from w -- bit-identical to addmm_ on CPU.
//   * The scale/min search, the 128 sequential column steps (rank-1 updates held in registers, 8 lanes
//     per row), the GGUF bit-pack and the dequantised write-back are fused in shared memory.
#include "tile.cuh"

namespace {

constexpr int R = 32;        // rows per CTA
constexpr int NT = 256;      // threads per CTA
constexpr int KP = 16;       // k's per pipeline piece
constexpr int S = 4;         // pipeline stages
constexpr int US_FLOATS = KP * 256;
constexpr int ES_FLOATS = R * KP;

struct LayerParams {
    float *W;
    const float *U;
    int d_row, d_col;
    SearchParams sp;
    uint8_t *qweight;
    uint16_t *d;
    uint8_t *sq;
    uint16_t *dmin;
    uint8_t *zq;
    uint8_t *packed;
    void *wdeq;
    int wdeq_dtype;
    uint32_t *flags;
};

struct __align__(16) Smem {
    float Wt[R * 256];                       // live super-block tile, later the dequantised values
    union {
        struct { float Us[S * US_FLOATS]; float Es[S * ES_FLOATS]; } pipe;
        float Ud[128 * 128];                 // diagonal block of U during the serial phase
    } u;
    float Et[R * 128];                       // errors of the current 128-column block
    uint8_t codes[R * 256];
    float gsc[R * 16];
    float gzr[R * 16];
    RowScales<R> rs;
};

// tile(8 rows x 4 cols per thread) -= E[:, kbeg:kend] * U[kbeg:kend, window]; the reference's addmm_ arithmetic.
// HALF: only the window's columns 128..255 are updated (warps with ch == 1 compute, all warps load).
template <bool HALF>
__device__ __forceinline__ void rank_update(float (&w)[8][4], const LayerParams &p, Smem &sm, int r0, int c,
                                            int kbeg, int kend, int tid, int rg, int ch, int lane) {
    const int P = (kend - kbeg) / KP;
    const float *__restrict__ U = p.U;
    const float *__restrict__ Wg = p.W;
    const size_t ld = (size_t)p.d_col;
    auto issue = [&](int pc) {
        if (pc < P) {
            const int k0 = kbeg + KP * pc, st = pc % S;
            float *us = sm.u.pipe.Us + st * US_FLOATS;
            float *es = sm.u.pipe.Es + st * ES_FLOATS;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int id = tid + NT * m, row = id >> 6, c16 = id & 63;
                if (!HALF || c16 >= 32) cp_async16(us + row * 256 + 4 * c16, U + (size_t)(k0 + row) * ld + c + 4 * c16);
            }
            if (tid < 128) {
                const int row = tid >> 2, part = tid & 3;
                const int gr = min(r0 + row, p.d_row - 1);
                cp_async16(es + row * KP + 4 * part, Wg + (size_t)gr * ld + k0 + 4 * part);
            }
        }
        cp_async_commit();
    };
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int s = 0; s < S - 1; ++s) issue(s);
    for (int pc = 0; pc < P; ++pc) {
        cp_async_wait<S - 2>();
        __syncthreads();
        issue(pc + S - 1);
        if (!HALF || ch == 1) {
            const float *us = sm.u.pipe.Us + (pc % S) * US_FLOATS + ch * 128 + 4 * lane;
            const float *es = sm.u.pipe.Es + (pc % S) * ES_FLOATS + (8 * rg) * KP;
#pragma unroll
            for (int kk = 0; kk < KP; kk += 4) {
                float4 e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = *reinterpret_cast<const float4 *>(es + i * KP + kk);
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const float4 u = *reinterpret_cast<const float4 *>(us + (kk + k2) * 256);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float ev = k2 == 0 ? e[i].x : k2 == 1 ? e[i].y : k2 == 2 ? e[i].z : e[i].w;
                        acc[i][0] = __fmaf_rn(ev, u.x, acc[i][0]);
                        acc[i][1] = __fmaf_rn(ev, u.y, acc[i][1]);
                        acc[i][2] = __fmaf_rn(ev, u.z, acc[i][2]);
                        acc[i][3] = __fmaf_rn(ev, u.w, acc[i][3]);
                    }
                }
            }
        }
        if ((pc & 7) == 7) {   // end of one earlier 128-column block: w <- w - acc  (gptq.py:270, alpha = -1)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    w[i][j] = __fsub_rn(w[i][j], acc[i][j]);
                    acc[i][j] = 0.0f;
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// The 128 sequential column steps of one block (gptq.py:229-268).  8 lanes per row; lane q8 holds the
// block's columns {8s + q8}.  Column i = 8s+q is broadcast from its owner, every lane of the row redoes the
// (cheap) quantise/err arithmetic, then updates its not-yet-consumed columns:  w -= fl(err * U[i, j])  (no FMA).
template <int QT>
__device__ __forceinline__ void serial_block(Smem &sm, int blk, int srow, int q8) {
    constexpr int GS = Fmt<QT>::GS;
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    float w[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) w[s] = sm.Wt[wt_idx(srow, blk * 128 + s * 8 + q8)];
    const float d = sm.rs.d[srow], dm = sm.rs.dm[srow];
    const float *Ud = sm.u.Ud;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int g = (blk * 128 + s * 8) / GS;
        const float sc = __fmul_rn(d, kq_code_to_f<QT>(sm.rs.sq[srow][g]));
        const float zz = __fmul_rn(dm, kq_code_to_f<QT>(sm.rs.zq[srow][g]));
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const int i = s * 8 + q;
            const float x = __shfl_sync(0xffffffffu, w[s], q, 8);
            const float qv = kq_quant(x, sc, zz, lo, hi);                       // :247-254
            const float wq = kq_dequant(qv, sc, zz);                            // :255-261
            const float err = __fdiv_rn(__fsub_rn(x, wq), Ud[i * 128 + i]);     // :264
            if (q8 == q) {
                sm.Et[srow * 128 + i] = err;                                    // :268
                sm.codes[srow * 256 + blk * 128 + i] = (uint8_t)(int8_t)(int)qv;  // :263
                sm.Wt[wt_idx(srow, blk * 128 + i)] = wq;                        // :266
            }
            const float *urow = Ud + i * 128 + q8;
#pragma unroll
            for (int s2 = s; s2 < 16; ++s2) w[s2] = __fsub_rn(w[s2], __fmul_rn(err, urow[s2 * 8]));  // :267
        }
    }
}

template <int QT>
__global__ void __launch_bounds__(NT, 1) gptq_layer_kernel(const LayerParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rg = warp >> 1, ch = warp & 1;            // rank-update mapping: rows 8rg..8rg+7, cols ch*128+4*lane..
    const int srow = warp * 4 + (lane >> 3), q8 = lane & 7;  // serial mapping
    const int r0 = blockIdx.x * R;
    const int nsb = p.d_col / GQ_QK_K, ng = p.d_col / GS;
    const size_t ld = (size_t)p.d_col;

    auto load_Ud = [&](int c1) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int id = tid + NT * m, row = id >> 5, c16 = id & 31;
            cp_async16(sm.u.Ud + row * 128 + 4 * c16, p.U + (size_t)(c1 + row) * ld + c1 + 4 * c16);
        }
        cp_async_commit();
    };
    auto store_E = [&](int c1) {   // errors of the block replace the consumed columns of W
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int id = tid + NT * m, row = id >> 5, c4 = id & 31;
            if (r0 + row < p.d_row)
                *reinterpret_cast<float4 *>(p.W + (size_t)(r0 + row) * ld + c1 + 4 * c4) =
                    *reinterpret_cast<const float4 *>(sm.Et + row * 128 + 4 * c4);
        }
    };

    for (int sb = 0; sb < nsb; ++sb) {
        const int c = sb * GQ_QK_K;
        float w[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gr = min(r0 + 8 * rg + i, p.d_row - 1);
            const float4 v = *reinterpret_cast<const float4 *>(p.W + (size_t)gr * ld + c + ch * 128 + 4 * lane);
            w[i][0] = v.x; w[i][1] = v.y; w[i][2] = v.z; w[i][3] = v.w;
        }
        __syncthreads();   // previous super-block is completely done with the shared buffers
        rank_update<false>(w, p, sm, r0, c, 0, c, tid, rg, ch, lane);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4 *>(sm.Wt + wt_idx4(8 * rg + i, ch * 32 + lane)) =
                make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
        __syncthreads();

        // scale / min search on the live tile (gptq.py:240-245 -> quant_utils.py:90-145); U's diagonal
        // block for the first 128 columns streams in underneath it.
        load_Ud(c);
        uint32_t vmask = 0, amask = 0;
        tile_search<QT, R, NT>(sm.Wt, sm.gsc, sm.gzr, p.sp, vmask, amask);
        publish_flags(p.flags ? p.flags + 2 * sb : nullptr, vmask, amask);
        __syncthreads();
        if (tid < R) {
            tile_finalize_row<QT, R>(tid, sm.gsc, sm.gzr, sm.rs);
            if (r0 + tid < p.d_row) {
                const size_t gr = (size_t)(r0 + tid);
                p.d[gr * nsb + sb] = sm.rs.dbits[tid];
                p.dmin[gr * nsb + sb] = sm.rs.dmbits[tid];
#pragma unroll
                for (int g = 0; g < GPR; ++g) {
                    p.sq[gr * ng + sb * GPR + g] = sm.rs.sq[tid][g];
                    p.zq[gr * ng + sb * GPR + g] = sm.rs.zq[tid][g];
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();

        serial_block<QT>(sm, 0, srow, q8);
        __syncthreads();
        store_E(c);
        __syncthreads();   // E of block 0 visible to the whole CTA; Ud is free again

        // the first block's rank-k update onto the super-block's second half
        if (ch == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = *reinterpret_cast<const float4 *>(sm.Wt + wt_idx4(8 * rg + i, 32 + lane));
                w[i][0] = v.x; w[i][1] = v.y; w[i][2] = v.z; w[i][3] = v.w;
            }
        }
        rank_update<true>(w, p, sm, r0, c, c, c + 128, tid, rg, ch, lane);
        if (ch == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4 *>(sm.Wt + wt_idx4(8 * rg + i, 32 + lane)) =
                    make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
        }
        load_Ud(c + 128);
        cp_async_wait<0>();
        __syncthreads();

        serial_block<QT>(sm, 1, srow, q8);
        __syncthreads();
        store_E(c + 128);

        // outputs of the finished super-block: codes, GGUF bytes, dequantised weights
        tile_emit<QT, R, NT>(sm.Wt, sm.codes, sm.rs, r0, p.d_row, ld, c, sb, nsb, p.qweight, p.packed, p.wdeq,
                             p.wdeq_dtype);
    }
}

template <int QT> int launch_layer(const LayerParams &p, cudaStream_t st) {
    const size_t smem = sizeof(Smem);
    GQ_CHECK_CUDA(cudaFuncSetAttribute(gptq_layer_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (p.d_row + R - 1) / R;
    gptq_layer_kernel<QT><<<grid, NT, smem, st>>>(p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

}  // namespace

void gq_fill_search_params(SearchParams &sp, int maxq, double rmin, double rdelta, int nstep);

extern "C" int gq_gptq_quantize(float *W, const float *U, int d_row, int d_col, int qtype, int block_size,
                                double rmin, double rdelta, int nstep, int mode, void *qweight, uint16_t *d,
                                void *sq, uint16_t *dmin, void *zq, uint8_t *packed, void *wdeq, int wdeq_dtype,
                                uint32_t *search_flags, gq_stream_t stream) {
    FmtInfo f;
    GQ_REQUIRE(gq_fmt_info(qtype, f), "gq_gptq_quantize: unknown q_type %d", qtype);
    GQ_REQUIRE(W && U && qweight && d && sq && dmin && zq, "gq_gptq_quantize: null pointer");
    GQ_REQUIRE(d_row > 0 && d_col > 0 && d_col % GQ_QK_K == 0, "gq_gptq_quantize: d_col=%d must be a positive multiple of 256", d_col);
    GQ_REQUIRE(nstep >= 0 && nstep < 64, "gq_gptq_quantize: nstep=%d out of range [0,63]", nstep);
    GQ_REQUIRE(((uintptr_t)W | (uintptr_t)U | (uintptr_t)qweight) % 16 == 0, "gq_gptq_quantize: W, U, qweight must be 16-byte aligned");
    GQ_REQUIRE(wdeq == nullptr || ((uintptr_t)wdeq % 16 == 0 && wdeq_dtype >= GQ_F32 && wdeq_dtype <= GQ_BF16),
               "gq_gptq_quantize: bad wdeq");
    if (block_size != 128) {
        gq_set_error("gq_gptq_quantize: block_size=%d not implemented (only 128, the run_quant.sh default)", block_size);
        return GQ_ERR_UNSUPPORTED;
    }
    if (mode != GQ_MODE_EXACT) {
        gq_set_error("gq_gptq_quantize: mode=%d not implemented in this build (GQ_MODE_EXACT only)", mode);
        return GQ_ERR_UNSUPPORTED;
    }
    LayerParams p;
    p.W = W; p.U = U; p.d_row = d_row; p.d_col = d_col;
    gq_fill_search_params(p.sp, (1 << f.bits) - 1, rmin, rdelta, nstep);
    p.qweight = (uint8_t *)qweight; p.d = d; p.sq = (uint8_t *)sq; p.dmin = dmin; p.zq = (uint8_t *)zq;
    p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = wdeq_dtype; p.flags = search_flags;
    cudaStream_t st = (cudaStream_t)stream;
    switch (qtype) {
    case GQ_Q2_K: return launch_layer<GQ_Q2_K>(p, st);
    case GQ_Q3_K: return launch_layer<GQ_Q3_K>(p, st);
    case GQ_Q4_K: return launch_layer<GQ_Q4_K>(p, st);
    case GQ_Q5_K: return launch_layer<GQ_Q5_K>(p, st);
    default: return launch_layer<GQ_Q6_K>(p, st);
    }
}
