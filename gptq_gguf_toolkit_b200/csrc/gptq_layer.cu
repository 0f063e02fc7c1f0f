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
#include "f32x2.cuh"
#include "gemm_tf32.cuh"
#include "tile.cuh"
#include "rank_update.cuh"
#include "exact_update_v2.cuh"
#include <cmath>
#include <cstdlib>

namespace {

using rk::R;             // rows per CTA, threads per CTA, k's per pipeline piece, pipeline stages: rank_update.cuh
using rk::NT;
using rk::KP;
using rk::S;
using rk::US_FLOATS;
using rk::ES_FLOATS;

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
    // GQ_MODE_FAST: the kernel handles super-blocks [sb_begin, sb_end) of an already updated W (the rank-k updates
    // between super-blocks run as tcgen05 GEMMs) and also emits the hi/lo TF32 split of its errors.
    int sb_begin, sb_end, fast;
    // skip_bulk: the contributions of all EARLIER super-blocks have already been applied to W by separate launches
    // (fast mode's tcgen05 GEMMs, or exact_update_kernel in the exact right-looking schedule); the kernel then only
    // does the search, the column steps and the in-super-block update of [sb_begin, sb_end).
    int skip_bulk;
    float *e_hi, *e_lo;     // (rows padded to 128) x 256, only in fast mode
    f2_t nz2;                  // {-0.0f, -0.0f}, deliberately a run-time value (see f2_mul_nofuse)
    // static_groups (gptq.py:184-196): d/dmin/sq/zq already hold the scales of EVERY super-block (searched on the
    // original W), the kernel only reads them.  perm (act_order, gptq.py:209-216; needs static_scales): W and U are in
    // permuted column order, loop column c is original column perm[c] and uses that column's group (:233-238);
    // qweight comes out in loop order, the fused pack / dequantised outputs are not available.
    int static_scales;
    const int *perm;
    unsigned long long *clk;   // optional (gq_debug_phase_clocks): 8 per-phase cycle counters summed over CTAs
};

// phase ids of the optional cycle counters
enum { PH_RANK = 0, PH_SEARCH, PH_FINAL, PH_SERIAL0, PH_MID, PH_SERIAL1, PH_EMIT, PH_COUNT };
struct PhaseClock {
    unsigned long long *out;
    long long t;
    __device__ __forceinline__ PhaseClock(unsigned long long *o) : out(o), t(0) {
        if (out && threadIdx.x == 0) t = clock64();
    }
    __device__ __forceinline__ void lap(int ph) {
        if (out && threadIdx.x == 0) {
            const long long n = clock64();
            atomicAdd(out + ph, (unsigned long long)(n - t));
            t = n;
        }
    }
};

struct __align__(16) Smem {
    float Wt[R * 256];                       // live super-block tile, later the dequantised values
    union {
        struct { float Us[S * US_FLOATS]; float Es[S * ES_FLOATS]; } pipe;
        float Ud[128 * 128];                 // diagonal block of U during the serial phase
    } u;
    float Et[R * 128];                       // errors of the current 128-column block
    float Wq[R * 128];                       // dequantised values of the current block (copied into Wt when it is done)
    uint8_t codes[R * 256];
    float gsc[R * 16];
    float gzr[R * 16];
    float dg_b[128];                         // diagonal of the current U block and its checked reciprocal (DivBy)
    float dg_y[128];
    float dummy_f[64];                       // sink of the non-owner lanes' stores in the serial phase
    uint8_t dummy_b[32];
    RowScales<R> rs;
    // act_order only -- kept LAST: launches without a permutation allocate the struct up to here (SMEM_NO_PERM)
    float pc_sc[R * 128];                    // per (row, column of the block) scale, zero, checked 1/scale
    float pc_zz[R * 128];
    float pc_y[R * 128];
};
constexpr size_t SMEM_NO_PERM = offsetof(Smem, pc_sc);

// Shared-memory layout of the (128 x 128) diagonal block of U for the serial phase: in row i the 16 values a
// lane (l8 = j & 7) needs -- columns j = 8s + l8 -- are contiguous (64 B), 16-byte chunks XOR-swizzled by (l8 >> 1) & 3
// so that the eight lanes of a row group read eight different bank groups with one LDS.128 each.
__device__ __forceinline__ int ud_idx(int i, int j) {
    const int l8 = j & 7, sgrp = j >> 3;
    return i * 128 + l8 * 16 + ((((sgrp >> 2) ^ (l8 >> 1)) & 3) << 2) + (sgrp & 3);
}

// The 128 sequential column steps of one block (gptq.py:229-268), all 8 warps (two per SM sub-partition, so that
// one warp's off-chain work fills the other's dependency stalls).  8 lanes per row; lane l8 holds the block's columns
// {8s + l8}, s = 0..15, as 8 packed pairs.  Column i = 8s+q is broadcast from its owner, every lane of the row redoes
// the (cheap) quantise/err arithmetic, then updates its not-yet-consumed columns:  w -= fl(err * U[i, j])
// (two roundings: f2_mul_nofuse then sub.rn.f32x2).
// The two divisions of the dependent chain -- (x + z) / max(s, eps) and (x - w_q) / U[i,i] -- use reciprocals prepared
// off the chain (DivBy, f32x2.cuh).  SAFE = false: branch-free (DivBy::div_fast, stores through a select-ed pointer);
// returns true if some quotient was outside the range in which div_fast is proven exact -- the caller then reruns the
// block with SAFE = true (IEEE fallback inside DivBy::div).  The block's initial values are read from Wt, its
// dequantised values go to `Wq` (a separate buffer), so a rerun starts from unchanged inputs.
template <int QT, bool SAFE>
__device__ __noinline__ bool serial_block_impl(Smem &sm, int blk, int warp, int lane, f2_t nz2, bool per_col) {
    constexpr int GS = Fmt<QT>::GS;
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    const int l8 = lane & 7, srow = warp * 4 + (lane >> 3);
    f2_t pr[8];
#pragma unroll
    for (int m = 0; m < 8; ++m)
        pr[m] = f2_pack(sm.Wt[wt_idx(srow, blk * 128 + 16 * m + l8)], sm.Wt[wt_idx(srow, blk * 128 + 16 * m + 8 + l8)]);
    const float d = sm.rs.d[srow], dm = sm.rs.dm[srow];
    const float *Ud = sm.u.Ud + l8 * 16;
    const int sw = (l8 >> 1) & 3;
    float *et = sm.Et + srow * 128, *wqo = sm.Wq + srow * 128;
    uint8_t *cd = sm.codes + srow * 256 + blk * 128;
    float sc = 0.0f, zz = 0.0f;
    DivBy ds = DivBy::make(1.0f);
    bool bad = false;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        if (!per_col && (8 * s) % GS == 0) {
            const int g = (blk * 128 + 8 * s) / GS;
            sc = __fmul_rn(d, kq_code_to_f<QT>(sm.rs.sq[srow][g]));
            zz = __fmul_rn(dm, kq_code_to_f<QT>(sm.rs.zq[srow][g]));
            ds = DivBy::make(fmaxf(sc, GQ_EPS));
        }
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const int i = 8 * s + q;
            // loads that do not depend on the chain first: U[i, my columns], the diagonal's reciprocal
            const float4 *urow = reinterpret_cast<const float4 *>(Ud + i * 128);
            float4 u[4];
#pragma unroll
            for (int c4 = s >> 2; c4 < 4; ++c4) u[c4] = urow[c4 ^ sw];
            DivBy du;
            du.b = sm.dg_b[i];
            du.y = sm.dg_y[i];
            if (per_col) {          // act_order: every column has its own group (gptq.py:233-238)
                sc = sm.pc_sc[srow * 128 + i];
                zz = sm.pc_zz[srow * 128 + i];
                ds.b = fmaxf(sc, GQ_EPS);
                ds.y = sm.pc_y[srow * 128 + i];
            }
            float plo, phi;
            f2_unpack(pr[s >> 1], plo, phi);
            const float x = __shfl_sync(0xffffffffu, (s & 1) ? phi : plo, q, 8);
            const float t = __fadd_rn(x, zz);
            const float qv = clampf(rintf(SAFE ? ds.div(t) : ds.div_fast(t, bad)), lo, hi);   // :247-254 (kq_quant)
            const float wq = kq_dequant(qv, sc, zz);                                          // :255-261
            const float num = __fsub_rn(x, wq);
            const float err = SAFE ? du.div(num) : du.div_fast(num, bad);                     // :264
            const f2_t e2 = f2_pack(err, err);
#pragma unroll
            for (int c4 = s >> 2; c4 < 4; ++c4) {                                             // :267
                if (2 * c4 >= (s >> 1)) pr[2 * c4] = f2_sub(pr[2 * c4], f2_mul_nofuse(e2, f2_pack(u[c4].x, u[c4].y), nz2));
                pr[2 * c4 + 1] = f2_sub(pr[2 * c4 + 1], f2_mul_nofuse(e2, f2_pack(u[c4].z, u[c4].w), nz2));
            }
            // outputs of column i: the owner lane stores, the others write to a per-lane dummy slot (no branch)
            const bool own = (l8 == q);
            float *ep = own ? et + i : sm.dummy_f + lane;
            float *wp = own ? wqo + i : sm.dummy_f + 32 + lane;
            uint8_t *cp = own ? cd + i : sm.dummy_b + lane;
            *ep = err;                                                                        // :268
            *wp = wq;                                                                         // :266
            *cp = (uint8_t)(int8_t)(int)qv;                                                   // :263
        }
    }
    return bad;
}

template <int QT>
__device__ __forceinline__ void serial_block(Smem &sm, int blk, int warp, int lane, f2_t nz2, bool per_col) {
    const bool bad = serial_block_impl<QT, false>(sm, blk, warp, lane, nz2, per_col);
    if (__any_sync(0xffffffffu, bad)) serial_block_impl<QT, true>(sm, blk, warp, lane, nz2, per_col);   // rare: exact IEEE divisions
    __syncwarp();
    // the block's dequantised values replace the consumed columns of the tile (this warp's 4 rows)
    const int srow = warp * 4 + (lane >> 3), l8 = lane & 7;
#pragma unroll
    for (int s = 0; s < 16; ++s) sm.Wt[wt_idx(srow, blk * 128 + 8 * s + l8)] = sm.Wq[srow * 128 + 8 * s + l8];
}

// (A register-capped build of this kernel for the panel launches of the right-looking schedule -- __launch_bounds__(256, 2),
// 128 registers, two co-resident CTAs per SM -- was measured on B200 and was 3-5 % SLOWER at every shape: the K-quant search
// of the 32-weight-group types spills, and the column steps lose the registers that keep their loads ahead of the chain.)
template <int QT>
__global__ void __launch_bounds__(NT, 1) gptq_layer_kernel(const LayerParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rg = warp >> 1, ch = warp & 1;            // rank-update mapping: rows 8rg..8rg+7, cols ch*128+4*lane..
    const int r0 = blockIdx.x * R;
    const int nsb = p.d_col / GQ_QK_K, ng = p.d_col / GS;
    const size_t ld = (size_t)p.d_col;

    // Diagonal (128 x 128) block of U -> shared memory in the serial phase's permuted layout (ud_idx).  Row i only
    // needs its columns j >= 32*(i/32) (the lanes read whole 16-byte chunks = 32-column spans at and right of the diagonal).
    auto load_Ud = [&](int c1) {
        for (int id = tid; id < 128 * 128; id += NT) {
            const int i = id >> 7, j = id & 127;
            if (j >= (i & ~31)) cp_async4(sm.u.Ud + ud_idx(i, j), p.U + (size_t)(c1 + i) * ld + c1 + j);
        }
        cp_async_commit();
    };
    auto diag_recip = [&]() {      // after Ud has landed: checked reciprocals of the diagonal (off the dependent chain)
        if (tid < 128) {
            const DivBy dv = DivBy::make(sm.u.Ud[ud_idx(tid, tid)]);
            sm.dg_b[tid] = dv.b;
            sm.dg_y[tid] = dv.y;
        }
    };
    auto store_E = [&](int c1) {   // errors of the block replace the consumed columns of W
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int id = tid + NT * m, row = id >> 5, c4 = id & 31;
            if (r0 + row < p.d_row) {
                const float4 e = *reinterpret_cast<const float4 *>(sm.Et + row * 128 + 4 * c4);
                *reinterpret_cast<float4 *>(p.W + (size_t)(r0 + row) * ld + c1 + 4 * c4) = e;
                if (p.fast) {      // operands of the next tcgen05 rank-256 update: hi = TF32 part, lo = remainder
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(e.x) & 0xFFFFE000u); l.x = __fsub_rn(e.x, h.x);
                    h.y = __uint_as_float(__float_as_uint(e.y) & 0xFFFFE000u); l.y = __fsub_rn(e.y, h.y);
                    h.z = __uint_as_float(__float_as_uint(e.z) & 0xFFFFE000u); l.z = __fsub_rn(e.z, h.z);
                    h.w = __uint_as_float(__float_as_uint(e.w) & 0xFFFFE000u); l.w = __fsub_rn(e.w, h.w);
                    const size_t o = (size_t)(r0 + row) * 256 + (c1 & 255) + 4 * c4;
                    *reinterpret_cast<float4 *>(p.e_hi + o) = h;
                    *reinterpret_cast<float4 *>(p.e_lo + o) = l;
                }
            }
        }
    };
    auto per_column_tables = [&](int c1) {      // act_order: scale / zero / checked reciprocal of every (row, column) of a block
        for (int id = tid; id < R * 128; id += NT) {
            const int row = id >> 7, i = id & 127;
            const int ocol = p.perm[c1 + i];
            const size_t gr = (size_t)min(r0 + row, p.d_row - 1);
            const float dd = __half2float(__ushort_as_half(p.d[gr * nsb + (ocol >> 8)]));
            const float dm = __half2float(__ushort_as_half(p.dmin[gr * nsb + (ocol >> 8)]));
            const float sc = __fmul_rn(dd, kq_code_to_f<QT>(p.sq[gr * ng + ocol / GS]));
            const float zz = __fmul_rn(dm, kq_code_to_f<QT>(p.zq[gr * ng + ocol / GS]));
            sm.pc_sc[id] = sc;
            sm.pc_zz[id] = zz;
            sm.pc_y[id] = DivBy::make(fmaxf(sc, GQ_EPS)).y;
        }
    };
    const bool per_col = p.perm != nullptr;
    PhaseClock pc(p.clk);
    for (int sb = p.sb_begin; sb < p.sb_end; ++sb) {
        const int c = sb * GQ_QK_K;
        float w[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gr = min(r0 + 8 * rg + i, p.d_row - 1);
            const float4 v = *reinterpret_cast<const float4 *>(p.W + (size_t)gr * ld + c + ch * 128 + 4 * lane);
            w[i][0] = v.x; w[i][1] = v.y; w[i][2] = v.z; w[i][3] = v.w;
        }
        __syncthreads();   // previous super-block is completely done with the shared buffers
        rank_update<false>(w, p, sm.u.pipe.Us, sm.u.pipe.Es, r0, c, 0, p.skip_bulk ? 0 : c, tid, rg, ch, lane);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4 *>(sm.Wt + wt_idx4(8 * rg + i, ch * 32 + lane)) =
                make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
        __syncthreads();
        pc.lap(PH_RANK);

        // scale / min search on the live tile (gptq.py:240-245 -> quant_utils.py:90-145); U's diagonal
        // block for the first 128 columns streams in underneath it.
        load_Ud(c);
        if (!p.static_scales) {
            uint32_t vmask = 0, amask = 0;
            tile_search<QT, R, NT>(sm.Wt, sm.gsc, sm.gzr, p.sp, vmask, amask);
            publish_flags(p.flags ? p.flags + 2 * sb : nullptr, vmask, amask);
        }
        cp_async_wait<0>();
        __syncthreads();
        pc.lap(PH_SEARCH);
        diag_recip();
        if (tid >= 128 && tid < 128 + R) {
            const int row = tid - 128;
            const size_t gr = (size_t)min(r0 + row, p.d_row - 1);
            if (!p.static_scales) {
                tile_finalize_row<QT, R>(row, sm.gsc, sm.gzr, sm.rs);
                if (r0 + row < p.d_row) {
                    p.d[gr * nsb + sb] = sm.rs.dbits[row];
                    p.dmin[gr * nsb + sb] = sm.rs.dmbits[row];
#pragma unroll
                    for (int g = 0; g < GPR; ++g) {
                        p.sq[gr * ng + sb * GPR + g] = sm.rs.sq[row][g];
                        p.zq[gr * ng + sb * GPR + g] = sm.rs.zq[row][g];
                    }
                }
            } else if (p.perm == nullptr) {       // static_groups: this super-block's scales were searched up front
                sm.rs.dbits[row] = p.d[gr * nsb + sb];
                sm.rs.dmbits[row] = p.dmin[gr * nsb + sb];
                sm.rs.d[row] = __half2float(__ushort_as_half(sm.rs.dbits[row]));
                sm.rs.dm[row] = __half2float(__ushort_as_half(sm.rs.dmbits[row]));
#pragma unroll
                for (int g = 0; g < GPR; ++g) {
                    sm.rs.sq[row][g] = p.sq[gr * ng + sb * GPR + g];
                    sm.rs.zq[row][g] = p.zq[gr * ng + sb * GPR + g];
                }
            }
        }
        if (p.perm != nullptr) per_column_tables(c);
        __syncthreads();
        pc.lap(PH_FINAL);

        serial_block<QT>(sm, 0, warp, lane, p.nz2, per_col);
        __syncthreads();
        pc.lap(PH_SERIAL0);
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
        rank_update<true>(w, p, sm.u.pipe.Us, sm.u.pipe.Es, r0, c, c, c + 128, tid, rg, ch, lane);
        if (ch == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4 *>(sm.Wt + wt_idx4(8 * rg + i, 32 + lane)) =
                    make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
        }
        load_Ud(c + 128);
        cp_async_wait<0>();
        __syncthreads();
        diag_recip();
        if (per_col) per_column_tables(c + 128);
        __syncthreads();
        pc.lap(PH_MID);

        serial_block<QT>(sm, 1, warp, lane, p.nz2, per_col);
        __syncthreads();
        pc.lap(PH_SERIAL1);
        store_E(c + 128);

        // outputs of the finished super-block: codes, GGUF bytes, dequantised weights
        tile_emit<QT, R, NT>(sm.Wt, sm.codes, sm.rs, r0, p.d_row, ld, c, sb, nsb, p.qweight, p.packed, p.wdeq,
                             p.wdeq_dtype);
        pc.lap(PH_EMIT);
    }
}

template <int QT> int launch_layer(const LayerParams &p, cudaStream_t st) {
    const int grid = (p.d_row + R - 1) / R;
    const size_t smem = p.perm == nullptr ? SMEM_NO_PERM : sizeof(Smem);
    GQ_CHECK_CUDA(cudaFuncSetAttribute(gptq_layer_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    gptq_layer_kernel<QT><<<grid, NT, smem, st>>>(p);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

// Exact right-looking schedule: the trailing update of ONE finished super-block (columns c .. c+255, whose propagated
// errors E sit in W[:, c:c+256]) onto every later column, as a grid over (256-column windows) x (32-row groups):
//     W[r, j] <- (W[r, j] - chain(E[r, c:c+128], U[c:c+128, j])) - chain(E[r, c+128:c+256], U[c+128:c+256, j])
// -- per element the very same two roundings-per-128-k's sequence that the left-looking loop of gptq_layer_kernel applies
// for these two blocks (rank_update<false> is shared), so the two schedules are bit-identical; what changes is who does
// the work: with few row groups (a row slice of a projection on one of several GPUs, d_col = 14336) the left-looking
// kernel leaves most SMs idle while each CTA walks its super-blocks alone, here every SM takes part in every update.
__global__ void __launch_bounds__(NT, 2) exact_update_kernel(const LayerParams p, const int c) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    exact_update_body(p, c, smem_raw);
}

int launch_exact_update(const LayerParams &p, int c, cudaStream_t st) {
    const size_t smem = (size_t)S * (US_FLOATS + ES_FLOATS) * sizeof(float);
    // per launch, like launch_layer: the attribute belongs to the current device's context, a process may use several
    GQ_CHECK_CUDA(cudaFuncSetAttribute(exact_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nwin = (p.d_col - c - GQ_QK_K) / GQ_QK_K;
    if (nwin <= 0) return GQ_OK;
    dim3 grid(nwin, (p.d_row + R - 1) / R);
    const char *v2 = getenv("GQ_UPDATE_V2");      // experimental variant, see exact_update_v2_kernel
    if (v2 && v2[0] == '1') {
        static_assert(upd2::R == R && upd2::KP == KP && upd2::S == S, "exact_update_v2.cuh must use the pipeline geometry of this file");
        const upd2::Params p2{p.W, p.U, p.d_row, p.d_col};
        GQ_CHECK_CUDA(cudaFuncSetAttribute(exact_update_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd2::SMEM_BYTES));
        exact_update_v2_kernel<<<grid, upd2::NT2, upd2::SMEM_BYTES, st>>>(p2, c);
        gq_count_launches(1);
        GQ_CHECK_CUDA(cudaGetLastError());
        return GQ_OK;
    }
    exact_update_kernel<<<grid, NT, smem, st>>>(p, c);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

// Which exact schedule?  Measured on B200 (profiles/r01_ncu_summary.md, Q4_K, both bit-identical): the right-looking
// schedule wins at every Llama-3-8B shape -- 28672 x 4096: 19.6 vs 21.5 ms, 4096 x 14336: 24.2 vs 27.9 ms, and by 2-3.5x on
// the row slices a rank of a multi-GPU run launches (512 x 14336: 7.7 vs 27.2 ms) -- because the trailing update runs
// with two co-resident CTAs per SM and no idle tail wave, while the left-looking kernel does it inside one 255-register
// CTA per SM.  The single left-looking launch is kept for narrow layers (fewer than 4 super-blocks), where two launches
// per super-block buy nothing.
bool exact_prefers_right_looking(int d_row, int d_col) {
    (void)d_row;
    return d_col / GQ_QK_K >= 4;
}

}  // namespace

void gq_fill_search_params(SearchParams &sp, int maxq, double rmin, double rdelta, int nstep);

// ---- optional kernel-level profiling of the fast path (bench.py: rank-k GEMM time measured live with CUDA events) ----
#include <vector>
namespace {
struct ProfEvent { cudaEvent_t a, b; int kind; };   // kind 0 = fused search/column-loop kernel, 1 = rank-k update launch (tcgen05 GEMM / exact_update_kernel)
bool g_prof_on = false;
std::vector<ProfEvent> g_prof;
struct ProfScope {
    cudaStream_t st; int kind; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(cudaStream_t s, int k) : st(s), kind(k) {
        if (g_prof_on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (a) { cudaEventRecord(b, st); g_prof.push_back({a, b, kind}); }
    }
};
}  // namespace

unsigned long long *g_phase_clk = nullptr;
// debug hook (not part of the reference-facing API): device array of 8 u64 that the column-loop kernel adds its
// per-phase cycle counts to (thread 0 of every CTA); nullptr switches the counters off.
extern "C" GQ_API void gq_debug_phase_clocks(unsigned long long *dev8) { g_phase_clk = dev8; }
extern "C" GQ_API void gq_profile_enable(int on) { g_prof_on = on != 0; }
// Synchronises, sums the recorded spans by kind (milliseconds, launch counts), clears the record.
extern "C" GQ_API int gq_profile_read(float ms[2], int counts[2]) {
    ms[0] = ms[1] = 0.f; counts[0] = counts[1] = 0;
    for (auto &e : g_prof) {
        if (cudaEventSynchronize(e.b) != cudaSuccess) return GQ_ERR_CUDA;
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        ms[e.kind] += t; counts[e.kind] += 1;
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    g_prof.clear();
    return GQ_OK;
}

namespace {

// Ut = U^T split into TF32 hi / lo parts (fast mode's B operand must be K-contiguous: B[n][k] = U[k][n] = Ut[n][k])
__global__ void __launch_bounds__(256) transpose_split_kernel(const float *__restrict__ U, int n, float *__restrict__ hi,
                                                              float *__restrict__ lo) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) t[r][tx] = U[(size_t)(r0 + r) * n + c0 + tx];
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        const float x = t[tx][c];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[(size_t)(c0 + c) * n + r0 + tx] = h;
        lo[(size_t)(c0 + c) * n + r0 + tx] = __fsub_rn(x, h);
    }
}

size_t fast_ws_bytes(int d_row, int d_col) {
    const size_t n = (size_t)d_col, mp = ((size_t)d_row + 127) / 128 * 128;
    return 2 * n * n * sizeof(float) + 2 * mp * 256 * sizeof(float) + 4096;
}

template <int QT> int run_layer(LayerParams p, int mode, void *ws, size_t ws_bytes, cudaStream_t st) {
    const int nsb = p.d_col / GQ_QK_K;
    p.fast = 0; p.skip_bulk = 0; p.e_hi = p.e_lo = nullptr; p.sb_begin = 0; p.sb_end = nsb;
    if (mode != GQ_MODE_FAST) {
        // Two bit-identical schedules of the exact arithmetic: ONE left-looking launch (each CTA applies all earlier
        // blocks to its own tile), or per super-block a panel launch + exact_update_kernel over the whole trailing part.
        bool right = mode == GQ_MODE_EXACT_RIGHT;
        if (mode == GQ_MODE_EXACT) {
            static int forced = -1;      // GQ_EXACT_SCHEDULE=left|right overrides the cost model (ablation runs)
            if (forced < 0) {
                const char *e = getenv("GQ_EXACT_SCHEDULE");
                forced = (e && e[0] == 'l') ? 1 : (e && e[0] == 'r') ? 2 : 0;
            }
            right = forced == 2 || (forced == 0 && exact_prefers_right_looking(p.d_row, p.d_col));
        }
        if (!right || nsb < 2) {
            ProfScope ps(st, 0);
            return launch_layer<QT>(p, st);
        }
        p.skip_bulk = 1;
        for (int sb = 0; sb < nsb; ++sb) {
            p.sb_begin = sb; p.sb_end = sb + 1;
            int rc;
            {
                ProfScope ps(st, 0);
                rc = launch_layer<QT>(p, st);
            }
            if (rc) return rc;
            {
                ProfScope ps(st, 1);
                rc = launch_exact_update(p, sb * GQ_QK_K, st);
            }
            if (rc) return rc;
        }
        return GQ_OK;
    }
    // ---- GQ_MODE_FAST: right-looking at super-block granularity.  Per 256-column super-block one launch of the fused
    // search / column-loop kernel, then ONE tcgen05 3xTF32 GEMM  W[:, c+256:] -= E[:, c:c+256] * U[c:c+256, c+256:].
    if (ws == nullptr || ws_bytes < fast_ws_bytes(p.d_row, p.d_col)) {
        gq_set_error("gq_gptq_quantize: fast mode needs %zu workspace bytes", fast_ws_bytes(p.d_row, p.d_col));
        return GQ_ERR_WORKSPACE;
    }
    const int n = p.d_col, mp = (p.d_row + 127) / 128 * 128;
    float *ut_hi = reinterpret_cast<float *>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    float *ut_lo = ut_hi + (size_t)n * n;
    float *e_hi = ut_lo + (size_t)n * n;
    float *e_lo = e_hi + (size_t)mp * 256;
    GQ_CHECK_CUDA(cudaMemsetAsync(e_hi, 0, 2 * (size_t)mp * 256 * sizeof(float), st));   // padded rows stay zero
    transpose_split_kernel<<<dim3(n / 32, n / 32), 256, 0, st>>>(p.U, n, ut_hi, ut_lo);
    gq_count_launches(1);
    p.fast = 1; p.skip_bulk = 1; p.e_hi = e_hi; p.e_lo = e_lo;
    for (int sb = 0; sb < nsb; ++sb) {
        p.sb_begin = sb; p.sb_end = sb + 1;
        int rc;
        {
            ProfScope ps(st, 0);
            rc = launch_layer<QT>(p, st);
        }
        if (rc) return rc;
        const int c = sb * GQ_QK_K, ntrail = n - c - GQ_QK_K;
        if (ntrail > 0) {
            tg::PreSplit A{e_hi, e_lo, 256, mp, 0, 0};
            tg::PreSplit B{ut_hi, ut_lo, n, n, c + GQ_QK_K, c};
            {
                ProfScope ps(st, 1);
                rc = tg::gemm_tf32x3_nt_presplit(A, B, p.W + c + GQ_QK_K, p.d_col, mp, p.d_row, ntrail, GQ_QK_K, -1.0f, 1.0f, st);
            }
            if (rc) return rc;
        }
    }
    return GQ_OK;
}

}  // namespace

extern "C" size_t gq_gptq_workspace_bytes(int d_row, int d_col, int mode) {
    return mode == GQ_MODE_FAST ? fast_ws_bytes(d_row, d_col) : 0;
}

int gq_search_all_superblocks(const float *W, int d_row, int d_col, int qtype, double rmin, double rdelta, int nstep,
                              uint16_t *d, uint16_t *dmin, void *sq, void *zq, uint32_t *search_flags, cudaStream_t st);

extern "C" int gq_gptq_quantize_ex(float *W, const float *U, int d_row, int d_col, int qtype, int block_size,
                                   double rmin, double rdelta, int nstep, int mode, int static_groups, const int *perm,
                                   void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq, uint8_t *packed, void *wdeq,
                                   int wdeq_dtype, uint32_t *search_flags, void *workspace, size_t ws_bytes, gq_stream_t stream) {
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
    GQ_REQUIRE(mode >= GQ_MODE_EXACT && mode <= GQ_MODE_EXACT_RIGHT, "gq_gptq_quantize: unknown mode %d", mode);
    GQ_REQUIRE(static_groups >= 0 && static_groups <= 2, "gq_gptq_quantize: static_groups=%d must be 0, 1 or 2", static_groups);
    if (qtype == GQ_Q3_K) { static_groups = 0; perm = nullptr; }      // gptq.py:204-206: Q3_K ignores both options
    GQ_REQUIRE(perm == nullptr || static_groups == 2,
               "gq_gptq_quantize: act_order (perm) needs static_groups = 2 (scales searched on the un-permuted W beforehand)");
    GQ_REQUIRE(perm == nullptr || (packed == nullptr && wdeq == nullptr),
               "gq_gptq_quantize: with perm the codes come out in loop order; pack / dequantise after un-permuting them");
    if ((static_groups || perm) && mode == GQ_MODE_FAST) {
        gq_set_error("gq_gptq_quantize: static_groups / act_order are implemented for GQ_MODE_EXACT only");
        return GQ_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (static_groups == 1) {     // gptq.py:184-196: all scales / zeros up front, on the weights as they are now
        const int rc = gq_search_all_superblocks(W, d_row, d_col, qtype, rmin, rdelta, nstep, d, dmin, sq, zq, search_flags, st);
        if (rc) return rc;
    }
    LayerParams p;
    p.W = W; p.U = U; p.d_row = d_row; p.d_col = d_col;
    gq_fill_search_params(p.sp, (1 << f.bits) - 1, rmin, rdelta, nstep);
    p.qweight = (uint8_t *)qweight; p.d = d; p.sq = (uint8_t *)sq; p.dmin = dmin; p.zq = (uint8_t *)zq;
    p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = wdeq_dtype; p.flags = search_flags;
    p.clk = g_phase_clk;
    p.nz2 = F2_NEG_ZERO2;
    p.static_scales = static_groups != 0; p.perm = perm;
    switch (qtype) {
    case GQ_Q2_K: return run_layer<GQ_Q2_K>(p, mode, workspace, ws_bytes, st);
    case GQ_Q3_K: return run_layer<GQ_Q3_K>(p, mode, workspace, ws_bytes, st);
    case GQ_Q4_K: return run_layer<GQ_Q4_K>(p, mode, workspace, ws_bytes, st);
    case GQ_Q5_K: return run_layer<GQ_Q5_K>(p, mode, workspace, ws_bytes, st);
    default: return run_layer<GQ_Q6_K>(p, mode, workspace, ws_bytes, st);
    }
}

extern "C" int gq_gptq_quantize(float *W, const float *U, int d_row, int d_col, int qtype, int block_size,
                                double rmin, double rdelta, int nstep, int mode, void *qweight, uint16_t *d,
                                void *sq, uint16_t *dmin, void *zq, uint8_t *packed, void *wdeq, int wdeq_dtype,
                                uint32_t *search_flags, void *workspace, size_t ws_bytes, gq_stream_t stream) {
    return gq_gptq_quantize_ex(W, U, d_row, d_col, qtype, block_size, rmin, rdelta, nstep, mode, 0, nullptr, qweight, d, sq, dmin,
                               zq, packed, wdeq, wdeq_dtype, search_flags, workspace, ws_bytes, stream);
}
