// gptq_layer_kernel.cuh -- the fused column-loop kernel of csrc/gptq_layer.cu (LayerParams, shared-memory layout, the 128
// sequential column steps, gptq_layer_kernel<QT>), in a header of its own so that the CPU suite can run the SHIPPED source on
// the SIMT emulator against the reference goldens (tests/test_simt_emu_cpu.py).  Included INSIDE gptq_layer.cu's anonymous
// namespace, after the `using rk::...` declarations -- it is not a stand-alone header.
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
    int dg_unsafe[4];                        // per warp of diag_recip: some dg_y is 0 (= "use the IEEE division") -> the block takes the SAFE path
    RowScales<R> rs;
    union {
        // act_order only: per (row, column of the block) scale, zero, checked 1/scale
        struct { float pc_sc[R * 128]; float pc_zz[R * 128]; float pc_y[R * 128]; };
        // without a permutation: the off-diagonal block U[c:c+128, c+128:c+256] of the in-super-block update, prefetched underneath
        // the scale search (mid_update_smem)
        float Uoff[128 * 128];
    };
};

// The 128 sequential column steps of one block (gptq.py:229-268), all 8 warps (two per SM sub-partition, so that one warp's
// off-chain work fills the other's dependency stalls).  8 lanes per row.  Lane l8 holds the block's columns in groups of four:
// slot m (0..3) = columns 32m + 4*l8 .. +3, as two packed pairs -- so the diagonal block of U sits in shared memory in plain
// row-major order (16-byte cp.async, and a lane's part of row i is one LDS.128 per slot; the eight lanes of a row read 128
// contiguous bytes).  The columns are consumed in order i = 32m + 4q + p (slot m, owner lane q, position p): at the start of a
// group every lane fetches the owner's four values (four shuffles, once per four columns) and then carries that group REDUNDANTLY
// through its four steps -- quantise, error, w[j] -= fl(err * U[i, j]) on the group's remaining columns -- so the dependent chain
// of a column step holds no shuffle at all: 16 fp32 operations.  Every lane also updates its own not-yet-consumed slots
// (two roundings per element: f2_mul_nofuse then sub.rn.f32x2; the already consumed columns of the slots it touches see U's
// zeros below the diagonal).
// The two divisions of the dependent chain -- (x + z) / max(s, eps) and (x - w_q) / U[i,i] -- use reciprocals prepared
// off the chain (DivBy, f32x2.cuh).  SAFE = false: branch-free Markstein quotients (div_chain) whose dividends are tracked by two
// integer min / max accumulators (DivRange); returns true if a dividend left the range in which the quotient is proven to equal
// the IEEE one, or a divisor has no checked reciprocal -- the caller then reruns the block with SAFE = true (IEEE fallback inside
// DivBy::div).  The block's initial values are read from Wt, its dequantised values go to `Wq` (a separate buffer), so a rerun
// starts from unchanged inputs.  per_col (act_order: every column has its own scale group) is a template parameter: the table
// loads would otherwise sit, predicated off, in every column step.
template <int QT, bool SAFE, bool per_col>
__device__ __noinline__ bool serial_block_impl(Smem &sm, int blk, int warp, int lane, f2_t nz2) {
    constexpr int GS = Fmt<QT>::GS;
    constexpr int NH = 32 / GS;              // scale groups per 32-column span (1 or 2)
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    const int l8 = lane & 7, srow = warp * 4 + (lane >> 3);
    f2_t pr[8];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float4 v = *reinterpret_cast<const float4 *>(sm.Wt + wt_idx4(srow, blk * 32 + 8 * m + l8));
        pr[2 * m] = f2_pack(v.x, v.y);
        pr[2 * m + 1] = f2_pack(v.z, v.w);
    }
    const float d = sm.rs.d[srow], dm = sm.rs.dm[srow];
    const float *Ud = sm.u.Ud + 4 * l8;
    float *et = sm.Et + srow * 128, *wqo = sm.Wq + srow * 128;
    uint8_t *cd = sm.codes + srow * 256 + blk * 128;
    float sc = 0.0f, zz = 0.0f;
    DivBy ds = DivBy::make(1.0f);
    bool bad = (sm.dg_unsafe[0] | sm.dg_unsafe[1] | sm.dg_unsafe[2] | sm.dg_unsafe[3]) != 0;   // a diagonal element of U without a checked reciprocal
    DivRange rng;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            if (!per_col) {
                const int g = (blk * 128 + 32 * m) / GS + h;
                sc = __fmul_rn(d, kq_code_to_f<QT>(sm.rs.sq[srow][g]));
                zz = __fmul_rn(dm, kq_code_to_f<QT>(sm.rs.zq[srow][g]));
                ds = DivBy::make(fmaxf(sc, GQ_EPS));
                bad = bad || ds.y == 0.0f;
            }
#pragma unroll 1
            for (int q = h * (8 / NH); q < (h + 1) * (8 / NH); ++q) {
                // the owner's four columns, replicated in every lane of the row
                float xg[4];
                {
                    float a0, a1, a2, a3;
                    f2_unpack(pr[2 * m], a0, a1);
                    f2_unpack(pr[2 * m + 1], a2, a3);
                    xg[0] = __shfl_sync(0xffffffffu, a0, q, 8);
                    xg[1] = __shfl_sync(0xffffffffu, a1, q, 8);
                    xg[2] = __shfl_sync(0xffffffffu, a2, q, 8);
                    xg[3] = __shfl_sync(0xffffffffu, a3, q, 8);
                }
                // the diagonal's checked reciprocals for the group's four columns, ahead of the chain
                const int i0 = 32 * m + 4 * q;
                const float4 dgb = *reinterpret_cast<const float4 *>(sm.dg_b + i0);
                const float4 dgy = *reinterpret_cast<const float4 *>(sm.dg_y + i0);
                float errs[4], wqs[4];
                uint32_t code4 = 0;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const int i = 32 * m + 4 * q + pp;
                    // loads that do not depend on the chain first: U[i, my columns], U[i, the group's columns], the diagonal's reciprocal
                    const float *urow = Ud + i * 128;
                    float4 u[4];
#pragma unroll
                    for (int c4 = m; c4 < 4; ++c4) u[c4] = *reinterpret_cast<const float4 *>(urow + 32 * c4);
                    const float4 ug = *reinterpret_cast<const float4 *>(sm.u.Ud + i * 128 + 32 * m + 4 * q);
                    DivBy du;
                    du.b = pp == 0 ? dgb.x : pp == 1 ? dgb.y : pp == 2 ? dgb.z : dgb.w;
                    du.y = pp == 0 ? dgy.x : pp == 1 ? dgy.y : pp == 2 ? dgy.z : dgy.w;
                    if (per_col) {          // act_order: every column has its own group (gptq.py:233-238)
                        sc = sm.pc_sc[srow * 128 + i];
                        zz = sm.pc_zz[srow * 128 + i];
                        ds.b = fmaxf(sc, GQ_EPS);
                        ds.y = sm.pc_y[srow * 128 + i];
                        if (!SAFE) bad = bad || ds.y == 0.0f;
                    }
                    const float x = xg[pp];
                    const float t = __fadd_rn(x, zz);
                    uint32_t qbits;
                    const float qv = kq_rint_clamp_bits(SAFE ? ds.div(t) : div_chain(t, ds.b, ds.y, rng), lo, hi, qbits);   // :247-254 (kq_quant)
                    const float wq = kq_dequant(qv, sc, zz);                                          // :255-261
                    const float num = __fsub_rn(x, wq);
                    const float err = SAFE ? du.div(num) : div_chain(num, du.b, du.y, rng);           // :264
                    // the group's remaining columns (the replicated copy): w -= fl(err * u), two roundings     :267
                    if (pp < 1) xg[1] = __fsub_rn(xg[1], __fmul_rn(err, ug.y));
                    if (pp < 2) xg[2] = __fsub_rn(xg[2], __fmul_rn(err, ug.z));
                    if (pp < 3) xg[3] = __fsub_rn(xg[3], __fmul_rn(err, ug.w));
                    const f2_t e2 = f2_pack(err, err);
#pragma unroll
                    for (int c4 = m; c4 < 4; ++c4) {                                                  // :267
                        pr[2 * c4] = f2_sub(pr[2 * c4], f2_mul_nofuse(e2, f2_pack(u[c4].x, u[c4].y), nz2));
                        pr[2 * c4 + 1] = f2_sub(pr[2 * c4 + 1], f2_mul_nofuse(e2, f2_pack(u[c4].z, u[c4].w), nz2));
                    }
                    errs[pp] = err;                                                                   // :268
                    wqs[pp] = wq;                                                                     // :266
                    code4 = __byte_perm(code4, qbits, pp == 0 ? 0x3214 : pp == 1 ? 0x3240 : pp == 2 ? 0x3410 : 0x4210);   // :263
                }
                // outputs of the group's four columns (identical in the row's eight lanes): one lane stores them, 16 + 16 + 4 bytes
                if (l8 == 0) {
                    *reinterpret_cast<float4 *>(et + i0) = make_float4(errs[0], errs[1], errs[2], errs[3]);
                    *reinterpret_cast<float4 *>(wqo + i0) = make_float4(wqs[0], wqs[1], wqs[2], wqs[3]);
                    *reinterpret_cast<uint32_t *>(cd + i0) = code4;
                }
            }
        }
    }
    return bad || !rng.ok();
}

template <int QT>
__device__ __forceinline__ void serial_block(Smem &sm, int blk, int warp, int lane, f2_t nz2, bool per_col) {
    // (two instantiations of the fast path: the act_order tables cost four predicated instructions per column otherwise)
    const bool bad = per_col ? serial_block_impl<QT, false, true>(sm, blk, warp, lane, nz2)
                             : serial_block_impl<QT, false, false>(sm, blk, warp, lane, nz2);
    if (__any_sync(0xffffffffu, bad)) {      // rare: exact IEEE divisions
        if (per_col) serial_block_impl<QT, true, true>(sm, blk, warp, lane, nz2);
        else serial_block_impl<QT, true, false>(sm, blk, warp, lane, nz2);
    }
    __syncwarp();
    // the block's dequantised values replace the consumed columns of the tile (this warp's 4 rows)
    const int srow = warp * 4 + (lane >> 3), l8 = lane & 7;
#pragma unroll
    for (int s = 0; s < 16; ++s) sm.Wt[wt_idx(srow, blk * 128 + 8 * s + l8)] = sm.Wq[srow * 128 + 8 * s + l8];
}

// The in-super-block update  tile[:, 128:256] -= E_0 (32 x 128) * U[c:c+128, c+128:c+256]  (gptq.py:270 for the super-block's first
// 128-column block) with BOTH operands already in shared memory: the errors are the serial phase's own output (Et), the block of U
// was prefetched underneath the scale search (Uoff).  All 8 warps take part -- warp w owns rows 4w..4w+3, lane l the columns
// 128 + 4l .. 128 + 4l + 3 -- where the streaming rank_update<true> kept only the four ch == 1 warps (two of the four SM
// sub-partitions) busy and paid eight pipeline barriers.  Same arithmetic per element: one fresh single-accumulator FMA chain over
// the 128 k's in ascending order, then one subtraction.
__device__ __forceinline__ void mid_update_smem(Smem &sm, int warp, int lane) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const float *ub = sm.Uoff + 4 * lane;
    const float *eb = sm.Et + (4 * warp) * 128;
#pragma unroll 2
    for (int kk = 0; kk < 128; kk += 4) {
        float4 e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = *reinterpret_cast<const float4 *>(eb + i * 128 + kk);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const float4 u = *reinterpret_cast<const float4 *>(ub + (kk + k2) * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float ev = k2 == 0 ? e[i].x : k2 == 1 ? e[i].y : k2 == 2 ? e[i].z : e[i].w;
                acc[i][0] = __fmaf_rn(ev, u.x, acc[i][0]);
                acc[i][1] = __fmaf_rn(ev, u.y, acc[i][1]);
                acc[i][2] = __fmaf_rn(ev, u.z, acc[i][2]);
                acc[i][3] = __fmaf_rn(ev, u.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 *wp = reinterpret_cast<float4 *>(sm.Wt + wt_idx4(4 * warp + i, 32 + lane));
        float4 v = *wp;
        v.x = __fsub_rn(v.x, acc[i][0]); v.y = __fsub_rn(v.y, acc[i][1]);
        v.z = __fsub_rn(v.z, acc[i][2]); v.w = __fsub_rn(v.w, acc[i][3]);
        *wp = v;
    }
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

    // Diagonal (128 x 128) block of U -> shared memory, row-major.  Row i only needs its columns j >= 32*(i/32) (the lanes read
    // whole 32-column spans at and right of the diagonal).
    auto load_Ud = [&](int c1) {
#pragma unroll 4
        for (int id = tid; id < 128 * 32; id += NT) {
            const int i = id >> 5, c4 = id & 31;
            if (4 * c4 >= (i & ~31)) cp_async16(sm.u.Ud + i * 128 + 4 * c4, p.U + (size_t)(c1 + i) * ld + c1 + 4 * c4);
        }
        cp_async_commit();
    };
    auto diag_recip = [&]() {      // after Ud has landed: checked reciprocals of the diagonal (off the dependent chain)
        if (tid < 128) {
            const DivBy dv = DivBy::make(sm.u.Ud[tid * 128 + tid]);
            sm.dg_b[tid] = dv.b;
            sm.dg_y[tid] = dv.y;
            const bool unsafe = __any_sync(0xffffffffu, dv.y == 0.0f);
            if ((tid & 31) == 0) sm.dg_unsafe[tid >> 5] = unsafe ? 1 : 0;
        }
    };
    auto load_Uoff = [&](int c1) {   // U[c1:c1+128, c1+128:c1+256] -> Uoff (row-major), 16 x 16-byte cp.async per thread
#pragma unroll 4
        for (int id = tid; id < 128 * 32; id += NT) {
            const int i = id >> 5, c4 = id & 31;
            cp_async16(sm.Uoff + i * 128 + 4 * c4, p.U + (size_t)(c1 + i) * ld + c1 + 128 + 4 * c4);
        }
        cp_async_commit();
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
        if (!per_col) load_Uoff(c);
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

        // the first block's rank-k update onto the super-block's second half
        if (!per_col) {
            // Ud is free (every warp is past serial_block 0): the next diagonal block streams in underneath the update, whose
            // operands are both in shared memory already
            load_Ud(c + 128);
            mid_update_smem(sm, warp, lane);
        } else {
            __syncthreads();   // E of block 0 visible (global) to the whole CTA
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
        }
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

