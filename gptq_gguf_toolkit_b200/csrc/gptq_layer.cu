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
#include "gemm_f16x3.cuh"
#include "tile.cuh"
#include "rank_update.cuh"
#include <cmath>
#include <cstdlib>

namespace {

using rk::R;             // rows per CTA, threads per CTA, k's per pipeline piece, pipeline stages: rank_update.cuh
using rk::NT;
using rk::KP;
using rk::S;
using rk::US_FLOATS;
using rk::ES_FLOATS;

#include "gptq_layer_kernel.cuh"      // LayerParams, Smem, serial_block<>, gptq_layer_kernel<QT>

template <int QT> int launch_layer(const LayerParams &p, cudaStream_t st) {
    const int grid = (p.d_row + R - 1) / R;
    const size_t smem = sizeof(Smem);
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
struct ProfEvent { cudaEvent_t a, b; int kind; };   // kind 0 = fused search/column-loop kernel, 1 = rank-k update launch (tcgen05 GEMM / exact_update_kernel), 2 = operand preparation of the tcgen05 GEMM
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
extern "C" GQ_API int gq_profile_read3(float ms[3], int counts[3]) {
    for (int k = 0; k < 3; ++k) { ms[k] = 0.f; counts[k] = 0; }
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
extern "C" GQ_API int gq_profile_read(float ms[2], int counts[2]) {      // kinds 0 and 1 only (kind 2 is dropped)
    float m3[3]; int c3[3];
    const int rc = gq_profile_read3(m3, c3);
    ms[0] = m3[0]; ms[1] = m3[1]; counts[0] = c3[0]; counts[1] = c3[1];
    return rc;
}

namespace {

// ---- GQ_MODE_FAST workspace: U^T as a split-fp16 operand (n rows of 2n halves) | its column scales | column-max scratch |
// the split errors of one update (rows padded to 128, up to FAST_KMAX k's) | their row scales
constexpr int FAST_KMAX = 1024;
struct FastWs { __half *ut; float *sb; unsigned int *cmax; __half *e16; float *sa; };
inline size_t al1k(size_t x) { return (x + 1023) / 1024 * 1024; }
size_t fast_ws_bytes(int d_row, int d_col) {
    const size_t n = (size_t)d_col, mp = ((size_t)d_row + 255) / 256 * 256;      // rows padded for the CTA-pair tiles
    return 1024 + al1k(4 * n * n) + 2 * al1k(4 * n) + al1k(mp * 2 * FAST_KMAX * 2) + al1k(4 * mp);
}
FastWs fast_ws_carve(void *ws, int d_row, int d_col) {
    const size_t n = (size_t)d_col, mp = ((size_t)d_row + 255) / 256 * 256;
    uint8_t *b = reinterpret_cast<uint8_t *>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    FastWs w;
    w.ut = (__half *)b; b += al1k(4 * n * n);
    w.sb = (float *)b; b += al1k(4 * n);
    w.cmax = (unsigned int *)b; b += al1k(4 * n);
    w.e16 = (__half *)b; b += al1k(mp * 2 * FAST_KMAX * 2);
    w.sa = (float *)b;
    return w;
}
// Super-blocks per trailing update (GQ_FAST_GROUP = 1, 2 or 4; read once per layer): inside a group of G super-blocks the updates stay
// inside the group (K = 256), the columns behind the group receive ONE update with K = 256 G -- G times fewer passes over
// the trailing part of W (the K = 256 update moves 64 flop per byte of W, HBM-bound well below the tensor pipe).
int fast_group() {
    const char *e = getenv("GQ_FAST_GROUP");
    const int v = e ? atoi(e) : 2;
    return (v == 1 || v == 2 || v == 4) ? v : 2;
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
    // ---- GQ_MODE_FAST: right-looking with the rank-k updates on tcgen05 (split-fp16 GEMM, gemm_f16x3.cu).  Per 256-column
    // super-block one launch of the fused search / column-loop kernel; its errors E (left in W[:, c:c+256]) then update
    //   * the rest of its group of G super-blocks:      W[:, c+256:gend]  -= E_c     U[c:c+256,  c+256:gend]    (K = 256)
    //   * after the group's last super-block, all later columns:  W[:, gend:] -= E_group U[g0:gend, gend:]      (K = 256 G)
    if (ws == nullptr || ws_bytes < fast_ws_bytes(p.d_row, p.d_col)) {
        gq_set_error("gq_gptq_quantize: fast mode needs %zu workspace bytes", fast_ws_bytes(p.d_row, p.d_col));
        return GQ_ERR_WORKSPACE;
    }
    const int n = p.d_col, mp = (p.d_row + 255) / 256 * 256, G = fast_group();
    const FastWs w = fast_ws_carve(ws, p.d_row, p.d_col);
    {
        ProfScope ps(st, 2);
        const int rc = th::transpose_split_upper_f16(p.U, n, w.ut, w.sb, w.cmax, st);
        if (rc) return rc;
    }
    auto update = [&](int k0, int K, int n0, int N) -> int {
        {
            ProfScope ps(st, 2);
            const int rc = th::split_rows_f16(p.W + k0, p.d_col, 0, p.d_row, mp, K, K, 1, w.e16, w.sa, st);
            if (rc) return rc;
        }
        const th::Split16 A{w.e16, w.sa, 2L * K, mp, 0, 0};
        const th::Split16 B{w.ut, w.sb, 2L * n, n, n0, k0};
        ProfScope ps(st, 1);
        return th::gemm_f16x3_nt_presplit(A, B, p.W + n0, p.d_col, mp, p.d_row, N, K, -1.0f, 1.0f, st);
    };
    p.skip_bulk = 1;
    for (int g0 = 0; g0 < nsb; g0 += G) {
        const int g1 = g0 + G < nsb ? g0 + G : nsb, gend = g1 * GQ_QK_K;
        for (int sb = g0; sb < g1; ++sb) {
            p.sb_begin = sb; p.sb_end = sb + 1;
            int rc;
            {
                ProfScope ps(st, 0);
                rc = launch_layer<QT>(p, st);
            }
            if (rc) return rc;
            const int c = sb * GQ_QK_K;
            if (gend - c - GQ_QK_K > 0) {
                rc = update(c, GQ_QK_K, c + GQ_QK_K, gend - c - GQ_QK_K);
                if (rc) return rc;
            }
        }
        if (n - gend > 0) {
            const int rc = update(g0 * GQ_QK_K, gend - g0 * GQ_QK_K, gend, n - gend);
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

int gq_gptq_blocksize(float *W, const float *U, int d_row, int d_col, int qtype, int B, double rmin, double rdelta, int nstep,
                      bool searched, void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq, uint8_t *packed, void *wdeq,
                      int wdeq_dtype, uint32_t *flags, cudaStream_t st);

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
    if (block_size != 128 && block_size != 32 && block_size != 64 && block_size != 256) {
        gq_set_error("gq_gptq_quantize: block_size=%d not implemented (32, 64, 128 and 256 are; run_quant.sh uses 128)", block_size);
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
    if (block_size != 128) {      // the other block sizes of the reference: plain right-looking schedule (gptq_blocksize.cu)
        if (mode == GQ_MODE_FAST || perm != nullptr || static_groups == 2) {
            gq_set_error("gq_gptq_quantize: block_size=%d is implemented for the exact arithmetic without act_order only", block_size);
            return GQ_ERR_UNSUPPORTED;
        }
        if (static_groups == 1) {
            const int rc = gq_search_all_superblocks(W, d_row, d_col, qtype, rmin, rdelta, nstep, d, dmin, sq, zq, search_flags, st);
            if (rc) return rc;
        }
        return gq_gptq_blocksize(W, U, d_row, d_col, qtype, block_size, rmin, rdelta, nstep, static_groups == 1, qweight, d, sq, dmin,
                                 zq, packed, wdeq, wdeq_dtype, search_flags, st);
    }
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
