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

#include "gptq_layer_kernel.cuh"      // LayerParams, Smem, serial_block<>, gptq_layer_kernel<QT>

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
