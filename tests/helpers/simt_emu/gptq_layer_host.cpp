// TEST-ONLY: the SHIPPED fused column-loop kernel (gptq_gguf_toolkit_b200/csrc/gptq_layer_kernel.cuh: scale search, the 256
// dependent column steps per super-block with their warp shuffles, in-super-block update, left-looking bulk update, GGUF pack)
// and the trailing-update body (rank_update.cuh) on the SIMT emulator, with the CUDA intrinsics from tests/helpers/host_shim.
#define SIMT_EMU 1
#define GQ_HOST_SHIM 1
#include "simt_emu.h"
#include "host_shim_intrinsics.h"
#include <deque>
#include <map>
#include <utility>

static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
// PRMT: byte n of the result = byte (selector nibble n) of {x (bytes 0-3), y (bytes 4-7)}; the replicate-sign mode (nibble bit 3) is not used
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t xy = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int n = 0; n < 4; ++n) r |= (uint32_t)((xy >> (8 * ((s >> (4 * n)) & 7))) & 0xFF) << (8 * n);
    return r;
}
static inline long long clock64() { return 0; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
// cp.async as late as legal (see exact_update_host.cpp)
namespace cpa {
struct Copy { void *dst; const void *src; int n; };
struct Thread { std::vector<Copy> open; std::deque<std::vector<Copy>> groups; };
static std::map<int, Thread> g_threads;
static Thread &me() { return g_threads[(int)threadIdx.x]; }
}  // namespace cpa
static inline void cp_async16(void *dst, const void *src) { cpa::me().open.push_back({dst, src, 16}); }
static inline void cp_async4(void *dst, const void *src) { cpa::me().open.push_back({dst, src, 4}); }
static inline void cp_async_commit() { auto &t = cpa::me(); t.groups.push_back(std::move(t.open)); t.open.clear(); }
template <int N> static inline void cp_async_wait() {
    auto &t = cpa::me();
    while ((int)t.groups.size() > N) {
        for (auto &c : t.groups.front()) std::memcpy(c.dst, c.src, c.n);
        t.groups.pop_front();
    }
}

#include "f32x2.cuh"
#include "tile.cuh"
#include "rank_update.cuh"
#include "rtn_native.cuh"

namespace {
alignas(16) uint8_t smem_raw[224 * 1024];      // the kernel's `extern __shared__ ... smem_raw[]` (declared inside this namespace)
using rk::R; using rk::NT; using rk::KP; using rk::S; using rk::US_FLOATS; using rk::ES_FLOATS;
#include "gptq_layer_kernel.cuh"

template <int QT> void run(LayerParams p, int right_looking) {
    static_assert(sizeof(Smem) <= sizeof(smem_raw), "shared memory array too small");
    const int nsb = p.d_col / 256, grid = (p.d_row + R - 1) / R;
    p.fast = 0; p.e_hi = p.e_lo = nullptr;
    if (!right_looking) {
        p.skip_bulk = 0; p.sb_begin = 0; p.sb_end = nsb;
        simt::launch(dim3(grid), dim3(NT), [&]() { gptq_layer_kernel<QT>(p); });
        return;
    }
    p.skip_bulk = 1;
    for (int sb = 0; sb < nsb; ++sb) {       // run_layer's exact right-looking schedule (csrc/gptq_layer.cu)
        p.sb_begin = sb; p.sb_end = sb + 1;
        simt::launch(dim3(grid), dim3(NT), [&]() { gptq_layer_kernel<QT>(p); });
        const int nwin = (p.d_col - sb * 256 - 256) / 256;
        if (nwin > 0) simt::launch(dim3(nwin, grid), dim3(NT), [&]() { exact_update_body(p, sb * 256, smem_raw); });
    }
}
}  // namespace

// ---- static_groups (gptq.py:184-196): all scales up front = the RTN search over the whole matrix (gq_search_all_superblocks)
namespace {
#include "rtn_structs.cuh"          // rtn.cu's own RtnParams / RtnSmem (uses R == 32 of this namespace)
RtnSmem g_rtn_sm;
template <int QT> void search_all(const RtnParams &p) {
    simt::launch(dim3((p.d_row + 31) / 32, p.nsb), dim3(256), [&]() { rtn_body<QT, 32, 256, RtnParams, RtnSmem>(p, g_rtn_sm); });
}
}  // namespace
extern "C" int run_search_all(int qtype, const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep, uint16_t *d,
                              uint16_t *dmin, uint8_t *sq, uint8_t *zq) {
    const int bits = qtype == GQ_Q2_K ? 2 : qtype == GQ_Q3_K ? 3 : qtype == GQ_Q4_K ? 4 : qtype == GQ_Q5_K ? 5 : 6;
    const int gs = (qtype == GQ_Q4_K || qtype == GQ_Q5_K) ? 32 : 16;
    RtnParams p;
    p.W = W; p.w_dtype = GQ_F32; p.ld_in = d_col; p.d_row = d_row; p.nsb = d_col / 256;
    p.sp.nstep = nstep;
    for (int i = 0; i <= nstep && i < 64; ++i) p.sp.num[i] = (float)(rmin + rdelta * (double)i + (double)((1 << bits) - 1));
    p.d = d; p.dmin = dmin; p.d_stride = p.nsb; p.sq = sq; p.zq = zq; p.sq_stride = d_col / gs;
    p.qweight = nullptr; p.packed = nullptr; p.wdeq = nullptr; p.wdeq_dtype = GQ_F32; p.flags = nullptr;
    switch (qtype) {
    case GQ_Q2_K: search_all<GQ_Q2_K>(p); return 0;
    case GQ_Q3_K: search_all<GQ_Q3_K>(p); return 0;
    case GQ_Q4_K: search_all<GQ_Q4_K>(p); return 0;
    case GQ_Q5_K: search_all<GQ_Q5_K>(p); return 0;
    case GQ_Q6_K: search_all<GQ_Q6_K>(p); return 0;
    }
    return -1;
}

// W (d_row x d_col fp32, clobbered with the propagated errors), U row-major upper; outputs like gq_gptq_quantize_ex (fp32 wdeq).
// static_scales != 0: d / dmin / sq / zq hold the scales on entry; perm (may be null): act_order, W and U in permuted order,
// qweight comes out in loop order, packed / wdeq must be null then.
extern "C" int run_gptq_layer_ex(int qtype, int right_looking, float *W, const float *U, int d_row, int d_col, double rmin, double rdelta,
                                 int nstep, int static_scales, const int *perm, uint8_t *qweight, uint16_t *d, uint8_t *sq, uint16_t *dmin,
                                 uint8_t *zq, uint8_t *packed, float *wdeq);
extern "C" int run_gptq_layer(int qtype, int right_looking, float *W, const float *U, int d_row, int d_col, double rmin, double rdelta,
                              int nstep, uint8_t *qweight, uint16_t *d, uint8_t *sq, uint16_t *dmin, uint8_t *zq, uint8_t *packed,
                              float *wdeq) {
    return run_gptq_layer_ex(qtype, right_looking, W, U, d_row, d_col, rmin, rdelta, nstep, 0, nullptr, qweight, d, sq, dmin, zq, packed, wdeq);
}
extern "C" int run_gptq_layer_ex(int qtype, int right_looking, float *W, const float *U, int d_row, int d_col, double rmin, double rdelta,
                                 int nstep, int static_scales, const int *perm, uint8_t *qweight, uint16_t *d, uint8_t *sq, uint16_t *dmin,
                                 uint8_t *zq, uint8_t *packed, float *wdeq) {
    const int bits = qtype == GQ_Q2_K ? 2 : qtype == GQ_Q3_K ? 3 : qtype == GQ_Q4_K ? 4 : qtype == GQ_Q5_K ? 5 : 6;
    LayerParams p;
    p.W = W; p.U = U; p.d_row = d_row; p.d_col = d_col;
    p.sp.nstep = nstep;
    for (int i = 0; i <= nstep && i < 64; ++i) p.sp.num[i] = (float)(rmin + rdelta * (double)i + (double)((1 << bits) - 1));
    p.qweight = qweight; p.d = d; p.sq = sq; p.dmin = dmin; p.zq = zq; p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = GQ_F32;
    p.flags = nullptr; p.clk = nullptr; p.nz2 = F2_NEG_ZERO2; p.static_scales = static_scales; p.perm = perm;
    switch (qtype) {
    case GQ_Q2_K: run<GQ_Q2_K>(p, right_looking); return 0;
    case GQ_Q3_K: run<GQ_Q3_K>(p, right_looking); return 0;
    case GQ_Q4_K: run<GQ_Q4_K>(p, right_looking); return 0;
    case GQ_Q5_K: run<GQ_Q5_K>(p, right_looking); return 0;
    case GQ_Q6_K: run<GQ_Q6_K>(p, right_looking); return 0;
    }
    return -1;
}

// ---- DivRange / div_chain (f32x2.cuh): the dividend bookkeeping of the column steps' fast divisions ----
// ok_out[i] = DivRange that has seen only a[i] is still ok();  q_out[i] = div_chain(a[i], b, 1/b)
extern "C" void run_div_chain(const float *a, int n, float b, int *ok_out, float *q_out) {
    const DivBy d = DivBy::make(b);
    for (int i = 0; i < n; ++i) {
        DivRange rg;
        q_out[i] = div_chain(a[i], d.b, d.y, rg);
        ok_out[i] = rg.ok() ? 1 : 0;
    }
}
