// TEST-ONLY: the trailing-update kernel body of the exact right-looking schedule (csrc/rank_update.cuh) on the SIMT
// emulator.  cp.async is emulated AS LATE AS LEGAL: a copy is only performed when a cp_async_wait<N> of the issuing thread
// retires its group (all but the newest N committed groups), so that reading a pipeline stage before the wait that covers it
// -- or waiting for too few groups -- shows up as stale data and a bit mismatch.
#define SIMT_EMU 1
#include "simt_emu.h"
#include <algorithm>
#include <deque>
#include <map>
#include <utility>
using std::min;
namespace cpa {
struct Copy { void *dst; const void *src; };
struct Thread { std::vector<Copy> open; std::deque<std::vector<Copy>> groups; };
static std::map<int, Thread> g_threads;        // keyed by the linear thread id of the block being run
static Thread &me() { return g_threads[(int)threadIdx.x]; }
static void reset() { g_threads.clear(); }
}  // namespace cpa
static inline void cp_async16(void *dst, const void *src) { cpa::me().open.push_back({dst, src}); }
static inline void cp_async_commit() { auto &t = cpa::me(); t.groups.push_back(std::move(t.open)); t.open.clear(); }
template <int N> static inline void cp_async_wait() {
    auto &t = cpa::me();
    while ((int)t.groups.size() > N) {
        for (auto &c : t.groups.front()) std::memcpy(c.dst, c.src, 16);
        t.groups.pop_front();
    }
}
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fsub_rn(float a, float b) { return a - b; }
alignas(16) unsigned char smem_raw[80 * 1024];       // the kernel's `extern __shared__ ... smem_raw[]`
#include "rank_update.cuh"          // the SHIPPED trailing-update body (exact_update_kernel) and rank_update<>
namespace upd { struct Params { float *W; const float *U; int d_row, d_col; }; }

// the shipped kernel: exact_update_kernel = exact_update_body<LayerParams> (csrc/gptq_layer.cu); 256 threads per CTA
extern "C" int run_exact_update_v1(float *W, const float *U, int d_row, int d_col, int c) {
    static_assert((size_t)rk::S * (rk::US_FLOATS + rk::ES_FLOATS) * sizeof(float) <= sizeof(smem_raw), "shared memory array too small");
    const int nwin = (d_col - c - 256) / 256;
    if (nwin <= 0) return 0;
    upd::Params p{W, U, d_row, d_col};
    simt::launch(dim3(nwin, (d_row + rk::R - 1) / rk::R), dim3(rk::NT), [&]() { exact_update_body(p, c, smem_raw); });
    return 0;
}
