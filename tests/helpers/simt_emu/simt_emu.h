// TEST-ONLY: a minimal SIMT emulator for running ONE-CTA-at-a-time CUDA kernels on the host, so that a kernel written without
// access to a GPU can be checked for indexing / algorithm errors before it costs GPU minutes.
//   * every CUDA thread of a block is a ucontext fiber; __syncthreads() yields until all fibers of the block have arrived;
//   * between barriers the fibers run one after the other in thread order, so data races that need a particular interleaving
//     are NOT detected -- this finds deterministic mistakes (wrong index, missing barrier that matters in thread order,
//     wrong formula), nothing more;
//   * shared memory: the kernel's `extern __shared__ ... raw[]` binds to the array the harness defines.
//   * warp shuffles (float), votes and __syncwarp() are warp-level barriers + an exchange buffer.
// Not supported: tensor cores, clusters, real atomics (harnesses provide what a kernel needs), timing.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

using std::isfinite;
using std::isnan;

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__
#define __restrict__

namespace simt {
// Scheduler: fibers run round-robin; a fiber that reaches a barrier (block-wide or warp-wide) registers its arrival and is
// only resumed past it when every live thread of the block / of its warp has arrived (generation counters), so barriers of
// different kinds and warp-divergent code (one warp shuffling while another waits at __syncthreads) keep CUDA's semantics.
struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = false; };
inline dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
inline ucontext_t g_sched;
inline std::vector<Fiber> *g_fibers = nullptr;
inline int g_current = -1, g_nt = 0;
inline std::function<void()> *g_body = nullptr;
inline int g_live = 0;                                  // threads of the block that have not returned yet
inline int g_block_arrived = 0;
inline unsigned long g_block_gen = 0;
inline std::vector<int> g_warp_arrived, g_warp_live;
inline std::vector<unsigned long> g_warp_gen;
inline std::vector<uint32_t> g_xchg;                    // one 32-bit slot per thread for shuffles / votes

inline void yield_to_scheduler() { swapcontext(&(*g_fibers)[g_current].ctx, &g_sched); }
inline void trampoline() {
    (*g_body)();
    const int w = g_current >> 5;
    (*g_fibers)[g_current].done = true;
    --g_live;
    --g_warp_live[w];
    // a thread that leaves must not keep others waiting forever: leaving counts as "no longer expected" at barriers
    if (g_live > 0 && g_block_arrived == g_live) { g_block_arrived = 0; ++g_block_gen; }
    if (g_warp_live[w] > 0 && g_warp_arrived[w] == g_warp_live[w]) { g_warp_arrived[w] = 0; ++g_warp_gen[w]; }
    yield_to_scheduler();
}
inline void syncthreads() {
    const unsigned long gen = g_block_gen;
    if (++g_block_arrived == g_live) { g_block_arrived = 0; ++g_block_gen; return; }
    while (g_block_gen == gen) yield_to_scheduler();
}
inline void syncwarp() {
    const int w = g_current >> 5;
    const unsigned long gen = g_warp_gen[w];
    if (++g_warp_arrived[w] == g_warp_live[w]) { g_warp_arrived[w] = 0; ++g_warp_gen[w]; return; }
    while (g_warp_gen[w] == gen) yield_to_scheduler();
}
// warp exchange: every lane publishes a 32-bit value, then reads what it needs (two warp barriers)
template <class F> inline uint32_t warp_exchange(uint32_t mine, F reader) {
    g_xchg[g_current] = mine;
    syncwarp();
    const uint32_t r = reader((g_current >> 5) << 5);   // reader gets the index of lane 0 of this warp in g_xchg
    syncwarp();
    return r;
}
inline float shfl(float v, int src, int width) {
    uint32_t bits; std::memcpy(&bits, &v, 4);
    const int lane = g_current & 31;
    const int from = (lane / width) * width + (src % width);
    const uint32_t r = warp_exchange(bits, [&](int base) { return g_xchg[base + from]; });
    float out; std::memcpy(&out, &r, 4); return out;
}
inline bool any(bool pred) {
    return warp_exchange(pred ? 1u : 0u, [&](int base) { uint32_t o = 0; for (int l = 0; l < 32 && base + l < g_nt; ++l) o |= g_xchg[base + l]; return o; }) != 0;
}
inline uint32_t reduce_or(uint32_t v) {
    return warp_exchange(v, [&](int base) { uint32_t o = 0; for (int l = 0; l < 32 && base + l < g_nt; ++l) o |= g_xchg[base + l]; return o; });
}

// Runs body() once per thread of every block of the grid; blocks one after the other.
inline void launch(dim3 grid, dim3 block, std::function<void()> body, size_t stack_bytes = 256 * 1024) {
    g_gridDim = grid; g_blockDim = block; g_body = &body;
    const int nt = (int)(block.x * block.y * block.z);
    g_nt = nt;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = dim3(bx, by, bz);
                std::vector<Fiber> fibers(nt);
                g_fibers = &fibers;
                g_live = nt; g_block_arrived = 0; g_block_gen = 0;
                const int nw = (nt + 31) / 32;
                g_warp_arrived.assign(nw, 0); g_warp_gen.assign(nw, 0); g_warp_live.assign(nw, 0);
                for (int t = 0; t < nt; ++t) ++g_warp_live[t >> 5];
                g_xchg.assign(nt, 0);
                for (int t = 0; t < nt; ++t) {
                    fibers[t].stack.resize(stack_bytes);
                    getcontext(&fibers[t].ctx);
                    fibers[t].ctx.uc_stack.ss_sp = fibers[t].stack.data();
                    fibers[t].ctx.uc_stack.ss_size = stack_bytes;
                    fibers[t].ctx.uc_link = &g_sched;
                    makecontext(&fibers[t].ctx, trampoline, 0);
                }
                unsigned long idle_passes = 0, last_progress = ~0ul;
                while (g_live > 0) {
                    const unsigned long progress = g_block_gen * 1000003ul + (unsigned long)g_live;
                    unsigned long wsum = 0;
                    for (auto g : g_warp_gen) wsum += g;
                    const unsigned long state = progress ^ (wsum << 20);
                    idle_passes = (state == last_progress) ? idle_passes + 1 : 0;
                    last_progress = state;
                    if (idle_passes > 4) {
                        std::fprintf(stderr, "simt_emu: deadlock -- %d live threads, %d at the block barrier\n", g_live, g_block_arrived);
                        std::abort();
                    }
                    for (int t = 0; t < nt; ++t) {
                        if (fibers[t].done) continue;
                        g_current = t;
                        g_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        swapcontext(&g_sched, &fibers[t].ctx);
                    }
                }
            }
}
}  // namespace simt

#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
#define __syncthreads() simt::syncthreads()
#define __syncwarp() simt::syncwarp()
#define __shfl_sync(mask, v, src, width) simt::shfl((v), (src), (width))
#define __any_sync(mask, pred) simt::any(pred)
#define __noinline__
