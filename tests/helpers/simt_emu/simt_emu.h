// TEST-ONLY: a minimal SIMT emulator for running ONE-CTA-at-a-time CUDA kernels on the host, so that a kernel written without
// access to a GPU can be checked for indexing / algorithm errors before it costs GPU minutes.
//   * every CUDA thread of a block is a ucontext fiber; __syncthreads() yields until all fibers of the block have arrived;
//   * between barriers the fibers run one after the other in thread order, so data races that need a particular interleaving
//     are NOT detected -- this finds deterministic mistakes (wrong index, missing barrier that matters in thread order,
//     wrong formula), nothing more;
//   * shared memory: the kernel's `extern __shared__ ... raw[]` binds to the array the harness defines.
// Not supported: warp shuffles / votes, atomics, cp.async, tensor cores, clusters.
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
struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = false; };
inline dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
inline ucontext_t g_sched;
inline std::vector<Fiber> *g_fibers = nullptr;
inline int g_current = -1;
inline std::function<void()> *g_body = nullptr;
inline void trampoline() {
    (*g_body)();
    (*g_fibers)[g_current].done = true;
    swapcontext(&(*g_fibers)[g_current].ctx, &g_sched);
}
inline void syncthreads() { swapcontext(&(*g_fibers)[g_current].ctx, &g_sched); }

// Runs body() once per thread of every block of the grid; blocks one after the other.
inline void launch(dim3 grid, dim3 block, std::function<void()> body, size_t stack_bytes = 256 * 1024) {
    g_gridDim = grid; g_blockDim = block; g_body = &body;
    const int nt = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = dim3(bx, by, bz);
                std::vector<Fiber> fibers(nt);
                g_fibers = &fibers;
                for (int t = 0; t < nt; ++t) {
                    fibers[t].stack.resize(stack_bytes);
                    getcontext(&fibers[t].ctx);
                    fibers[t].ctx.uc_stack.ss_sp = fibers[t].stack.data();
                    fibers[t].ctx.uc_stack.ss_size = stack_bytes;
                    fibers[t].ctx.uc_link = &g_sched;
                    makecontext(&fibers[t].ctx, trampoline, 0);
                }
                int remaining = nt;
                while (remaining > 0) {          // one pass = every live fiber runs up to its next barrier (or to the end)
                    int finished_now = 0, live = 0;
                    for (int t = 0; t < nt; ++t) {
                        if (fibers[t].done) continue;
                        ++live;
                        g_current = t;
                        g_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        swapcontext(&g_sched, &fibers[t].ctx);
                        if (fibers[t].done) ++finished_now;
                    }
                    if (finished_now != 0 && finished_now != live) {
                        std::fprintf(stderr, "simt_emu: %d of %d threads left the kernel while the others wait at a barrier\n", finished_now, live);
                        std::abort();
                    }
                    remaining -= finished_now;
                }
            }
}
}  // namespace simt

#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
#define __syncthreads() simt::syncthreads()
