// TEST-ONLY: gptq_gguf_toolkit_b200/csrc/exact_update_v2.cuh (an experimental kernel that has not run on a GPU yet) on the SIMT
// emulator.  cp.async is emulated by an immediate 16-byte copy (so a missing wait would NOT be noticed; wrong indices are).
#define SIMT_EMU 1
#include "simt_emu.h"
#include <algorithm>
using std::min;
static inline void cp_async16(void *dst, const void *src) { std::memcpy(dst, src, 16); }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fsub_rn(float a, float b) { return a - b; }
alignas(16) unsigned char smem_raw[80 * 1024];       // the kernel's `extern __shared__ ... smem_raw[]`
#include "exact_update_v2.cuh"

extern "C" int run_exact_update_v2(float *W, const float *U, int d_row, int d_col, int c) {
    static_assert(upd2::SMEM_BYTES <= sizeof(smem_raw), "shared memory array too small");
    const int nwin = (d_col - c - 256) / 256;
    if (nwin <= 0) return 0;
    upd2::Params p{W, U, d_row, d_col};
    simt::launch(dim3(nwin, (d_row + upd2::R - 1) / upd2::R), dim3(upd2::NT2), [&]() { exact_update_v2_kernel(p, c); });
    return 0;
}
