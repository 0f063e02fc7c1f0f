// TEST-ONLY: gptq_gguf_toolkit_b200/csrc/chol_diag_v4.cuh (the two-level diagonal-block kernel of gq_prepare) on the SIMT emulator.
#define SIMT_EMU 1
#include "simt_emu.h"
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
namespace cd4 { alignas(16) unsigned char raw[160 * 1024]; }   // the kernel's `extern __shared__ ... raw[]` (declared inside namespace cd4)
#include "chol_diag_v4.cuh"
using namespace cd4;

// A: (n x n) row-major with the 128 x 128 SPD block at (k0, k0); outputs like the kernel's.
extern "C" int run_chol_diag_v4(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    static_assert(sizeof(Smem4) <= sizeof(raw), "shared memory array too small");
    simt::launch(dim3(1), dim3(T4), [&]() { chol_diag_v4_kernel(A, Binv, BinvT, ld, k0, not_pd); });
    return 0;
}
