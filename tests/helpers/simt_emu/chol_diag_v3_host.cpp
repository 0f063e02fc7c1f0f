// TEST-ONLY: gptq_gguf_toolkit_b200/csrc/chol_diag_v3.cuh (a kernel that has not run on a GPU yet) on the SIMT emulator.
#define SIMT_EMU 1
#include "simt_emu.h"
namespace cd3 { alignas(16) unsigned char raw[160 * 1024]; }   // the kernel's `extern __shared__ ... raw[]` (declared inside namespace cd3)
#include "chol_diag_v3.cuh"
using namespace cd3;

// A: (n x n) row-major with the 128 x 128 SPD block at (k0, k0); outputs like the kernel's.
extern "C" int run_chol_diag_v3(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    static_assert(sizeof(Smem3) <= sizeof(raw), "shared memory array too small");
    simt::launch(dim3(1), dim3(T3), [&]() { chol_diag_v3_kernel(A, Binv, BinvT, ld, k0, not_pd); });
    return 0;
}
