// TEST-ONLY: gptq_gguf_toolkit_b200/csrc/chol_diag_v2.cuh (the SHIPPED diagonal-block kernel of gq_prepare) on the SIMT emulator.
#define SIMT_EMU 1
#include "simt_emu.h"
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
namespace {
constexpr int NB = 128;
alignas(16) uint8_t raw[160 * 1024];          // the kernel's `extern __shared__ ... raw[]`
#include "chol_diag_v2.cuh"
}  // namespace

extern "C" int run_chol_diag_v2(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    static_assert(sizeof(DiagSmem2) <= sizeof(raw), "shared memory array too small");
    simt::launch(dim3(1), dim3(DT2), [&]() { chol_diag_v2_kernel(A, Binv, BinvT, ld, k0, not_pd); });
    return 0;
}
