// TEST-ONLY host shim: lets g++ compile gptq_gguf_toolkit_b200/csrc/kquant_bf16.cuh (device code) for the CPU, so that the
// arithmetic of the experimental bf16 scale search can be checked against the reference golden without a GPU
// (tests/test_kquant_bf16_source_cpu.py).  It stands in for kquant.cuh / common.cuh with just what that header uses; every CUDA
// intrinsic maps to the IEEE operation it denotes (compile with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define GQ_EPS 1e-9f
#define GQ_QK_K 256

struct SearchParams { int nstep; float num[64]; };

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

struct __nv_bfloat16 { uint16_t bits; };
static inline __nv_bfloat16 __float2bfloat16_rn(float x) {
    uint32_t u; std::memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) <= 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
    return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline float __bfloat162float(__nv_bfloat16 b) { uint32_t u = (uint32_t)b.bits << 16; float r; std::memcpy(&r, &u, 4); return r; }
struct __half { _Float16 v; };
static inline __half __float2half_rn(float x) { return __half{(_Float16)x}; }
static inline uint16_t __half_as_ushort(__half h) { uint16_t b; std::memcpy(&b, &h.v, 2); return b; }

// torch Tensor.sum(dim=1) over GS in {16,32}: 8 lane accumulators folded in order (same as kquant.cuh)
template <int GS, class F> inline float kq_sum8(F f) {
    float s = 0.0f;
    for (int l = 0; l < 8; ++l) {
        float a = f(l);
        for (int k = l + 8; k < GS; k += 8) a = __fadd_rn(a, f(k));
        s = __fadd_rn(s, a);
    }
    return s;
}

// the format constants kquant_bf16.cuh reads (GGML_QUANT_SIZES, quant_utils.py:19-26)
enum { GQ_Q2_K = 10, GQ_Q3_K = 11, GQ_Q4_K = 12, GQ_Q5_K = 13, GQ_Q6_K = 14 };
template <int QT> struct Fmt;
template <> struct Fmt<GQ_Q2_K> { static constexpr int BITS = 2, GS = 16, SMQ = 15; static constexpr bool ASYM = true; };
template <> struct Fmt<GQ_Q3_K> { static constexpr int BITS = 3, GS = 16, SMQ = 31; static constexpr bool ASYM = false; };
template <> struct Fmt<GQ_Q4_K> { static constexpr int BITS = 4, GS = 32, SMQ = 63; static constexpr bool ASYM = true; };
template <> struct Fmt<GQ_Q5_K> { static constexpr int BITS = 5, GS = 32, SMQ = 63; static constexpr bool ASYM = true; };
template <> struct Fmt<GQ_Q6_K> { static constexpr int BITS = 6, GS = 16, SMQ = 63; static constexpr bool ASYM = false; };
