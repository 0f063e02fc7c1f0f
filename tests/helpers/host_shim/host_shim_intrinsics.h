// TEST-ONLY host shim: lets g++ compile the device headers gptq_gguf_toolkit_b200/csrc/kquant.cuh and kquant_bf16.cuh for the CPU,
// so that the K-quant arithmetic of the product source is checked against the reference goldens without a GPU
// (tests/test_kquant_source_cpu.py, tests/test_kquant_bf16_source_cpu.py).  It stands in for common.cuh with just what they use; every CUDA
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

static inline uint32_t __float_as_uint(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
// a + b rounded towards zero: the sum of two floats is exact in double whenever their exponents are less than ~29 apart (all
// uses: a small non-negative value plus 2^23); the double is then truncated to float
static inline float __fadd_rz(float a, float b) {
    const double d = (double)a + (double)b;
    float r = (float)d;
    if (std::fabs((double)r) > std::fabs(d)) r = std::nextafterf(r, 0.0f);
    return r;
}
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
// ---- what tile.cuh / rtn_native.cuh need on top (used together with tests/helpers/simt_emu/simt_emu.h) ----
enum { GQ_F32 = 0, GQ_F16 = 1, GQ_BF16 = 2 };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
struct __half2 { __half x, y; };
static inline __half2 __floats2half2_rn(float a, float b) { return __half2{__float2half_rn(a), __float2half_rn(b)}; }
static inline uint32_t __reduce_or_sync(uint32_t, uint32_t v) { return v; }      // publish_flags is not exercised on the host
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { uint32_t o = *p; *p |= v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }

static inline __half __ushort_as_half(uint16_t b) { __half h; std::memcpy(&h.v, &b, 2); return h; }
static inline float __half2float(__half h) { return (float)h.v; }

// the format constants kquant_bf16.cuh reads (GGML_QUANT_SIZES, quant_utils.py:19-26)
enum { GQ_Q2_K = 10, GQ_Q3_K = 11, GQ_Q4_K = 12, GQ_Q5_K = 13, GQ_Q6_K = 14 };
template <int QT> struct Fmt;
// (copied from csrc/common.cuh; tests/test_kquant_source_cpu.py::test_shim_format_table_matches_common_cuh keeps them in step)
template <> struct Fmt<GQ_Q2_K> { static constexpr int BITS = 2, QMIN = 0, QMAX = 3, SMQ = 15, GS = 16, ASYM = 1, TS = 84; };
template <> struct Fmt<GQ_Q3_K> { static constexpr int BITS = 3, QMIN = -4, QMAX = 3, SMQ = 31, GS = 16, ASYM = 0, TS = 110; };
template <> struct Fmt<GQ_Q4_K> { static constexpr int BITS = 4, QMIN = 0, QMAX = 15, SMQ = 63, GS = 32, ASYM = 1, TS = 144; };
template <> struct Fmt<GQ_Q5_K> { static constexpr int BITS = 5, QMIN = 0, QMAX = 31, SMQ = 63, GS = 32, ASYM = 1, TS = 176; };
template <> struct Fmt<GQ_Q6_K> { static constexpr int BITS = 6, QMIN = -32, QMAX = 31, SMQ = 63, GS = 16, ASYM = 0, TS = 210; };
static inline float load_as_f32(const void *p, long idx, int dtype) {       // common.cuh:85-89
    if (dtype == GQ_F32) return ((const float *)p)[idx];
    if (dtype == GQ_F16) return __half2float(((const __half *)p)[idx]);
    return __bfloat162float(((const __nv_bfloat16 *)p)[idx]);
}
