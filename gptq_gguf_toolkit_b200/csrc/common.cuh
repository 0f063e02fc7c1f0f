// common.cuh -- shared helpers for libgq (sm_100a only).
//
// The whole library is compiled with -fmad=false: a*b+c is NEVER contracted.  Every FMA in
// the code base is an explicit fmaf()/__fmaf_rn().  Division, sqrt and reciprocal are the IEEE
// correctly rounded versions (nvcc defaults -prec-div=true -prec-sqrt=true -ftz=false).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gq.h"

#define GQ_QK_K 256
#define GQ_EPS 1e-9f

// ---------------------------------------------------------------------------------------------
// Format registry (reference: quant_utils.py:19-26 GGML_QUANT_SIZES; gguf type sizes)
// ---------------------------------------------------------------------------------------------
template <int QT> struct Fmt;
template <> struct Fmt<GQ_Q2_K> { static constexpr int BITS = 2, QMIN = 0, QMAX = 3, SMQ = 15, GS = 16, ASYM = 1, TS = 84; };
template <> struct Fmt<GQ_Q3_K> { static constexpr int BITS = 3, QMIN = -4, QMAX = 3, SMQ = 31, GS = 16, ASYM = 0, TS = 110; };
template <> struct Fmt<GQ_Q4_K> { static constexpr int BITS = 4, QMIN = 0, QMAX = 15, SMQ = 63, GS = 32, ASYM = 1, TS = 144; };
template <> struct Fmt<GQ_Q5_K> { static constexpr int BITS = 5, QMIN = 0, QMAX = 31, SMQ = 63, GS = 32, ASYM = 1, TS = 176; };
template <> struct Fmt<GQ_Q6_K> { static constexpr int BITS = 6, QMIN = -32, QMAX = 31, SMQ = 63, GS = 16, ASYM = 0, TS = 210; };

struct FmtInfo { int bits, qmin, qmax, smq, gs, asym, ts; };
inline bool gq_fmt_info(int qt, FmtInfo &f) {
    switch (qt) {
    case GQ_Q2_K: f = {2, 0, 3, 15, 16, 1, 84}; return true;
    case GQ_Q3_K: f = {3, -4, 3, 31, 16, 0, 110}; return true;
    case GQ_Q4_K: f = {4, 0, 15, 63, 32, 1, 144}; return true;
    case GQ_Q5_K: f = {5, 0, 31, 63, 32, 1, 176}; return true;
    case GQ_Q6_K: f = {6, -32, 31, 63, 16, 0, 210}; return true;
    }
    return false;
}

// Search parameters shared by all kernels.  num[i] = fp32(rmin + rdelta*i + maxq) evaluated in
// double on the host exactly like the Python expression at quant_utils.py:241.
struct SearchParams {
    int nstep;      // candidates are i = 0..nstep (none if nstep < 1)
    float num[64];
};

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, never throws)
// ---------------------------------------------------------------------------------------------
void gq_set_error(const char *fmt, ...);
void gq_count_launches(int n);   // bookkeeping behind gq_launch_count()
#define GQ_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            gq_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GQ_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)
#define GQ_REQUIRE(cond, ...)           \
    do {                                \
        if (!(cond)) {                  \
            gq_set_error(__VA_ARGS__);  \
            return GQ_ERR_INVALID;      \
        }                               \
    } while (0)

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// cp.async (LDGSTS) 16-byte copy, L2-only caching.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// load a scalar of dtype as fp32 (exact widening)
__device__ __forceinline__ float load_as_f32(const void *p, long idx, int dtype) {
    if (dtype == GQ_F32) return ((const float *)p)[idx];
    if (dtype == GQ_F16) return __half2float(((const __half *)p)[idx]);
    return __bfloat162float(((const __nv_bfloat16 *)p)[idx]);
}
// store fp32 as dtype (round to nearest even), == torch .to(dtype)
__device__ __forceinline__ void store_from_f32(void *p, long idx, int dtype, float v) {
    if (dtype == GQ_F32) ((float *)p)[idx] = v;
    else if (dtype == GQ_F16) ((__half *)p)[idx] = __float2half_rn(v);
    else ((__nv_bfloat16 *)p)[idx] = __float2bfloat16_rn(v);
}
