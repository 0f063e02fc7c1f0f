// f32x2.cuh -- Blackwell packed-pair fp32 arithmetic (PTX *.f32x2, SASS FFMA2 / FMUL2 / FADD2) and an exact
// division by a pre-inverted divisor.
//
// sm_100 issues a scalar FFMA/FMUL/FADD warp instruction only every second cycle per SM sub-partition; the
// packed forms do two IEEE round-to-nearest operations per lane in the same slot, element-wise bit-identical
// to the scalar `_rn` intrinsics.  Everything here keeps the library's numerical contract (-fmad=false, every
// rounding explicit): f2_fma == two __fmaf_rn, f2_mul_nofuse == two __fmul_rn, f2_add/f2_sub == two __fadd_rn/__fsub_rn.
#pragma once
#ifndef SIMT_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

typedef unsigned long long f2_t;    // two fp32 in one 64-bit register pair: {lo, hi}

#ifdef SIMT_EMU      // the CPU suite runs the kernels on an emulator (tests/helpers/simt_emu): same element-wise IEEE operations
#define F2_NEG_ZERO2 0x8000000080000000ull
static inline f2_t f2_pack(float lo, float hi) { uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return (f2_t)a | ((f2_t)b << 32); }
static inline void f2_unpack(f2_t v, float &lo, float &hi) { uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32); memcpy(&lo, &a, 4); memcpy(&hi, &b, 4); }
static inline f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    float al, ah, bl, bh, cl, ch; f2_unpack(a, al, ah); f2_unpack(b, bl, bh); f2_unpack(c, cl, ch);
    return f2_pack(fmaf(al, bl, cl), fmaf(ah, bh, ch));
}
static inline f2_t f2_mul_nofuse(f2_t a, f2_t b, f2_t nz) { return f2_fma(a, b, nz); }
static inline f2_t f2_add(f2_t a, f2_t b) { float al, ah, bl, bh; f2_unpack(a, al, ah); f2_unpack(b, bl, bh); return f2_pack(al + bl, ah + bh); }
static inline f2_t f2_sub(f2_t a, f2_t b) { float al, ah, bl, bh; f2_unpack(a, al, ah); f2_unpack(b, bl, bh); return f2_pack(al - bl, ah - bh); }
#else
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// Two separately rounded products.  ptxas (12.9) contracts  mul.rn.f32x2 + add/sub.rn.f32x2  into one FFMA2 even under
// --fmad=false (the scalar forms are left alone), and it also folds fma(a, b, -0.0) back into a multiply first.  The
// product is therefore formed as fma(a, b, nz) with nz = {-0.0f, -0.0f} passed in at RUN TIME (kernel parameter), which
// ptxas cannot fold: a*b + (-0) == RN(a*b) bit for bit (including the sign of a zero product), and an FMA cannot be
// fused with the following add.  Same instruction count as FMUL2 + FADD2.
#define F2_NEG_ZERO2 0x8000000080000000ull
__device__ __forceinline__ f2_t f2_mul_nofuse(f2_t a, f2_t b, f2_t nz) {
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(nz));
    return d;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
    f2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) {
    f2_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

#endif  // SIMT_EMU

// ---------------------------------------------------------------------------------------------
// Exact a / b with the divisor's reciprocal hoisted out of a dependent chain.
//   y = RN(1/b) (correctly rounded, __frcp_rn);  q0 = RN(a*y);  r = a - q0*b (exact in one FMA);
//   q = RN(q0 + r*y)  is the correctly rounded quotient (Markstein) whenever nothing over/underflows and
//   b's significand is not all ones.  DivBy::make() checks the divisor once, div() checks the dividend;
//   anything outside the safe range takes the IEEE division, so the result is ALWAYS == __fdiv_rn(a, b).
// The dependent chain shrinks from ~55 cycles (MUFU.RCP + Newton + FCHK) to three FP ops.
// ---------------------------------------------------------------------------------------------
struct DivBy {
    float b, y;     // y == 0 marks "use the IEEE division"
    __device__ __forceinline__ static DivBy make(float b) {
        DivBy d;
        d.b = b;
        const float ab = fabsf(b);
        const bool safe = ab > 0x1p-60f && ab < 0x1p60f && (__float_as_uint(b) & 0x7FFFFFu) != 0x7FFFFFu;
        d.y = safe ? __frcp_rn(b) : 0.0f;
        return d;
    }
    __device__ __forceinline__ float div(float a) const {
        const float aa = fabsf(a);
        if (y != 0.0f && aa < 0x1p60f && aa > 0x1p-60f) {     // zeros (sign of zero!) take the IEEE path
            const float q0 = __fmul_rn(a, y);
            const float r = __fmaf_rn(-q0, b, a);
            return __fmaf_rn(r, y, q0);
        }
        return __fdiv_rn(a, b);
    }
    // Branch-free variant for latency-critical straight-line code: returns the Markstein quotient and ORs `bad` with
    // "this quotient is not guaranteed to equal a / b" (dividend out of the safe range, or unsafe divisor).  A zero
    // dividend is handled exactly (q0 = a*y already carries the IEEE sign of the zero).  The caller must redo the work
    // with div() when `bad` comes back set.
    __device__ __forceinline__ float div_fast(float a, bool &bad) const {
        const float aa = fabsf(a);
        const float q0 = __fmul_rn(a, y);
        const float r = __fmaf_rn(-q0, b, a);
        const float q = __fmaf_rn(r, y, q0);
        bad = bad || !(aa < 0x1p60f) || (aa <= 0x1p-60f && a != 0.0f) || y == 0.0f;
        return a == 0.0f ? q0 : q;
    }
};

// The same Markstein quotient for the dependent chain of the column steps, with the bookkeeping reduced to integer min / max
// accumulators (no predicates, no selects):  `ok()` afterwards says whether EVERY dividend seen was zero or inside
// (2^-60, 2^60) -- u2 = bits(a) << 1 drops the sign; a zero dividend gives u2 - 1 = 0xFFFFFFFF and never lowers the minimum.
// The divisor's own check (y != 0, DivBy::make) is the caller's, once per divisor instead of once per quotient.
// For a zero dividend the result is a zero whose SIGN may differ from IEEE's; the callers' next operation is either the magic-number
// rint of kq_rint_clamp_bits (which returns +0 for both) or a product that is subtracted from a weight (w - (+-0) == w; a zero
// weight can change its sign, and again meets the rint first), so no output value depends on it.
struct DivRange {
    uint32_t hi2 = 0u, lo2m1 = 0xFFFFFFFFu;
    __device__ __forceinline__ void see(float a) {
        const uint32_t u2 = __float_as_uint(a) << 1;
        hi2 = max(hi2, u2);
        lo2m1 = min(lo2m1, u2 - 1u);
    }
    __device__ __forceinline__ bool ok() const {       // all |a| < 2^60 and all non-zero |a| > 2^-60
        return hi2 < (0x5D800000u << 1) && lo2m1 >= (0x21800000u << 1);
    }
};
__device__ __forceinline__ float div_chain(float a, float b, float y, DivRange &rg) {
    rg.see(a);
    const float q0 = __fmul_rn(a, y);
    const float r = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(r, y, q0);
}
