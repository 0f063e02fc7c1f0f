// kquant.cuh -- device-side K-quant numerics: scale/min search, (de)quantise, GGUF bit-pack.
//
// Bit-exact restatement of the reference's CPU arithmetic (quant/gptq/src/quant_utils.py and
// packing_utils.py of IST-DASLab/gptq-gguf-toolkit); every rounding is spelled out with an _rn
// intrinsic so that neither -fmad nor instruction selection can change it.
#pragma once
#ifndef GQ_HOST_SHIM      // the CPU suite compiles this header for the host through a shim of the intrinsics (tests/helpers/host_shim)
#include "common.cuh"
#endif

// torch Tensor.sum(dim=1) over GS in {16,32} contiguous fp32 == 8 lane accumulators
// lane[l] = ((x[l]+x[l+8])+x[l+16])+x[l+24], folded in order starting from 0.
template <int GS, class F> __device__ __forceinline__ float kq_sum8(F f) {
    float s = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        float a = f(l);
#pragma unroll
        for (int k = l + 8; k < GS; k += 8) a = __fadd_rn(a, f(k));
        s = __fadd_rn(s, a);
    }
    return s;
}

// clamp(rint(v), lo, hi) without the conversion pipe (FRND issues at a quarter of the FADD rate and has four times its latency, and
// this sits on the dependent chain of the column steps): t = (v + 1.5*2^23) - 1.5*2^23 is round-to-nearest-even for |v| <= 2^22
// (the sum lands in [2^23, 2^24], where the spacing is 1 and the magic constant is even, so ties go to the even integer; the
// subtraction is exact), and for anything larger -- or infinite -- t stays beyond +-2^22 on v's side, so the clamp returns the same
// bound; NaN stays NaN and leaves fmaxf/fminf as lo either way.  Needs |lo|, |hi| <= 2^22.  The only difference from rintf is the
// sign of a zero result (+0.0 here where rintf(-0.3f) is -0.0f), which no later operation of the path turns into a different value.
__device__ __forceinline__ float kq_rint_clamp(float v, float lo, float hi) {
    const float magic = 12582912.0f;
    return clampf(__fsub_rn(__fadd_rn(v, magic), magic), lo, hi);
}

// The same, also returning the bit pattern of the clamped sum BEFORE the magic constant is taken off again: its low byte is the
// two's-complement code byte of the result (1.5 * 2^23 = 0x4B400000 has a zero low byte), so callers pack codes with one PRMT
// instead of F2I + shift + mask + or.
__device__ __forceinline__ float kq_rint_clamp_bits(float v, float lo, float hi, uint32_t &bits) {
    const float magic = 12582912.0f;
    const float t = clampf(__fadd_rn(v, magic), __fadd_rn(magic, lo), __fadd_rn(magic, hi));
    bits = __float_as_uint(t);
    return __fsub_rn(t, magic);
}

// float((uint8(L) ** 2) mod 256) for an integer-valued L in [0, MAXQ]  (quant_utils.py:246: the codes are uint8, so the square wraps).
// L*L is exact in fp32; up to MAXQ = 15 it never reaches 256, beyond that 256*floor(L*L/256) is taken off -- the floor by an
// addition of 2^23 rounded towards zero -- so the sum s_l2 needs no float <-> int conversions.
template <int MAXQ> __device__ __forceinline__ float kq_sq_u8(float L) {
    const float l2 = __fmul_rn(L, L);
    if constexpr (MAXQ * MAXQ < 256) {
        return l2;
    } else {
        const float f = __fsub_rn(__fadd_rz(__fmul_rn(l2, 0.00390625f), 8388608.0f), 8388608.0f);
        return __fmaf_rn(-256.0f, f, l2);
    }
}

// make_k_quants (quant_utils.py:199-274): asymmetric weighted least-squares search for one group.
// vmask/amask: bit i set when candidate i had D > eps / was accepted (see gq.h search_flags).
template <int GS, int MAXQ>
__device__ __forceinline__ void kq_search_asym(const float (&x)[GS], const SearchParams &sp, float &out_scale,
                                               float &out_zero, uint32_t &vmask, uint32_t &amask) {
    const float fmaxq = (float)MAXQ;
    const float sum_x2 = kq_sum8<GS>([&](int k) { return __fmul_rn(x[k], x[k]); });          // :203
    const float av_x = __fsqrt_rn(__fdiv_rn(sum_x2, (float)GS));                               // :204
    float w[GS];
    float mn = x[0], mx = x[0];
#pragma unroll
    for (int k = 0; k < GS; ++k) {
        w[k] = __fadd_rn(av_x, fabsf(x[k]));                                                   // :205
        mn = fminf(mn, x[k]);
        mx = fmaxf(mx, x[k]);
    }
    mn = fminf(mn, 0.0f);                                                                      // :210
    const bool isconst = (mx == mn);                                                           // :211
    const float sum_w = kq_sum8<GS>([&](int k) { return w[k]; });                              // :214
    const float sum_x = kq_sum8<GS>([&](int k) { return __fmul_rn(w[k], x[k]); });             // :215
    float scale = __fdiv_rn(__fsub_rn(mx, mn), fmaxq);                                         // :218
    if (isconst) scale = 0.0f;                                                                 // :219
    const float iscale = __frcp_rn(fmaxf(scale, GQ_EPS));                                      // :220
    float best_err = kq_sum8<GS>([&](int k) {                                                  // :223-232
        float q = kq_rint_clamp(__fmul_rn(__fsub_rn(x[k], mn), iscale), 0.0f, fmaxq);
        if (isconst) q = 0.0f;
        const float diff = __fsub_rn(__fadd_rn(__fmul_rn(scale, q), mn), x[k]);
        return __fmul_rn(w[k], __fmul_rn(diff, diff));
    });
    float xmin = mn;  // aliases best_min (:228): every accepted candidate moves the grid origin
    float best_scale = scale;
    if (sp.nstep >= 1) {
        for (int i = 0; i <= sp.nstep; ++i) {                                                  // :240
            // :241  python_scalar / tensor == reciprocal(tensor) * fp32(scalar)
            // :243  L = 0 for a constant group: all its x equal mn (finite), so a zero inverse scale gives (x - xmin) * 0 = +-0 -> +0
            const float is = isconst ? 0.0f : __fmul_rn(__frcp_rn(fmaxf(__fsub_rn(mx, xmin), GQ_EPS)), sp.num[i]);
            float L[GS];
#pragma unroll
            for (int k = 0; k < GS; ++k) L[k] = kq_rint_clamp(__fmul_rn(__fsub_rn(x[k], xmin), is), 0.0f, fmaxq);   // :242
            const float s_l = kq_sum8<GS>([&](int k) { return __fmul_rn(w[k], L[k]); });            // :245
            const float s_l2 = kq_sum8<GS>([&](int k) { return __fmul_rn(w[k], kq_sq_u8<MAXQ>(L[k])); });   // :246  uint8 ** 2 wraps mod 256
            const float s_xl = kq_sum8<GS>([&](int k) { return __fmul_rn(__fmul_rn(w[k], x[k]), L[k]); });  // :247
            const float D = __fsub_rn(__fmul_rn(sum_w, s_l2), __fmul_rn(s_l, s_l));                 // :249
            if (D > GQ_EPS) vmask |= (1u << i);                                                     // :250
            float sc = __fdiv_rn(__fsub_rn(__fmul_rn(sum_w, s_xl), __fmul_rn(sum_x, s_l)), D);      // :254
            float m2 = __fdiv_rn(__fsub_rn(__fmul_rn(s_l2, sum_x), __fmul_rn(s_l, s_xl)), D);       // :255
            if (m2 > 0.0f) {                                                                        // :257-260
                sc = __fdiv_rn(s_xl, fmaxf(s_l2, GQ_EPS));
                m2 = 0.0f;
            }
            const float cand = kq_sum8<GS>([&](int k) {                                             // :262-264
                const float diff = __fsub_rn(__fadd_rn(__fmul_rn(sc, L[k]), m2), x[k]);
                return __fmul_rn(w[k], __fmul_rn(diff, diff));
            });
            if (cand < best_err) {                                                                  // :266-270
                best_err = cand;
                best_scale = sc;
                xmin = m2;
                amask |= (1u << i);
            }
        }
    }
    out_scale = best_scale;
    out_zero = -xmin;                                                                               // :273
}

// make_quants (quant_utils.py:147-197), quant_scale == "absmax": symmetric scale, zero == 0.
template <int GS, int MAXQ>
__device__ __forceinline__ void kq_search_sym(const float (&x)[GS], float &out_scale, float &out_zero) {
    float mn = x[0], mx = x[0];
#pragma unroll
    for (int k = 1; k < GS; ++k) { mn = fminf(mn, x[k]); mx = fmaxf(mx, x[k]); }
    mx = fmaxf(fabsf(mn), mx);                          // :153
    if (mn < 0.0f) mn = -mx;                            // :154-156
    if (mn == mx) { mn = -1.0f; mx = 1.0f; }            // :157-159
    out_scale = __fdiv_rn(__fsub_rn(mx, mn), (float)MAXQ);  // :161
    out_zero = 0.0f;                                    // :195
}

template <int QT, int GS = Fmt<QT>::GS>
__device__ __forceinline__ void kq_group_search(const float (&x)[GS], const SearchParams &sp, float &scale,
                                                float &zero, uint32_t &vmask, uint32_t &amask) {
    constexpr int MAXQ = (1 << Fmt<QT>::BITS) - 1;
    if constexpr (Fmt<QT>::ASYM) kq_search_asym<GS, MAXQ>(x, sp, scale, zero, vmask, amask);
    else kq_search_sym<GS, MAXQ>(x, scale, zero);
}

// Super-block double quantisation of the group scales (quant_utils.py:117-143) for ONE row.
// gs/gz: the 256/GS group scales / zeros.  Writes fp16 bit patterns and integer codes.
template <int QT>
__device__ __forceinline__ void kq_row_finalize(const float *gs, const float *gz, uint16_t &d_bits,
                                                uint16_t &dmin_bits, uint8_t *sq, uint8_t *zq) {
    constexpr int GPR = GQ_QK_K / Fmt<QT>::GS;
    const float smq = (float)Fmt<QT>::SMQ;
    float ms = gs[0], mz = gz[0];
#pragma unroll
    for (int g = 1; g < GPR; ++g) { ms = fmaxf(ms, gs[g]); mz = fmaxf(mz, gz[g]); }          // :121
    d_bits = __half_as_ushort(__float2half_rn(__fdiv_rn(ms, smq)));                            // :124
    dmin_bits = __half_as_ushort(__float2half_rn(__fdiv_rn(mz, smq)));                         // :125
    const float inv_s = ms > 0.0f ? __fmul_rn(__frcp_rn(ms), smq) : 0.0f;                      // :128
    const float inv_z = mz > 0.0f ? __fmul_rn(__frcp_rn(mz), smq) : 0.0f;                      // :129
#pragma unroll
    for (int g = 0; g < GPR; ++g) {
        sq[g] = (uint8_t)(int)kq_rint_clamp(__fmul_rn(inv_s, gs[g]), 0.0f, smq);               // :132-137
        zq[g] = (uint8_t)(int)kq_rint_clamp(__fmul_rn(inv_z, gz[g]), 0.0f, smq);               // :138-143
    }
}

// quantize / dequantize (quant_utils.py:34-46).  s = d*sq, z = dmin*zq already formed (each one rounding).
__device__ __forceinline__ float kq_quant(float x, float s, float z, float lo, float hi) {
    return kq_rint_clamp(__fdiv_rn(__fadd_rn(x, z), fmaxf(s, GQ_EPS)), lo, hi);
}
__device__ __forceinline__ float kq_dequant(float q, float s, float z) {
    return __fsub_rn(__fmul_rn(s, q), z);
}

// ---------------------------------------------------------------------------------------------
// GGUF block bytes (packing_utils.py:8-326).  Pure function of the byte index b in [0, TS):
// q  : the 256 codes of the super-block as stored (u8, or i8 bit patterns for Q3_K/Q6_K)
// sq, zq: the group codes; d/dmin: fp16 bit patterns.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t kq_scale_min_byte(const uint8_t *sc, const uint8_t *mn, int j) {  // :8-30
    if (j < 4) return (uint8_t)(sc[j] | ((sc[4 + j] >> 4) << 6));
    if (j < 8) return (uint8_t)(mn[j - 4] | ((mn[j] >> 4) << 6));
    return (uint8_t)((sc[j - 4] & 0x0F) | ((mn[j - 4] & 0x0F) << 4));
}

template <int QT>
__device__ __forceinline__ uint8_t kq_pack_byte(int b, const uint8_t *q, const uint8_t *sq, const uint8_t *zq,
                                                uint16_t d, uint16_t dmin) {
    if constexpr (QT == GQ_Q2_K) {            // scales[16] qs[64] d dmin            :33-77
        if (b < 16) return (uint8_t)((sq[b] & 0x0F) | ((zq[b] & 0x0F) << 4));
        if (b < 80) {
            const int c = (b - 16) >> 5, l = (b - 16) & 31;
            const uint8_t *p = q + 128 * c + l;
            return (uint8_t)(p[0] | (p[32] << 2) | (p[64] << 4) | (p[96] << 6));
        }
        if (b < 82) return (uint8_t)(d >> (8 * (b - 80)));
        return (uint8_t)(dmin >> (8 * (b - 82)));
    } else if constexpr (QT == GQ_Q3_K) {     // hmask[32] qs[64] scales[12] d         :80-142
        if (b < 32) {
            uint8_t m = 0;
#pragma unroll
            for (int g = 0; g < 8; ++g) m |= (uint8_t)((((int)(int8_t)q[32 * g + b] + 4) > 3) << g);
            return m;
        }
        if (b < 96) {
            const int c = (b - 32) >> 5, l = (b - 32) & 31;
            const uint8_t *p = q + 128 * c + l;
            auto lo2 = [](uint8_t v) { return (uint8_t)(((int)(int8_t)v + 4) & 3); };
            return (uint8_t)(lo2(p[0]) | (lo2(p[32]) << 2) | (lo2(p[64]) << 4) | (lo2(p[96]) << 6));
        }
        if (b < 108) {
            const int j = b - 96;
            auto L = [&](int g) { return (uint8_t)((int)(int8_t)sq[g] + 32); };
            if (j < 8) return (uint8_t)((L(j) & 0x0F) | ((L(j + 8) & 0x0F) << 4));
            const int m = j - 8;
            return (uint8_t)(((L(m) >> 4) & 3) | (((L(m + 4) >> 4) & 3) << 2) | (((L(m + 8) >> 4) & 3) << 4) |
                             (((L(m + 12) >> 4) & 3) << 6));
        }
        return (uint8_t)(d >> (8 * (b - 108)));
    } else if constexpr (QT == GQ_Q4_K) {     // d dmin scales[12] qs[128]             :145-190
        if (b < 2) return (uint8_t)(d >> (8 * b));
        if (b < 4) return (uint8_t)(dmin >> (8 * (b - 2)));
        if (b < 16) return kq_scale_min_byte(sq, zq, b - 4);
        const int c = (b - 16) >> 5, l = (b - 16) & 31;
        return (uint8_t)(q[64 * c + l] | (q[64 * c + 32 + l] << 4));
    } else if constexpr (QT == GQ_Q5_K) {     // d dmin scales[12] qh[32] ql[128]      :193-262
        if (b < 2) return (uint8_t)(d >> (8 * b));
        if (b < 4) return (uint8_t)(dmin >> (8 * (b - 2)));
        if (b < 16) return kq_scale_min_byte(sq, zq, b - 4);
        if (b < 48) {
            const int l = b - 16;
            uint8_t m = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                m |= (uint8_t)(((q[64 * c + l] > 15) << (2 * c)) | ((q[64 * c + 32 + l] > 15) << (2 * c + 1)));
            return m;
        }
        const int c = (b - 48) >> 5, l = (b - 48) & 31;
        return (uint8_t)((q[64 * c + l] & 0x0F) | ((q[64 * c + 32 + l] & 0x0F) << 4));
    } else {                                  // Q6_K: ql[128] qh[64] scales[16] d      :265-326
        auto u = [&](int j) { return (uint8_t)((int)(int8_t)q[j] + 32); };
        if (b < 128) {
            const int c = b >> 6, r = b & 63;
            if (r < 32) return (uint8_t)((u(128 * c + r) & 0xF) | ((u(128 * c + 64 + r) & 0xF) << 4));
            const int l = r - 32;
            return (uint8_t)((u(128 * c + 32 + l) & 0xF) | ((u(128 * c + 96 + l) & 0xF) << 4));
        }
        if (b < 192) {
            const int c = (b - 128) >> 5, l = (b - 128) & 31;
            return (uint8_t)(((u(128 * c + l) >> 4) & 3) | (((u(128 * c + 32 + l) >> 4) & 3) << 2) |
                             (((u(128 * c + 64 + l) >> 4) & 3) << 4) | (((u(128 * c + 96 + l) >> 4) & 3) << 6));
        }
        if (b < 208) return sq[b - 192];
        return (uint8_t)(d >> (8 * (b - 208)));
    }
}

// float value of a stored code byte
template <int QT> __device__ __forceinline__ float kq_code_to_f(uint8_t b) {
    if constexpr (Fmt<QT>::ASYM) return (float)b;
    else return (float)(int8_t)b;
}
