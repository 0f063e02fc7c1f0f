// kquant_bf16.cuh -- the K-quant scale search in the reference's BF16 (or FP16) arithmetic (written at the end of round 1, validated on B200 in round 2:
// after the GPU budget was spent; the CPU restatement of the same arithmetic that the tests check against is pinned bit for
// bit to the reference -- tests/golden/rtn_bf16.npz -- and is what this code has to match on the first GPU run of round 2).
//
// Why: for embed_tokens / lm_head the reference calls get_scale_and_zero on the weight in its ORIGINAL dtype
// (quant/gptq/src/quantizer.py:303-305).  For a bf16 model every torch op of make_k_quants / make_quants /
// get_scale_and_zero (quant_utils.py:90-274) therefore computes in fp32 and rounds its result to bf16 (RNE), reductions
// accumulate in fp32 and round once, and `python_scalar / tensor` is reciprocal(tensor) ROUNDED, times the fp32 scalar,
// rounded again.  The fp32-arithmetic search of kquant.cuh gives different scales on ~10 % of the Q4_K groups.
// This header deliberately DUPLICATES the search of kquant.cuh instead of parametrising it: the fp32 path is the validated
// contract of the GPTQ layers and must not be touched by an unvalidated variant.
#pragma once
#ifndef GQ_HOST_SHIM      // tests compile this header for the host through a shim of the intrinsics (tests/helpers/host_shim)
#include "kquant.cuh"
#endif

// RND: the weight's dtype (GQ_BF16 / GQ_F16 codes of gq.h: 2 = bf16, 1 = fp16); every op of the search rounds to it.
// The fp16 rule set is the same one (pinned against the reference on CPU as well: tests/golden/rtn_f16.npz).
#define GQ_RND_BF16 2
#define GQ_RND_F16 1
template <int RND> __device__ __forceinline__ float rbn(float x) {
    if constexpr (RND == GQ_RND_BF16) return __bfloat162float(__float2bfloat16_rn(x));
    else return __half2float(__float2half_rn(x));
}
#define rb16(x) rbn<RND>(x)

template <int GS, int RND, class F> __device__ __forceinline__ float kqb_sum_(F f) {      // fp32 accumulation, ONE rounding at the end
    return rbn<RND>(kq_sum8<GS>(f));
}
#define kqb_sum kqb_sum_

template <int GS, int MAXQ, int RND>
__device__ __forceinline__ void kqb_search_asym(const float (&x)[GS], const SearchParams &sp, float &out_scale, float &out_zero) {
    const float fmaxq = (float)MAXQ;
    const float eps_t = rb16(GQ_EPS);                                                           // clamp_min(eps) on a bf16 tensor
    const float sum_x2 = kqb_sum<GS, RND>([&](int k) { return rb16(__fmul_rn(x[k], x[k])); });     // :203
    const float av_x = rb16(__fsqrt_rn(rb16(__fdiv_rn(sum_x2, (float)GS))));                    // :204
    float w[GS];
    float mn = x[0], mx = x[0];
#pragma unroll
    for (int k = 0; k < GS; ++k) {
        w[k] = rb16(__fadd_rn(av_x, fabsf(x[k])));                                              // :205
        mn = fminf(mn, x[k]);
        mx = fmaxf(mx, x[k]);
    }
    mn = fminf(mn, 0.0f);                                                                       // :210
    const bool isconst = (mx == mn);                                                            // :211
    const float sum_w = kqb_sum<GS, RND>([&](int k) { return w[k]; });                               // :214
    const float sum_x = kqb_sum<GS, RND>([&](int k) { return rb16(__fmul_rn(w[k], x[k])); });      // :215
    float scale = rb16(__fdiv_rn(rb16(__fsub_rn(mx, mn)), fmaxq));                              // :218
    if (isconst) scale = 0.0f;                                                                  // :219
    const float iscale = rb16(__frcp_rn(fmaxf(scale, eps_t)));                                  // :220
    float best_err = kqb_sum<GS, RND>([&](int k) {                                                   // :223-232
        float q = kq_rint_clamp(rb16(__fmul_rn(rb16(__fsub_rn(x[k], mn)), iscale)), 0.0f, fmaxq);
        if (isconst) q = 0.0f;
        const float diff = rb16(__fsub_rn(rb16(__fadd_rn(rb16(__fmul_rn(scale, q)), mn)), x[k]));
        return rb16(__fmul_rn(w[k], rb16(__fmul_rn(diff, diff))));
    });
    float xmin = mn;      // aliases best_min (:228)
    float best_scale = scale;
    if (sp.nstep >= 1) {
        for (int i = 0; i <= sp.nstep; ++i) {                                                   // :240
            // :241  python_scalar / bf16 tensor == round(round(reciprocal(tensor)) * fp32(scalar))
            // (:243  L = 0 for a constant group through a zero inverse scale, as in kquant.cuh)
            const float is = isconst ? 0.0f : rb16(__fmul_rn(rb16(__frcp_rn(fmaxf(rb16(__fsub_rn(mx, xmin)), eps_t))), sp.num[i]));
            float L[GS];
#pragma unroll
            for (int k = 0; k < GS; ++k) L[k] = kq_rint_clamp(rb16(__fmul_rn(rb16(__fsub_rn(x[k], xmin)), is)), 0.0f, fmaxq);   // :242
            const float s_l = kqb_sum<GS, RND>([&](int k) { return rb16(__fmul_rn(w[k], L[k])); });                  // :245
            const float s_l2 = kqb_sum<GS, RND>([&](int k) { return rb16(__fmul_rn(w[k], kq_sq_u8<MAXQ>(L[k]))); });   // :246  uint8 ** 2 wraps mod 256
            const float s_xl = kqb_sum<GS, RND>([&](int k) { return rb16(__fmul_rn(rb16(__fmul_rn(w[k], x[k])), L[k])); });   // :247
            const float D = rb16(__fsub_rn(rb16(__fmul_rn(sum_w, s_l2)), rb16(__fmul_rn(s_l, s_l))));                  // :249
            // (no whole-call skip of a candidate, like the fp32 search: see gq.h search_flags)
            float sc = rb16(__fdiv_rn(rb16(__fsub_rn(rb16(__fmul_rn(sum_w, s_xl)), rb16(__fmul_rn(sum_x, s_l)))), D));  // :254
            float m2 = rb16(__fdiv_rn(rb16(__fsub_rn(rb16(__fmul_rn(s_l2, sum_x)), rb16(__fmul_rn(s_l, s_xl)))), D));   // :255
            if (m2 > 0.0f) {                                                                                             // :257-260
                sc = rb16(__fdiv_rn(s_xl, fmaxf(s_l2, eps_t)));
                m2 = 0.0f;
            }
            const float cand = kqb_sum<GS, RND>([&](int k) {                                                                  // :262-264
                const float diff = rb16(__fsub_rn(rb16(__fadd_rn(rb16(__fmul_rn(sc, L[k])), m2)), x[k]));
                return rb16(__fmul_rn(w[k], rb16(__fmul_rn(diff, diff))));
            });
            if (cand < best_err) {                                                                                       // :266-270
                best_err = cand;
                best_scale = sc;
                xmin = m2;
            }
        }
    }
    out_scale = best_scale;
    out_zero = -xmin;                                                                                                    // :273
}

template <int GS, int MAXQ, int RND>
__device__ __forceinline__ void kqb_search_sym(const float (&x)[GS], float &out_scale, float &out_zero) {
    float mn = x[0], mx = x[0];
#pragma unroll
    for (int k = 1; k < GS; ++k) { mn = fminf(mn, x[k]); mx = fmaxf(mx, x[k]); }
    mx = fmaxf(fabsf(mn), mx);                          // :153
    if (mn < 0.0f) mn = -mx;                            // :154-156
    if (mn == mx) { mn = -1.0f; mx = 1.0f; }            // :157-159
    out_scale = rb16(__fdiv_rn(rb16(__fsub_rn(mx, mn)), (float)MAXQ));  // :161
    out_zero = 0.0f;                                    // :195
}

// Super-block double quantisation of the group scales (quant_utils.py:117-143) for ONE row, bf16 arithmetic.
template <int QT, int RND>
__device__ __forceinline__ void kqb_row_finalize(const float *gs, const float *gz, uint16_t &d_bits, uint16_t &dmin_bits,
                                                 uint8_t *sq, uint8_t *zq) {
    constexpr int GPR = GQ_QK_K / Fmt<QT>::GS;
    const float smq = (float)Fmt<QT>::SMQ;
    float ms = gs[0], mz = gz[0];
#pragma unroll
    for (int g = 1; g < GPR; ++g) { ms = fmaxf(ms, gs[g]); mz = fmaxf(mz, gz[g]); }             // :121
    d_bits = __half_as_ushort(__float2half_rn(rb16(__fdiv_rn(ms, smq))));                         // :124 (bf16 -> fp16)
    dmin_bits = __half_as_ushort(__float2half_rn(rb16(__fdiv_rn(mz, smq))));                      // :125
    const float inv_s = ms > 0.0f ? rb16(__fmul_rn(rb16(__frcp_rn(ms)), smq)) : 0.0f;             // :128
    const float inv_z = mz > 0.0f ? rb16(__fmul_rn(rb16(__frcp_rn(mz)), smq)) : 0.0f;             // :129
#pragma unroll
    for (int g = 0; g < GPR; ++g) {
        sq[g] = (uint8_t)(int)kq_rint_clamp(rb16(__fmul_rn(inv_s, gs[g])), 0.0f, smq);            // :132-137
        zq[g] = (uint8_t)(int)kq_rint_clamp(rb16(__fmul_rn(inv_z, gz[g])), 0.0f, smq);            // :138-143
    }
}

#undef rb16
#undef kqb_sum
