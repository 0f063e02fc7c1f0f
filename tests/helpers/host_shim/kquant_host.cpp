// TEST-ONLY: the DEVICE source gptq_gguf_toolkit_b200/csrc/kquant.cuh (scale search, double quantisation of the scales,
// quantise / dequantise, GGUF bit-pack) compiled for the host through host_shim_intrinsics.h.
#define GQ_HOST_SHIM 1
#include "host_shim_intrinsics.h"
#include "kquant.cuh"

template <int QT>
static void rtn(const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep, uint8_t *qweight, uint16_t *d,
                uint16_t *dmin, uint8_t *sq, uint8_t *zq, uint8_t *packed, float *wdeq) {
    constexpr int GS = Fmt<QT>::GS, GPR = GQ_QK_K / GS, MAXQ = (1 << Fmt<QT>::BITS) - 1, TS = Fmt<QT>::TS;
    SearchParams sp;
    sp.nstep = nstep;
    for (int i = 0; i <= nstep && i < 64; ++i) sp.num[i] = (float)(rmin + rdelta * (double)i + (double)MAXQ);
    const int nsb = d_col / GQ_QK_K, ng = d_col / GS;
    const float lo = (float)Fmt<QT>::QMIN, hi = (float)Fmt<QT>::QMAX;
    for (int r = 0; r < d_row; ++r)
        for (int sb = 0; sb < nsb; ++sb) {
            float gs[16], gz[16];
            uint32_t vmask = 0, amask = 0;
            for (int g = 0; g < GPR; ++g) {
                float x[GS];
                for (int k = 0; k < GS; ++k) x[k] = W[(long)r * d_col + sb * GQ_QK_K + g * GS + k];
                kq_group_search<QT>(x, sp, gs[g], gz[g], vmask, amask);
            }
            uint16_t db, dmb;
            uint8_t *sqr = sq + (long)r * ng + sb * GPR, *zqr = zq + (long)r * ng + sb * GPR;
            kq_row_finalize<QT>(gs, gz, db, dmb, sqr, zqr);
            d[(long)r * nsb + sb] = db;
            dmin[(long)r * nsb + sb] = dmb;
            const float df = __half2float(__ushort_as_half(db)), dmf = __half2float(__ushort_as_half(dmb));
            uint8_t *codes = qweight + (long)r * d_col + sb * GQ_QK_K;
            for (int col = 0; col < GQ_QK_K; ++col) {
                const int g = col / GS;
                const float s = __fmul_rn(df, kq_code_to_f<QT>(sqr[g])), z = __fmul_rn(dmf, kq_code_to_f<QT>(zqr[g]));
                const float q = kq_quant(W[(long)r * d_col + sb * GQ_QK_K + col], s, z, lo, hi);
                codes[col] = (uint8_t)(int8_t)(int)q;
                if (wdeq) wdeq[(long)r * d_col + sb * GQ_QK_K + col] = kq_dequant(q, s, z);
            }
            if (packed)
                for (int b = 0; b < TS; ++b) packed[((long)r * nsb + sb) * TS + b] = kq_pack_byte<QT>(b, codes, sqr, zqr, db, dmb);
        }
}

extern "C" int host_rtn(int qtype, const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep, uint8_t *qweight,
                        uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq, uint8_t *packed, float *wdeq) {
    switch (qtype) {
    case GQ_Q2_K: rtn<GQ_Q2_K>(W, d_row, d_col, rmin, rdelta, nstep, qweight, d, dmin, sq, zq, packed, wdeq); return 0;
    case GQ_Q3_K: rtn<GQ_Q3_K>(W, d_row, d_col, rmin, rdelta, nstep, qweight, d, dmin, sq, zq, packed, wdeq); return 0;
    case GQ_Q4_K: rtn<GQ_Q4_K>(W, d_row, d_col, rmin, rdelta, nstep, qweight, d, dmin, sq, zq, packed, wdeq); return 0;
    case GQ_Q5_K: rtn<GQ_Q5_K>(W, d_row, d_col, rmin, rdelta, nstep, qweight, d, dmin, sq, zq, packed, wdeq); return 0;
    case GQ_Q6_K: rtn<GQ_Q6_K>(W, d_row, d_col, rmin, rdelta, nstep, qweight, d, dmin, sq, zq, packed, wdeq); return 0;
    }
    return -1;
}

// ---- direct checks of the conversion-free primitives (tests/test_kquant_source_cpu.py) ----
// kq_rint_clamp(v, lo, hi) against clampf(rintf(v), lo, hi) on a caller-supplied array of raw float bit patterns; returns the
// number of inputs whose results differ as VALUES (+0.0 == -0.0; NaN results cannot occur: the clamp maps NaN to lo in both).
extern "C" long host_check_rint_clamp(const uint32_t *bits, long n, float lo, float hi) {
    long bad = 0;
    for (long i = 0; i < n; ++i) {
        float v;
        std::memcpy(&v, &bits[i], 4);
        const float a = kq_rint_clamp(v, lo, hi), b = clampf(rintf(v), lo, hi);
        if (!(a == b)) ++bad;
    }
    return bad;
}
// the same for EVERY bit pattern in [first, last], both signs
extern "C" long host_check_rint_clamp_range(uint32_t first, uint32_t last, float lo, float hi) {
    long bad = 0;
    for (uint64_t u = first; u <= last; ++u)
        for (uint32_t sign = 0; sign < 2; ++sign) {
            const uint32_t b32 = (uint32_t)u | (sign << 31);
            float v;
            std::memcpy(&v, &b32, 4);
            const float a = kq_rint_clamp(v, lo, hi), b = clampf(rintf(v), lo, hi);
            if (!(a == b)) ++bad;
        }
    return bad;
}
// kq_sq_u8<MAXQ>(L) against float((uint8(L) ** 2) mod 256) for every integer L in [0, MAXQ]
extern "C" int host_check_sq_u8() {
    int bad = 0;
    for (int l = 0; l <= 3; ++l) bad += kq_sq_u8<3>((float)l) != (float)((l * l) & 255);
    for (int l = 0; l <= 15; ++l) bad += kq_sq_u8<15>((float)l) != (float)((l * l) & 255);
    for (int l = 0; l <= 31; ++l) bad += kq_sq_u8<31>((float)l) != (float)((l * l) & 255);
    for (int l = 0; l <= 255; ++l) bad += kq_sq_u8<255>((float)l) != (float)((l * l) & 255);
    return bad;
}
