/*
 * gq_oracle.c -- CPU restatement of the reference GPTQ -> GGUF K-quant hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may include, link or
 * call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg use it, and there only as the checker / the CPU arm.
 *
 * Parity pin: the reference (IST-DASLab/gptq-gguf-toolkit @7a38bc5) ships no tests and no
 * golden vectors for this path, so the oracle is pinned against outputs of the reference
 * itself run on CPU in the build container (tests/golden/make_golden.py imports
 * /root/reference/quant/gptq/src and commits the vectors under tests/golden/).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * quant/gptq/src/ of the reference).  All arithmetic is IEEE binary32, round to nearest
 * even, NO fused multiply-add except where the reference's BLAS uses one (the rank-k
 * update); compile with -ffp-contract=off.
 *
 * Order-of-operations facts this file relies on (measured against torch 2.11 CPU):
 *   - Tensor.sum(dim=1) over 16/32 fp32 values = 8 lane accumulators + ordered fold.
 *   - python_scalar / tensor  = reciprocal(tensor) * fp32(scalar)  (two roundings).
 *   - tensor / python_scalar, tensor / tensor = true division.
 *   - uint8_tensor ** 2 wraps mod 256.
 *   - addr_ with a strided vec2 = no FMA;  addmm_ (K=128) = single-accumulator FMA chain.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QK_K 256
#define ORC_EPS 1e-9f

/* ------------------------------------------------------------------------------------------
 * Format registry: quant_utils.py:19-26 (GGML_QUANT_SIZES) and gguf type sizes.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int bits, qmin, qmax, scale_maxq, group_size, asym, type_size;
} orc_fmt_t;

static int orc_fmt(int qtype, orc_fmt_t *f) {
    switch (qtype) {
    case 10: *f = (orc_fmt_t){2, 0, 3, 15, 16, 1, 84}; return 0;    /* Q2_K */
    case 11: *f = (orc_fmt_t){3, -4, 3, 31, 16, 0, 110}; return 0;  /* Q3_K */
    case 12: *f = (orc_fmt_t){4, 0, 15, 63, 32, 1, 144}; return 0;  /* Q4_K */
    case 13: *f = (orc_fmt_t){5, 0, 31, 63, 32, 1, 176}; return 0;  /* Q5_K */
    case 14: *f = (orc_fmt_t){6, -32, 31, 63, 16, 0, 210}; return 0; /* Q6_K */
    }
    return -1;
}

int orc_format(int qtype, int *out7) {
    orc_fmt_t f;
    if (orc_fmt(qtype, &f)) return -1;
    out7[0] = f.bits; out7[1] = f.qmin; out7[2] = f.qmax; out7[3] = f.scale_maxq;
    out7[4] = f.group_size; out7[5] = f.asym; out7[6] = f.type_size;
    return 0;
}

/* fp32 -> fp16 (RN-even) -> bits, and back */
static inline uint16_t f2h(float x) { _Float16 h = (_Float16)x; uint16_t b; memcpy(&b, &h, 2); return b; }
static inline float h2f(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return (float)h; }

/* torch sum(dim=1) over n in {16,32} contiguous fp32: 8 lanes, then ordered fold. */
static inline float sum8(const float *v, int n) {
    float lane[8];
    for (int l = 0; l < 8; ++l) {
        float a = v[l];
        for (int k = l + 8; k < n; k += 8) a = a + v[k];
        lane[l] = a;
    }
    float s = 0.0f;
    for (int l = 0; l < 8; ++l) s = s + lane[l];
    return s;
}

static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* bf16-arithmetic mode of the scale search (orc_rtn_quantize_bf16): the reference runs get_scale_and_zero on the weight in
 * its ORIGINAL dtype (quantizer.py:303-305), so for a bf16 model every torch op of make_k_quants / make_quants /
 * get_scale_and_zero computes in fp32 and rounds its result to bf16 (round to nearest even), reductions accumulate in fp32 and
 * round once, and `python_scalar / tensor` is reciprocal(tensor) ROUNDED, then times the fp32 scalar, rounded again.
 * Pinned op by op and end to end against the reference on CPU (tests/golden/make_golden_rtn_bf16.py): all four scale
 * tensors bit-identical for the five types.  RB() is the identity in the fp32 mode, i.e. the fp32 contract is untouched.
 * The switch is a file-scope flag set by the entry point (test infrastructure: not re-entrant). */
static int g_bf16 = 0;      /* 0: fp32 arithmetic, 1: bf16, 2: fp16 (every op rounds to that dtype) */
static inline float bf16_rne(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return x;                 /* NaN */
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    float r; memcpy(&r, &u, 4); return r;
}
static inline float RB(float x) { return g_bf16 == 1 ? bf16_rne(x) : g_bf16 == 2 ? (float)(_Float16)x : x; }
#define ORC_EPS_T (g_bf16 ? RB(ORC_EPS) : ORC_EPS)                /* clamp_min(eps) on a tensor of that dtype (0 in fp16) */

/* ------------------------------------------------------------------------------------------
 * make_k_quants: quant_utils.py:199-274  (Q2_K, Q4_K, Q5_K; asymmetric weighted LSQ search)
 * Operates on all G groups of one get_scale_and_zero call at once because the
 * `if not valid.any(): continue` test (quant_utils.py:250-252) is global over the call.
 * x: group g starts at x + g*gstride_hi ... we pass an accessor via row/grp strides.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float xmin;       /* aliases best_min (quant_utils.py:228) */
    float xmax;
    float best_scale;
    float best_err;
    float sum_w, sum_x;
    float s_l, s_l2, s_xl; /* per-iteration sums */
    int isconst;
} kq_state_t;

static void group_weights(const float *x, int n, float *w) {
    float t[32];
    for (int k = 0; k < n; ++k) t[k] = RB(x[k] * x[k]);
    float sum_x2 = RB(sum8(t, n));                  /* :203 */
    float av_x = RB(sqrtf(RB(sum_x2 / (float)n)));  /* :204 (IEEE sqrt) */
    for (int k = 0; k < n; ++k) w[k] = RB(av_x + fabsf(x[k])); /* :205 */
}

static void make_k_quants_call(const float *x, long row_stride, int rows, int n, int maxq,
                               double rmin, double rdelta, int nstep,
                               float *scale_out, float *zero_out, uint32_t *flags) {
    const int gpr = QK_K / n;           /* groups per row */
    const long G = (long)rows * gpr;
    const float fmaxq = (float)maxq;
    kq_state_t *st = (kq_state_t *)malloc(sizeof(kq_state_t) * (size_t)G);

#pragma omp parallel for schedule(static)
    for (long g = 0; g < G; ++g) {
        const float *xg = x + (g / gpr) * row_stride + (g % gpr) * n;
        kq_state_t *s = &st[g];
        float w[32], t[32];
        group_weights(xg, n, w);
        float mn = xg[0], mx = xg[0];
        for (int k = 1; k < n; ++k) { mn = fminf(mn, xg[k]); mx = fmaxf(mx, xg[k]); }
        mn = fminf(mn, 0.0f);                        /* :210 */
        s->isconst = (mx == mn);                     /* :211 */
        s->sum_w = RB(sum8(w, n));                   /* :214 */
        for (int k = 0; k < n; ++k) t[k] = RB(w[k] * xg[k]);
        s->sum_x = RB(sum8(t, n));                   /* :215 */
        float scale = RB(RB(mx - mn) / fmaxq);       /* :218 */
        if (s->isconst) scale = 0.0f;                /* :219 */
        float iscale = RB(1.0f / fmaxf(scale, ORC_EPS_T)); /* :220 */
        for (int k = 0; k < n; ++k) {
            float q = clampf(rintf(RB(RB(xg[k] - mn) * iscale)), 0.0f, fmaxq); /* :223 */
            if (s->isconst) q = 0.0f;                /* :225 */
            float diff = RB(RB(RB(scale * q) + mn) - xg[k]);   /* :230 */
            t[k] = RB(w[k] * RB(diff * diff));       /* :231-232 */
        }
        s->best_err = RB(sum8(t, n));
        s->xmin = mn; s->xmax = mx; s->best_scale = scale;
    }

    if (nstep >= 1) {                                 /* :235 */
        for (int i = 0; i <= nstep; ++i) {            /* :240 */
            /* python double arithmetic, then cast to fp32 when multiplied into the tensor */
            const float num = (float)(rmin + rdelta * (double)i + (double)maxq);
            int any_valid = 0;
#pragma omp parallel for schedule(static) reduction(| : any_valid)
            for (long g = 0; g < G; ++g) {
                const float *xg = x + (g / gpr) * row_stride + (g % gpr) * n;
                kq_state_t *s = &st[g];
                float w[32], a[32], b[32], c[32];
                group_weights(xg, n, w);
                /* :241  scalar / tensor  ==  reciprocal(tensor) * scalar */
                float is = RB(RB(1.0f / fmaxf(RB(s->xmax - s->xmin), ORC_EPS_T)) * num);
                for (int k = 0; k < n; ++k) {
                    float qf = clampf(rintf(RB(RB(xg[k] - s->xmin) * is)), 0.0f, fmaxq); /* :242 */
                    uint8_t L = s->isconst ? 0 : (uint8_t)qf;                      /* :243 */
                    uint8_t L2 = (uint8_t)(L * L);      /* :246 uint8 ** 2 wraps */
                    a[k] = RB(w[k] * (float)L);             /* :245 */
                    b[k] = RB(w[k] * (float)L2);            /* :246 */
                    c[k] = RB(RB(w[k] * xg[k]) * (float)L); /* :247 */
                }
                s->s_l = RB(sum8(a, n)); s->s_l2 = RB(sum8(b, n)); s->s_xl = RB(sum8(c, n));
                float D = RB(RB(s->sum_w * s->s_l2) - RB(s->s_l * s->s_l));   /* :249 */
                if (D > ORC_EPS) any_valid |= 1;                   /* :250 */
            }
            if (!any_valid) continue;                              /* :251-252 */
            if (flags) flags[0] |= (1u << i);
            int any_acc = 0;
#pragma omp parallel for schedule(static) reduction(| : any_acc)
            for (long g = 0; g < G; ++g) {
                const float *xg = x + (g / gpr) * row_stride + (g % gpr) * n;
                kq_state_t *s = &st[g];
                float w[32], t[32];
                group_weights(xg, n, w);
                float is = RB(RB(1.0f / fmaxf(RB(s->xmax - s->xmin), ORC_EPS_T)) * num);
                float D = RB(RB(s->sum_w * s->s_l2) - RB(s->s_l * s->s_l));
                float sc = RB(RB(RB(s->sum_w * s->s_xl) - RB(s->sum_x * s->s_l)) / D);  /* :254 */
                float mn = RB(RB(RB(s->s_l2 * s->sum_x) - RB(s->s_l * s->s_xl)) / D);   /* :255 */
                if (mn > 0.0f) {                                           /* :257-260 */
                    sc = RB(s->s_xl / fmaxf(s->s_l2, ORC_EPS_T));
                    mn = 0.0f;
                }
                for (int k = 0; k < n; ++k) {
                    float qf = clampf(rintf(RB(RB(xg[k] - s->xmin) * is)), 0.0f, fmaxq);
                    uint8_t L = s->isconst ? 0 : (uint8_t)qf;
                    float diff = RB(RB(RB(sc * (float)L) + mn) - xg[k]);   /* :262 */
                    t[k] = RB(w[k] * RB(diff * diff));                     /* :263-264 */
                }
                float cand = RB(sum8(t, n));
                if (cand < s->best_err) {                                  /* :266-270 */
                    s->best_err = cand; s->best_scale = sc; s->xmin = mn;  /* xmin IS best_min */
                    any_acc |= 1;
                }
            }
            if (flags && any_acc) flags[1] |= (1u << i);
        }
    }
    for (long g = 0; g < G; ++g) {
        scale_out[g] = st[g].best_scale;
        zero_out[g] = -st[g].xmin;                   /* :273 */
    }
    free(st);
}

/* make_quants: quant_utils.py:147-197 (Q3_K, Q6_K), quant_scale == absmax only. */
static void make_quants_call(const float *x, long row_stride, int rows, int n, int maxq,
                             float *scale_out, float *zero_out) {
    const int gpr = QK_K / n;
    const long G = (long)rows * gpr;
#pragma omp parallel for schedule(static)
    for (long g = 0; g < G; ++g) {
        const float *xg = x + (g / gpr) * row_stride + (g % gpr) * n;
        float mn = xg[0], mx = xg[0];
        for (int k = 1; k < n; ++k) { mn = fminf(mn, xg[k]); mx = fmaxf(mx, xg[k]); }
        mx = fmaxf(fabsf(mn), mx);                   /* :153 */
        if (mn < 0.0f) mn = -mx;                     /* :154-156 */
        if (mn == mx) { mn = -1.0f; mx = 1.0f; }     /* :157-159 */
        scale_out[g] = RB(RB(mx - mn) / (float)maxq);      /* :161 */
        zero_out[g] = 0.0f;                          /* :195 */
    }
}

/* ------------------------------------------------------------------------------------------
 * Quantizer.get_scale_and_zero: quant_utils.py:90-145.
 * x: (rows, 256) fp32 with row stride `row_stride` (elements).
 * Outputs (each strided by *_stride elements per row):
 *   d, dmin : fp16 bit patterns, one per row
 *   sq, zq  : gpr bytes per row (uint8 for Q2/4/5_K, int8 for Q3/6_K; same bit pattern, values >= 0)
 * flags[2] (optional): bit i of flags[0] = candidate i had some group with D > eps;
 *                      bit i of flags[1] = candidate i was accepted by some group.
 * ---------------------------------------------------------------------------------------- */
int orc_get_scale_and_zero(const float *x, long row_stride, int rows, int qtype,
                           double rmin, double rdelta, int nstep,
                           uint16_t *d, long d_stride, uint16_t *dmin, long dmin_stride,
                           uint8_t *sq, long sq_stride, uint8_t *zq, long zq_stride,
                           uint32_t *flags) {
    orc_fmt_t f;
    if (orc_fmt(qtype, &f)) return -1;
    const int n = f.group_size, gpr = QK_K / n, maxq = (1 << f.bits) - 1;
    const long G = (long)rows * gpr;
    float *gs = (float *)malloc(sizeof(float) * (size_t)G);
    float *gz = (float *)malloc(sizeof(float) * (size_t)G);
    if (flags) flags[0] = flags[1] = 0;
    if (f.asym) make_k_quants_call(x, row_stride, rows, n, maxq, rmin, rdelta, nstep, gs, gz, flags);
    else make_quants_call(x, row_stride, rows, n, maxq, gs, gz);
    const float smq = (float)f.scale_maxq;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const float *s = gs + (long)r * gpr, *z = gz + (long)r * gpr;
        float ms = s[0], mz = z[0];
        for (int g = 1; g < gpr; ++g) { ms = fmaxf(ms, s[g]); mz = fmaxf(mz, z[g]); }  /* :121 */
        d[r * d_stride] = f2h(RB(ms / smq));                                             /* :124 */
        dmin[r * dmin_stride] = f2h(RB(mz / smq));                                       /* :125 */
        float inv_s = ms > 0.0f ? RB(RB(1.0f / ms) * smq) : 0.0f;                        /* :128 */
        float inv_z = mz > 0.0f ? RB(RB(1.0f / mz) * smq) : 0.0f;                        /* :129 */
        for (int g = 0; g < gpr; ++g) {
            sq[r * sq_stride + g] = (uint8_t)(int)clampf(rintf(RB(inv_s * s[g])), 0.0f, smq); /* :132-137 */
            zq[r * zq_stride + g] = (uint8_t)(int)clampf(rintf(RB(inv_z * z[g])), 0.0f, smq); /* :138-143 */
        }
    }
    free(gs); free(gz);
    return 0;
}

/* quantize / dequantize: quant_utils.py:34-46 */
static inline float q_quant(float x, float d, float sq, float dm, float zq, float lo, float hi) {
    float q = rintf((x + dm * zq) / fmaxf(d * sq, ORC_EPS));
    return clampf(q, lo, hi);
}
static inline float q_dequant(float q, float d, float sq, float dm, float zq) {
    return (d * sq) * q - (dm * zq);
}

static inline float code_to_f(uint8_t b, int is_signed) { return is_signed ? (float)(int8_t)b : (float)b; }

/* ------------------------------------------------------------------------------------------
 * GPTQ.step: gptq.py:146-295.
 * W (d_row, d_col) fp32 row-major, IN: weights after quantization_pre_step (ORIGINAL column order);
 * OUT: the dequantised weights in the original column order (gptq.py:266 writes w_q back into w; with
 * act_order w lives in permuted order and dequantize_linear_weight of the un-permuted codes is what the
 * caller writes back, quantizer.py:257-264 -- the same values).
 * U = H_inv_cho, upper triangular, element (i,j) at U[i*u_rs + j*u_cs]; with act_order it belongs to the
 * PERMUTED Hessian H[perm][:, perm] (gptq.py:212-213).
 * static_groups (gptq.py:184-196): all scales / zeros are searched up front on the original W.
 * perm (act_order, gptq.py:209-216; requires static_groups, :45-46): column c of the loop is original
 * column perm[c]; its group / super-group are perm[c]/group_size, perm[c]/256 (:233-235); the codes are
 * un-permuted at the end (:276-277).  NULL = no act_order.
 * Q3_K ignores both options (gptq.py:204-206) -- the caller passes 0 / NULL for it.
 * Outputs: qweight (d_row,d_col) codes (uint8 / int8 bit patterns); d,dmin (d_row, d_col/256) fp16
 * bits; sq,zq (d_row, d_col/group_size).
 * ---------------------------------------------------------------------------------------- */
int orc_gptq_step_ex(float *W, const float *U, long u_rs, long u_cs, int d_row, int d_col, int qtype,
                     int block_size, double rmin, double rdelta, int nstep, int static_groups, const int *perm,
                     uint8_t *qweight, uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq,
                     uint32_t *flags /* 2 per super-block or NULL */) {
    orc_fmt_t f;
    if (orc_fmt(qtype, &f)) return -1;
    if (d_col % QK_K) return -2;
    if (perm && !static_groups) return -3;
    const int n = f.group_size, ng = d_col / n, nsb = d_col / QK_K, gpr = QK_K / n;
    const int is_signed = !f.asym;
    const float lo = (float)f.qmin, hi = (float)f.qmax;
    /* row-major private copy of U so that the hot loops are contiguous */
    float *Ur = (float *)malloc(sizeof(float) * (size_t)d_col * d_col);
    for (long i = 0; i < d_col; ++i)
        for (long j = 0; j < d_col; ++j) Ur[i * d_col + j] = U[i * u_rs + j * u_cs];
    float *errs = (float *)malloc(sizeof(float) * (size_t)d_row * block_size);

    if (static_groups) {                                            /* gptq.py:184-196, on the original w */
        for (int s = 0; s < nsb; ++s)
            orc_get_scale_and_zero(W + (long)s * QK_K, d_col, d_row, qtype, rmin, rdelta, nstep,
                                   d + s, nsb, dmin + s, nsb, sq + (long)s * gpr, ng, zq + (long)s * gpr, ng,
                                   flags ? flags + 2 * s : NULL);
    }
    if (perm) {                                                     /* gptq.py:211: w = w[:, perm] */
        float *tmp = (float *)malloc(sizeof(float) * d_col);
        for (int r = 0; r < d_row; ++r) {
            float *wrow = W + (long)r * d_col;
            for (int c = 0; c < d_col; ++c) tmp[c] = wrow[perm[c]];
            memcpy(wrow, tmp, sizeof(float) * d_col);
        }
        free(tmp);
    }
    uint8_t *qw = perm ? (uint8_t *)malloc((size_t)d_row * d_col) : qweight;

    for (int c1 = 0; c1 < d_col; c1 += block_size) {               /* gptq.py:222 */
        const int c2 = c1 + block_size < d_col ? c1 + block_size : d_col;
        const int ncols = c2 - c1;
        /* Super-block searches that start inside this block read the LIVE matrix w
         * (gptq.py:240-241), i.e. without this block's rank-1 updates (those only touch w_blk). */
        if (!static_groups)
            for (int i = 0; i < ncols; ++i) {
                const int col = c1 + i;
                if (col % QK_K == 0) {
                    const int s = col / QK_K;
                    orc_get_scale_and_zero(W + col, d_col, d_row, qtype, rmin, rdelta, nstep,
                                           d + s, nsb, dmin + s, nsb,
                                           sq + (long)s * gpr, ng, zq + (long)s * gpr, ng,
                                           flags ? flags + 2 * s : NULL);
                }
            }
#pragma omp parallel for schedule(static)
        for (int r = 0; r < d_row; ++r) {
            float wb[1024];
            float *wrow = W + (long)r * d_col;
            float *er = errs + (long)r * block_size;
            memcpy(wb, wrow + c1, sizeof(float) * ncols);            /* gptq.py:225 */
            for (int i = 0; i < ncols; ++i) {
                const int col = c1 + i;
                const int ocol = perm ? perm[col] : col;             /* :233-238 group of the ORIGINAL column */
                const int s = ocol / QK_K, g = ocol / n;
                const float *urow = Ur + (long)col * d_col + c1;     /* U[col, c1:c2] */
                const float dd = h2f(d[(long)r * nsb + s]), dm = h2f(dmin[(long)r * nsb + s]);
                const float fs = code_to_f(sq[(long)r * ng + g], is_signed);
                const float fz = code_to_f(zq[(long)r * ng + g], is_signed);
                const float x = wb[i];
                const float q = q_quant(x, dd, fs, dm, fz, lo, hi);  /* :247-254 */
                const float wq = q_dequant(q, dd, fs, dm, fz);       /* :255-261 */
                qw[(long)r * d_col + col] = (uint8_t)(int8_t)(int)q; /* :263 */
                const float err = (x - wq) / urow[i];                /* :264 */
                wrow[col] = wq;                                      /* :266 */
                const float nerr = -err;                             /* alpha=-1 folded into vec1 */
                for (int j = i; j < ncols; ++j) {                    /* :267 addr_, no FMA */
                    float p = nerr * urow[j];
                    wb[j] = wb[j] + p;
                }
                er[i] = err;                                         /* :268 */
            }
            /* :270 addmm_: single-accumulator FMA chain over k, then one subtract */
            const int ntrail = d_col - c2;
            float *wt = wrow + c2;
            for (int j0 = 0; j0 < ntrail; j0 += 64) {
                const int jn = ntrail - j0 < 64 ? ntrail - j0 : 64;
                float acc[64];
                for (int j = 0; j < jn; ++j) acc[j] = 0.0f;
                for (int k = 0; k < ncols; ++k) {
                    const float e = er[k];
                    const float *uk = Ur + (long)(c1 + k) * d_col + c2 + j0;
                    for (int j = 0; j < jn; ++j) acc[j] = fmaf(e, uk[j], acc[j]);
                }
                for (int j = 0; j < jn; ++j) wt[j0 + j] = wt[j0 + j] - acc[j];
            }
        }
    }
    if (perm) {                                                     /* gptq.py:276-277: qweight[:, invperm]; same for w */
        float *tmp = (float *)malloc(sizeof(float) * d_col);
        for (int r = 0; r < d_row; ++r) {
            float *wrow = W + (long)r * d_col;
            for (int c = 0; c < d_col; ++c) {
                qweight[(long)r * d_col + perm[c]] = qw[(long)r * d_col + c];
                tmp[perm[c]] = wrow[c];
            }
            memcpy(wrow, tmp, sizeof(float) * d_col);
        }
        free(tmp); free(qw);
    }
    free(errs); free(Ur);
    return 0;
}

int orc_gptq_step(float *W, const float *U, long u_rs, long u_cs, int d_row, int d_col, int qtype,
                  int block_size, double rmin, double rdelta, int nstep,
                  uint8_t *qweight, uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq, uint32_t *flags) {
    return orc_gptq_step_ex(W, U, u_rs, u_cs, d_row, d_col, qtype, block_size, rmin, rdelta, nstep, 0, NULL,
                            qweight, d, dmin, sq, zq, flags);
}

/* ------------------------------------------------------------------------------------------
 * Quantizer._quant_non_block_module: quantizer.py:278-330 (RTN K-quant, fp32 weights).
 * W is read-only.
 * ---------------------------------------------------------------------------------------- */
int orc_rtn_quantize(const float *W, int d_row, int d_col, int qtype,
                     double rmin, double rdelta, int nstep,
                     uint8_t *qweight, uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq) {
    orc_fmt_t f;
    if (orc_fmt(qtype, &f)) return -1;
    if (d_col % QK_K) return -2;
    const int n = f.group_size, ng = d_col / n, nsb = d_col / QK_K, gpr = QK_K / n;
    const int is_signed = !f.asym;
    const float lo = (float)f.qmin, hi = (float)f.qmax;
    for (int s = 0; s < nsb; ++s)                                       /* quantizer.py:302-309 */
        orc_get_scale_and_zero(W + (long)s * QK_K, d_col, d_row, qtype, rmin, rdelta, nstep,
                               d + s, nsb, dmin + s, nsb, sq + (long)s * gpr, ng, zq + (long)s * gpr, ng, NULL);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < d_row; ++r)
        for (int c = 0; c < d_col; ++c) {                               /* quantizer.py:323 */
            const int s = c / QK_K, g = c / n;
            float q = q_quant(W[(long)r * d_col + c], h2f(d[(long)r * nsb + s]),
                              code_to_f(sq[(long)r * ng + g], is_signed), h2f(dmin[(long)r * nsb + s]),
                              code_to_f(zq[(long)r * ng + g], is_signed), lo, hi);
            qweight[(long)r * d_col + c] = (uint8_t)(int8_t)(int)q;
        }
    return 0;
}

/* The same for a bf16 weight (values passed widened to fp32): the scale search runs in bf16 arithmetic (see RB above), the
 * final quantize() in fp32 like the reference's (quant_utils.py:34-40 promotes bf16 + fp32 to fp32). */
int orc_rtn_quantize_bf16(const float *W, int d_row, int d_col, int qtype,
                          double rmin, double rdelta, int nstep,
                          uint8_t *qweight, uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq) {
    g_bf16 = 1;
    const int rc = orc_rtn_quantize(W, d_row, d_col, qtype, rmin, rdelta, nstep, qweight, d, dmin, sq, zq);
    g_bf16 = 0;
    return rc;
}
/* ... and for an fp16 weight (probe only: agreement with the reference is reported by
 * tests/golden/check_oracle_vs_reference_large.py, not asserted -- see DESIGN.md section 2). */
int orc_rtn_quantize_fp16(const float *W, int d_row, int d_col, int qtype,
                          double rmin, double rdelta, int nstep,
                          uint8_t *qweight, uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq) {
    g_bf16 = 2;
    const int rc = orc_rtn_quantize(W, d_row, d_col, qtype, rmin, rdelta, nstep, qweight, d, dmin, sq, zq);
    g_bf16 = 0;
    return rc;
}

/* dequantize_linear_weight: quant_utils.py:277-310 */
int orc_dequantize(int qtype, const uint8_t *qweight, const uint16_t *d, const uint8_t *sq,
                   const uint16_t *dmin, const uint8_t *zq, int d_row, int d_col, float *out) {
    orc_fmt_t f;
    if (orc_fmt(qtype, &f)) return -1;
    const int n = f.group_size, ng = d_col / n, nsb = d_col / QK_K;
    const int is_signed = !f.asym;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < d_row; ++r)
        for (int c = 0; c < d_col; ++c) {
            const int s = c / QK_K, g = c / n;
            out[(long)r * d_col + c] =
                q_dequant(code_to_f(qweight[(long)r * d_col + c], is_signed), h2f(d[(long)r * nsb + s]),
                          code_to_f(sq[(long)r * ng + g], is_signed), h2f(dmin[(long)r * nsb + s]),
                          code_to_f(zq[(long)r * ng + g], is_signed));
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Byte packers: packing_utils.py:8-326.  Inputs are NOT modified (the reference's pack_Q3K /
 * pack_Q6K add the +4/+32 offsets in place; we add them on the fly).
 * out: (d_row, d_col/256 * type_size) bytes.
 * ---------------------------------------------------------------------------------------- */
static void pack_scale_min(const uint8_t *sc, const uint8_t *mn, uint8_t *p) { /* :8-30 */
    for (int j = 0; j < 4; ++j) {
        p[j] = (uint8_t)(sc[j] | ((sc[4 + j] >> 4) << 6));
        p[4 + j] = (uint8_t)(mn[j] | ((mn[4 + j] >> 4) << 6));
        p[8 + j] = (uint8_t)((sc[4 + j] & 0x0F) | ((mn[4 + j] & 0x0F) << 4));
    }
}

int orc_pack(int qtype, const uint8_t *qweight, const uint16_t *d, const uint8_t *sq,
             const uint16_t *dmin, const uint8_t *zq, int d_row, int d_col, uint8_t *out) {
    orc_fmt_t f;
    if (orc_fmt(qtype, &f)) return -1;
    if (d_col % QK_K) return -2;
    const int nsb = d_col / QK_K, gpr = QK_K / f.group_size;
    const long nblk = (long)d_row * nsb;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < nblk; ++b) {
        const uint8_t *q = qweight + b * QK_K;
        const uint8_t *s = sq + b * gpr, *z = zq + b * gpr;
        uint8_t *o = out + b * f.type_size;
        uint8_t u[QK_K];
        memset(o, 0, (size_t)f.type_size);
        switch (qtype) {
        case 10: /* Q2_K :33-77  scales[16] qs[64] d dmin */
            for (int j = 0; j < 16; ++j) o[j] = (uint8_t)((s[j] & 0x0F) | ((z[j] & 0x0F) << 4));
            for (int c = 0; c < 2; ++c)
                for (int l = 0; l < 32; ++l)
                    o[16 + 32 * c + l] = (uint8_t)(q[128 * c + l] | (q[128 * c + 32 + l] << 2) |
                                                   (q[128 * c + 64 + l] << 4) | (q[128 * c + 96 + l] << 6));
            memcpy(o + 80, &d[b], 2); memcpy(o + 82, &dmin[b], 2);
            break;
        case 11: /* Q3_K :80-142  hmask[32] qs[64] scales[12] d */
            for (int j = 0; j < QK_K; ++j) u[j] = (uint8_t)((int8_t)q[j] + 4);
            for (int j = 0; j < QK_K; ++j)
                if (u[j] > 3) { o[j % 32] |= (uint8_t)(1 << (j / 32)); u[j] = (uint8_t)(u[j] - 4); }
            for (int c = 0; c < 2; ++c)
                for (int l = 0; l < 32; ++l)
                    o[32 + 32 * c + l] = (uint8_t)(u[128 * c + l] | (u[128 * c + 32 + l] << 2) |
                                                   (u[128 * c + 64 + l] << 4) | (u[128 * c + 96 + l] << 6));
            for (int j = 0; j < 16; ++j) {
                uint8_t lj = (uint8_t)((int8_t)s[j] + 32);
                uint8_t lo4 = lj & 0x0F, hi2 = (lj >> 4) & 0x03;
                if (j < 8) o[96 + j] |= lo4; else o[96 + j - 8] |= (uint8_t)(lo4 << 4);
                o[96 + 8 + (j % 4)] |= (uint8_t)(hi2 << (2 * (j / 4)));
            }
            memcpy(o + 108, &d[b], 2);
            break;
        case 12: /* Q4_K :145-190  d dmin scales[12] qs[128] */
            memcpy(o, &d[b], 2); memcpy(o + 2, &dmin[b], 2);
            pack_scale_min(s, z, o + 4);
            for (int c = 0; c < 4; ++c)
                for (int l = 0; l < 32; ++l)
                    o[16 + 32 * c + l] = (uint8_t)(q[64 * c + l] | (q[64 * c + 32 + l] << 4));
            break;
        case 13: /* Q5_K :193-262  d dmin scales[12] qh[32] ql[128] */
            memcpy(o, &d[b], 2); memcpy(o + 2, &dmin[b], 2);
            pack_scale_min(s, z, o + 4);
            for (int c = 0; c < 4; ++c)
                for (int l = 0; l < 32; ++l) {
                    uint8_t l1 = q[64 * c + l], l2 = q[64 * c + 32 + l];
                    if (l1 > 15) { o[16 + l] |= (uint8_t)(1 << (2 * c)); l1 = (uint8_t)(l1 - 16); }
                    if (l2 > 15) { o[16 + l] |= (uint8_t)(2 << (2 * c)); l2 = (uint8_t)(l2 - 16); }
                    o[48 + 32 * c + l] = (uint8_t)(l1 | (l2 << 4));
                }
            break;
        case 14: /* Q6_K :265-326  ql[128] qh[64] scales[16] d */
            for (int j = 0; j < QK_K; ++j) u[j] = (uint8_t)((int8_t)q[j] + 32);
            for (int c = 0; c < 2; ++c)
                for (int l = 0; l < 32; ++l) {
                    uint8_t v0 = u[128 * c + l], v1 = u[128 * c + 32 + l];
                    uint8_t v2 = u[128 * c + 64 + l], v3 = u[128 * c + 96 + l];
                    o[64 * c + l] = (uint8_t)((v0 & 0xF) | ((v2 & 0xF) << 4));
                    o[64 * c + 32 + l] = (uint8_t)((v1 & 0xF) | ((v3 & 0xF) << 4));
                    o[128 + 32 * c + l] = (uint8_t)(((v0 >> 4) & 3) | (((v1 >> 4) & 3) << 2) |
                                                    (((v2 >> 4) & 3) << 4) | (((v3 >> 4) & 3) << 6));
                }
            memcpy(o + 192, s, 16);
            memcpy(o + 208, &d[b], 2);
            break;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * GPTQ.update: gptq.py:80-114.   H = beta*H + alpha * X^T X   (X: n_tok x d_col fp32).
 * The reference calls MKL sgemm whose K-blocking at K=2048 is not characterised, so this
 * boundary (B3) is statistical; the oracle accumulates X^T X in double and rounds once.
 * ---------------------------------------------------------------------------------------- */
int orc_hessian_update(float *H, const float *X, long n_tok, int d_col, float beta, float alpha) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < d_col; ++i)
        for (int j = 0; j < d_col; ++j) {
            double acc = 0.0;
            for (long t = 0; t < n_tok; ++t) acc += (double)X[t * d_col + i] * (double)X[t * d_col + j];
            H[(long)i * d_col + j] = (float)((double)beta * (double)H[(long)i * d_col + j] + (double)alpha * acc);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * quantization_pre_step + _prepare: gptq.py:123-143, 305-324; linalg_utils.py:8-12.
 * H (n,n) fp32 in/out (masked + damped, as the reference leaves it), W (d_row,n) in/out
 * (dead columns zeroed), U_out (n,n) fp32 ROW-major upper Cholesky factor of inv(H).
 * LAPACK bits are not reproducible (the reference differs from itself across thread
 * counts), so the factorisation is done in double and rounded: statistical boundary B2.
 * Returns 0, or 1 if H is not positive definite (the reference then uses U = I, gptq.py:321-323).
 * ---------------------------------------------------------------------------------------- */
int orc_prepare(float *H, float *W, int d_row, int n, float rel_damp, float *U_out) {
    for (int i = 0; i < n; ++i)                               /* gptq.py:134-141 */
        if (H[(long)i * n + i] == 0.0f) {
            H[(long)i * n + i] = 1.0f;
            for (int r = 0; r < d_row; ++r) W[(long)r * n + i] = 0.0f;
        }
    for (int j = 0; j < n; ++j) {                             /* gptq.py:308-313 */
        int allz = 1;
        for (int r = 0; r < d_row && allz; ++r) allz = (W[(long)r * n + j] == 0.0f);
        if (allz) {
            for (int k = 0; k < n; ++k) { H[(long)j * n + k] = 0.0f; H[(long)k * n + j] = 0.0f; }
            H[(long)j * n + j] = 1.0f;
        }
    }
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += (double)H[(long)i * n + i];
    const float damp = rel_damp * (float)(mean / n);          /* gptq.py:315 */
    for (int i = 0; i < n; ++i) H[(long)i * n + i] = H[(long)i * n + i] + damp;

    double *A = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *B = (double *)malloc(sizeof(double) * (size_t)n * n);
    int fail = 0;
    /* L = chol(H) lower, in A */
    for (long i = 0; i < (long)n * n; ++i) A[i] = (double)H[i];
    for (int j = 0; j < n && !fail; ++j) {
        double s = A[(long)j * n + j];
        for (int k = 0; k < j; ++k) s -= A[(long)j * n + k] * A[(long)j * n + k];
        if (!(s > 0.0)) { fail = 1; break; }
        const double ljj = sqrt(s);
        A[(long)j * n + j] = ljj;
#pragma omp parallel for schedule(static)
        for (int i = j + 1; i < n; ++i) {
            double t = A[(long)i * n + j];
            for (int k = 0; k < j; ++k) t -= A[(long)i * n + k] * A[(long)j * n + k];
            A[(long)i * n + j] = t / ljj;
        }
    }
    if (!fail) {
        /* B = inv(L) (lower), column by column */
#pragma omp parallel for schedule(dynamic, 8)
        for (int c = 0; c < n; ++c) {
            for (int i = 0; i < c; ++i) B[(long)i * n + c] = 0.0;
            B[(long)c * n + c] = 1.0 / A[(long)c * n + c];
            for (int i = c + 1; i < n; ++i) {
                double t = 0.0;
                for (int k = c; k < i; ++k) t -= A[(long)i * n + k] * B[(long)k * n + c];
                B[(long)i * n + c] = t / A[(long)i * n + i];
            }
        }
        /* A = inv(H) = inv(L)^T inv(L)  (symmetric; fill lower) */
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                double t = 0.0;
                for (int k = i; k < n; ++k) t += B[(long)k * n + i] * B[(long)k * n + j];
                A[(long)i * n + j] = t;
            }
        /* upper Cholesky of inv(H): U^T U = inv(H); compute lower L2 = chol(inv(H)), U = L2^T */
        for (int j = 0; j < n && !fail; ++j) {
            double s = A[(long)j * n + j];
            for (int k = 0; k < j; ++k) s -= A[(long)j * n + k] * A[(long)j * n + k];
            if (!(s > 0.0)) { fail = 1; break; }
            const double ljj = sqrt(s);
            A[(long)j * n + j] = ljj;
#pragma omp parallel for schedule(static)
            for (int i = j + 1; i < n; ++i) {
                double t = A[(long)i * n + j];
                for (int k = 0; k < j; ++k) t -= A[(long)i * n + k] * A[(long)j * n + k];
                A[(long)i * n + j] = t / ljj;
            }
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (fail) U_out[(long)i * n + j] = (i == j) ? 1.0f : 0.0f;
            else U_out[(long)i * n + j] = (j >= i) ? (float)A[(long)j * n + i] : 0.0f;
        }
    free(A); free(B);
    return fail;
}

int orc_abi_version(void) { return 1; }
