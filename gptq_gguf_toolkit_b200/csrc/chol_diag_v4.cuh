// chol_diag_v4.cuh -- two-level diagonal-block kernel of gq_prepare (csrc/linalg.cu), the default since round 2.
// Same contract as chol_diag_v2 / v3: factor A[k0:k0+128, k0:k0+128] = L L^T in place (lower), inv(L) to Binv, inv(L)^T to
// BinvT (may be null); a non-positive pivot sets *not_pd and is replaced by 1.  (The part of A's block above the diagonal is
// written as zeros; nothing reads it.)
//
// v2 / v3 sweep the 128 columns one by one with a block-wide barrier per column (88.7 / 70.4 us on B200: ~1000 cycles per
// column, most of it barrier + shared-memory round trips on the dependent chain).  Here the dependent chain lives in ONE warp:
//   * per 32-column panel, warp 0 factors the 32 x 32 diagonal block in registers (lane = row, column broadcasts by shuffle:
//     per column one pivot shuffle, one rsqrt, and 31-j independent shuffle + FMA pairs -- no barrier, no shared memory);
//   * the rows below solve  x L_bb^T = p  by substitution, one thread per row, everything in registers (straight-line code,
//     L_bb read as LDS.128 broadcasts);
//   * three block-wide barriers per panel instead of 32 + 2;
//   * a panel's rank-32 update is applied to the next block column at once and to the columns right of it by warps 1..7
//     WHILE warp 0 factors the next diagonal block;
//   * inv(L): the four 32 x 32 diagonal inverses by forward substitution (one thread per column, registers only), the
//     off-diagonal blocks by doubling (32 -> 64 -> 128) with 4 x 4 register tiles.
// The CPU suite runs this source on the SIMT emulator against fp64 (tests/test_simt_emu_cpu.py).
#pragma once

namespace cd4 {
#ifdef CD4_PROFILE      // profiles/microbench/chol_diag_v4.cu: per-phase cycle counts of thread 0
__device__ unsigned long long cd4_clk[8];
#define CD4_LAP(ph) do { if (threadIdx.x == 0) { const long long n_ = clock64(); cd4_clk[ph] += (unsigned long long)(n_ - t_); t_ = n_; } } while (0)
#else
#define CD4_LAP(ph) do { } while (0)
#endif
constexpr int NB = 128, LS = 132, PW = 32, T4 = 256;
constexpr int LS_ = LS;
struct Smem4 { float L[NB * LS]; float X[NB * LS]; float inv[NB]; };

// Shuffles of the warp-level factorisation as volatile asm: ptxas keeps volatile asm statements in program order, and the order
// matters -- the next pivot's broadcast has to enter the shuffle pipe EARLY in a column's burst of shuffles, not behind it.
__device__ __forceinline__ float shfl_idx(float v, int src) {
#ifdef SIMT_EMU
    return __shfl_sync(0xffffffffu, v, src, 32);
#else
    float r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(v), "r"(src));
    return r;
#endif
}
// 1 / sqrt(p) to ~1 ulp for p in [1e-30, 3e38]: MUFU.RSQ (no denormal fix-up) + one Newton step,  y <- y + y (1/2 - (p/2) y y)
__device__ __forceinline__ float rsqrt_nr(float p) {
#ifdef SIMT_EMU
    const float y = rsqrtf(p);
#else
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(p));
#endif
    const float h = 0.5f * p;
    return fmaf(y, fmaf(-h * y, y, 0.5f), y);
}
// One forward-substitution step with four independent partial sums (the single-accumulator form is a 4-cycle-per-term
// dependent chain):  returns  v - sum_{k < j} x[k] l[k]   with l = 32 consecutive floats of shared memory, j compile-time.
template <int J>
__device__ __forceinline__ float subst_dot(float v, const float (&x)[32], const float *l) {
    float a0 = v, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int q = 0; q < (J + 3) / 4; ++q) {
        const float4 l4 = *reinterpret_cast<const float4 *>(l + 4 * q);
        if (4 * q < J) a0 = fmaf(-x[4 * q], l4.x, a0);
        if (4 * q + 1 < J) a1 = fmaf(-x[4 * q + 1], l4.y, a1);
        if (4 * q + 2 < J) a2 = fmaf(-x[4 * q + 2], l4.z, a2);
        if (4 * q + 3 < J) a3 = fmaf(-x[4 * q + 3], l4.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}
template <int J> struct SubstLoop {      // x[j] = (x0[j] - sum_{k<j} x[k] L[j][k]) * inv[j]  for j = J .. 31, in order
    static __device__ __forceinline__ void panel(float (&x)[32], const float *Lbb, const float *inv) {
        x[J] = subst_dot<J>(x[J], x, Lbb + J * LS_) * inv[J];
        SubstLoop<J + 1>::panel(x, Lbb, inv);
    }
    // column cc of inv(L_bb): x[i] = (delta(i, cc) - sum_{k<i} L[i][k] x[k]) * inv[i], zero above the diagonal
    static __device__ __forceinline__ void inverse(float (&x)[32], const float *Lbb, const float *inv, int cc) {
        const float v = subst_dot<J>(J == cc ? 1.0f : 0.0f, x, Lbb + J * LS_);
        x[J] = (J >= cc) ? v * inv[J] : 0.0f;
        SubstLoop<J + 1>::inverse(x, Lbb, inv, cc);
    }
};
template <> struct SubstLoop<32> {
    static __device__ __forceinline__ void panel(float (&)[32], const float *, const float *) {}
    static __device__ __forceinline__ void inverse(float (&)[32], const float *, const float *, int) {}
};

// L[r0 + i][c0 + c] -= sum_{k = kb}^{kb+31} L[r0 + i][k] L[c0 + c][k]  for i < nr, c < nc, entries with column <= row only
// (the rank-32 update of a rectangle of the lower triangle).  4 x 4 register tiles whose rows / columns are nr/4 resp. nc/4
// apart, so that consecutive lanes read consecutive rows of the column operand (conflict-free LDS.128 along k) and lanes with
// the same row operand read it as a broadcast.  `t` = index of the calling thread among the `nthreads` taking part.
__device__ __forceinline__ void syrk_tiles(float *L, int r0, int nr, int c0, int nc, int kb, int t, int nthreads) {
    const int ntr = nr / 4, ntc = nc / 4;
    for (int tt = t; tt < ntr * ntc; tt += nthreads) {
        const int ti = tt / ntc, tc = tt - ti * ntc;
        const int i0 = r0 + ti, j0 = c0 + tc;
        float acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.0f;
#pragma unroll 2
        for (int k = kb; k < kb + 32; k += 4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) av[x] = *reinterpret_cast<const float4 *>(&L[(i0 + ntr * x) * LS + k]);
#pragma unroll
            for (int y = 0; y < 4; ++y) bv[y] = *reinterpret_cast<const float4 *>(&L[(j0 + ntc * y) * LS + k]);
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    acc[x][y] = fmaf(av[x].x, bv[y].x, acc[x][y]);
                    acc[x][y] = fmaf(av[x].y, bv[y].y, acc[x][y]);
                    acc[x][y] = fmaf(av[x].z, bv[y].z, acc[x][y]);
                    acc[x][y] = fmaf(av[x].w, bv[y].w, acc[x][y]);
                }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int i = i0 + ntr * x, c = j0 + ntc * y;
                if (c <= i) L[i * LS + c] -= acc[x][y];
            }
    }
}

// C[i][c] = sign * sum_{k = kbeg}^{kend-1} A[i][k] B[k][c] for one 4 x 4 tile at (i0, c0); A, B, C are blocks of the shared
// arrays (row stride LS), kbeg / kend multiples of 4.  Used for the off-diagonal blocks of inv(L).
__device__ __forceinline__ void tile_prod(const float *A, const float *B, float *C, int i0, int c0, int kbeg, int kend, float sign) {
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.0f;
#pragma unroll 2
    for (int k = kbeg; k < kend; k += 4) {
        float4 av[4], bv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) av[x] = *reinterpret_cast<const float4 *>(&A[(i0 + x) * LS + k]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) bv[kk] = *reinterpret_cast<const float4 *>(&B[(k + kk) * LS + c0]);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            acc[x][0] = fmaf(av[x].x, bv[0].x, acc[x][0]); acc[x][1] = fmaf(av[x].x, bv[0].y, acc[x][1]);
            acc[x][2] = fmaf(av[x].x, bv[0].z, acc[x][2]); acc[x][3] = fmaf(av[x].x, bv[0].w, acc[x][3]);
            acc[x][0] = fmaf(av[x].y, bv[1].x, acc[x][0]); acc[x][1] = fmaf(av[x].y, bv[1].y, acc[x][1]);
            acc[x][2] = fmaf(av[x].y, bv[1].z, acc[x][2]); acc[x][3] = fmaf(av[x].y, bv[1].w, acc[x][3]);
            acc[x][0] = fmaf(av[x].z, bv[2].x, acc[x][0]); acc[x][1] = fmaf(av[x].z, bv[2].y, acc[x][1]);
            acc[x][2] = fmaf(av[x].z, bv[2].z, acc[x][2]); acc[x][3] = fmaf(av[x].z, bv[2].w, acc[x][3]);
            acc[x][0] = fmaf(av[x].w, bv[3].x, acc[x][0]); acc[x][1] = fmaf(av[x].w, bv[3].y, acc[x][1]);
            acc[x][2] = fmaf(av[x].w, bv[3].z, acc[x][2]); acc[x][3] = fmaf(av[x].w, bv[3].w, acc[x][3]);
        }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x)
        *reinterpret_cast<float4 *>(&C[(i0 + x) * LS + c0]) =
            make_float4(sign * acc[x][0], sign * acc[x][1], sign * acc[x][2], sign * acc[x][3]);
}

__global__ void __launch_bounds__(T4) chol_diag_v4_kernel(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem4 &s = *reinterpret_cast<Smem4 *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *Ab = A + (size_t)k0 * ld + k0;
    float *Bb = Binv + (size_t)k0 * ld + k0;
#ifdef CD4_PROFILE
    long long t_ = clock64();
#endif
    {   // the block's lower triangle -> shared memory: 16 independent 16-byte loads per thread, all in flight together
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int id = tid + T4 * u, i = id >> 5, j4 = id & 31;
            v[u] = *reinterpret_cast<const float4 *>(Ab + (size_t)i * ld + 4 * j4);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int id = tid + T4 * u, i = id >> 5, j = 4 * (id & 31);
            *reinterpret_cast<float4 *>(&s.L[i * LS + j]) =
                make_float4(j <= i ? v[u].x : 0.0f, j + 1 <= i ? v[u].y : 0.0f, j + 2 <= i ? v[u].z : 0.0f, j + 3 <= i ? v[u].w : 0.0f);
        }
    }
    __syncthreads();
    CD4_LAP(0);
    for (int o = 0; o < NB; o += PW) {
        // ---- (a) warp 0: Cholesky of the 32 x 32 diagonal block, lane = row, the row's 32 entries in registers.
        //      Meanwhile warps 1..7 apply the PREVIOUS panel (columns o-32 .. o-1) to everything right of the next block
        //      column (rows, columns >= o+32): that part is not needed before the next-but-one panel.
        if (warp == 0) {
            float a[32];
            {
                const float4 *src = reinterpret_cast<const float4 *>(&s.L[(o + lane) * LS + o]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = src[q];
                    a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
                }
            }
            float pj = shfl_idx(a[0], 0);
            float my_iv = 1.0f;
            bool any_bad = false;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const bool bad = !(pj > 1e-30f && pj < 3e38f);      // not positive (or absurd): the reference's not-PD case
                any_bad |= bad;
                if (bad) pj = 1.0f;
                const float iv = rsqrt_nr(pj);
                if (lane == j) my_iv = iv;
                // column j of L: the diagonal is sqrt(p) = p / sqrt(p); lanes above the diagonal carry finite junk that no
                // lane below ever reads (the shuffles below only take from lanes k > j)
                const float lj = (lane == j) ? pj * iv : a[j] * iv;
                a[j] = lj;
                // the column's shuffles, in THIS order: a few first so that column j+1's own update has its operand by the
                // time the next pivot's broadcast is issued, then that broadcast, then the rest
                constexpr int LEAD = 6;
                float ck[32];
#pragma unroll
                for (int k = j + 1; k < 32 && k <= j + LEAD; ++k) ck[k] = shfl_idx(lj, k);
                float pn = 1.0f;
                if (j + 1 < 32) {
                    a[j + 1] = fmaf(-lj, ck[j + 1], a[j + 1]);
                    pn = shfl_idx(a[j + 1], j + 1);
                }
#pragma unroll
                for (int k = j + LEAD + 1; k < 32; ++k) ck[k] = shfl_idx(lj, k);
#pragma unroll
                for (int k = j + 2; k < 32; ++k) a[k] = fmaf(-lj, ck[k], a[k]);
                pj = pn;
            }
            s.inv[o + lane] = my_iv;
            if (any_bad && lane == 0) *not_pd = 1;
            float4 *dst = reinterpret_cast<float4 *>(&s.L[(o + lane) * LS + o]);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(4 * q <= lane ? a[4 * q] : 0.0f, 4 * q + 1 <= lane ? a[4 * q + 1] : 0.0f,
                                     4 * q + 2 <= lane ? a[4 * q + 2] : 0.0f, 4 * q + 3 <= lane ? a[4 * q + 3] : 0.0f);
        } else if (o > 0 && o + PW < NB) {
            const int far = o + PW, S = NB - far;                 // square region [far, NB) x [far, NB), panel k in [o-32, o)
            syrk_tiles(s.L, far, S, far, S, o - PW, tid - 32, T4 - 32);
        } else if (o + PW == NB && warp <= 3) {
            // last panel: warps 1..3 invert the three finished diagonal blocks (one thread per column) meanwhile
            const int ob = 32 * (warp - 1);
            float x[32];
            SubstLoop<0>::inverse(x, &s.L[ob * LS + ob], &s.inv[ob], lane);
#pragma unroll
            for (int i = 0; i < 32; ++i) s.X[(ob + i) * LS + ob + lane] = x[i];
        }
        __syncthreads();
        CD4_LAP(1);
        const int lo = o + PW, R = NB - lo;
        if (R > 0) {
            // ---- (b) rows below the block: x L_bb^T = p by substitution, one thread per row, registers only ----
            if (tid < R) {
                const int r = lo + tid;
                float x[32];
                float4 *row = reinterpret_cast<float4 *>(&s.L[r * LS + o]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = row[q];
                    x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                }
                SubstLoop<0>::panel(x, &s.L[o * LS + o], &s.inv[o]);
#pragma unroll
                for (int q = 0; q < 8; ++q) row[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            }
            __syncthreads();
            CD4_LAP(2);
            // ---- (c) this panel onto the NEXT block column only (rows >= lo, columns lo .. lo+31); the columns right of
            //      it follow under the next panel's (a) ----
            syrk_tiles(s.L, lo, R, lo, PW, o, tid, T4);
            __syncthreads();
            CD4_LAP(3);
        }
    }
    // ---- L is final: back into A (lower, zeros above) as coalesced 16-byte rows; the stores drain under the inverse ----
#pragma unroll 4
    for (int u = 0; u < 16; ++u) {
        const int id = tid + T4 * u, i = id >> 5, j = 4 * (id & 31);
        const float4 l = *reinterpret_cast<const float4 *>(&s.L[i * LS + j]);
        *reinterpret_cast<float4 *>(Ab + (size_t)i * ld + j) =
            make_float4(j <= i ? l.x : 0.0f, j + 1 <= i ? l.y : 0.0f, j + 2 <= i ? l.z : 0.0f, j + 3 <= i ? l.w : 0.0f);
    }
    // ---- X = inv(L): the last diagonal 32 x 32 block by forward substitution, one thread per column, registers only (the
    //      other three were done under the last panel); entries above a column's diagonal come out as exact zeros ----
    if (warp == 0) {
        const int ob = NB - PW;
        float x[32];
        SubstLoop<0>::inverse(x, &s.L[ob * LS + ob], &s.inv[ob], lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) s.X[(ob + i) * LS + ob + lane] = x[i];
    }
    __syncthreads();
    CD4_LAP(4);
    // ---- off-diagonal blocks by doubling:  inv([[A, 0], [C, B]]) = [[Ai, 0], [-Bi C Ai, Bi]].  Level 1 joins the 32-blocks
    //      (0,1) and (2,3), level 2 the two 64-blocks; T = C Ai goes to the (unused) mirror block of X above the diagonal.
    //      4 x 4 tiles; the k ranges skip the zero parts of the triangular factors. ----
    if (tid < 128) {
        const int pr = tid >> 6, ti = (tid >> 3) & 7, tc = tid & 7, p = 64 * pr, q = p + 32;
        // T[i][c] = sum_{k >= c} L[q+i][p+k] X[p+k][p+c]
        tile_prod(&s.L[q * LS + p], &s.X[p * LS + p], &s.X[p * LS + q], 4 * ti, 4 * tc, 4 * tc, 32, 1.0f);
    }
    __syncthreads();
    if (tid < 128) {
        const int pr = tid >> 6, ti = (tid >> 3) & 7, tc = tid & 7, p = 64 * pr, q = p + 32;
        // X[q+i][p+c] = -sum_{k <= i} X[q+i][q+k] T[k][c]
        tile_prod(&s.X[q * LS + q], &s.X[p * LS + q], &s.X[q * LS + p], 4 * ti, 4 * tc, 0, 4 * ti + 4, -1.0f);
    }
    __syncthreads();
    {
        const int ti = tid >> 4, tc = tid & 15;
        tile_prod(&s.L[64 * LS], &s.X[0], &s.X[64], 4 * ti, 4 * tc, 4 * tc, 64, 1.0f);
        __syncthreads();
        tile_prod(&s.X[64 * LS + 64], &s.X[64], &s.X[64 * LS], 4 * ti, 4 * tc, 0, 4 * ti + 4, -1.0f);
    }
    __syncthreads();
    CD4_LAP(5);
    // ---- inv(L) to Binv as coalesced 16-byte rows ----
#pragma unroll 4
    for (int u = 0; u < 16; ++u) {
        const int id = tid + T4 * u, i = id >> 5, j = 4 * (id & 31);
        const float4 x = *reinterpret_cast<const float4 *>(&s.X[i * LS + j]);
        *reinterpret_cast<float4 *>(Bb + (size_t)i * ld + j) =
            make_float4(j <= i ? x.x : 0.0f, j + 1 <= i ? x.y : 0.0f, j + 2 <= i ? x.z : 0.0f, j + 3 <= i ? x.w : 0.0f);
    }
    if (BinvT != nullptr) {
        // inv(L)^T: transposed through shared memory (L's array is free now) so that the global stores are full rows too
        __syncthreads();
#pragma unroll 4
        for (int u = 0; u < 64; ++u) {
            const int id = tid + T4 * u, j = id & 127, i = id >> 7;
            s.L[i * LS + j] = s.X[j * LS + i];
        }
        __syncthreads();
        float *Tb = BinvT + (size_t)k0 * ld + k0;
#pragma unroll 4
        for (int u = 0; u < 16; ++u) {
            const int id = tid + T4 * u, i = id >> 5, j = 4 * (id & 31);
            const float4 x = *reinterpret_cast<const float4 *>(&s.L[i * LS + j]);
            *reinterpret_cast<float4 *>(Tb + (size_t)i * ld + j) =
                make_float4(i <= j ? x.x : 0.0f, i <= j + 1 ? x.y : 0.0f, i <= j + 2 ? x.z : 0.0f, i <= j + 3 ? x.w : 0.0f);
        }
    }
    CD4_LAP(6);
}
}  // namespace cd4
