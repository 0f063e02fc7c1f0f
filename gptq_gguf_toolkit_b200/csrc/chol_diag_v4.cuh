// chol_diag_v4.cuh -- two-level diagonal-block kernel of gq_prepare (csrc/linalg.cu), the default since round 2.
// Same contract as chol_diag_v2 / v3: factor A[k0:k0+128, k0:k0+128] = L L^T in place (lower), inv(L) to Binv, inv(L)^T to
// BinvT (may be null); a non-positive pivot sets *not_pd and is replaced by 1.
//
// v2 / v3 sweep the 128 columns one by one with a block-wide barrier per column (88.7 / 70.4 us on B200: ~1000 cycles per
// column, most of it barrier + shared-memory round trips on the dependent chain).  Here the dependent chain lives in ONE warp:
//   * per 32-column panel, warp 0 factors the 32 x 32 diagonal block in registers (lane = row, column broadcasts by shuffle:
//     per column one pivot shuffle, one rsqrt, and 31-j independent shuffle + FMA pairs -- no barrier, no shared memory);
//   * the rows below solve  x L_bb^T = p  by substitution, one thread per row, everything in registers (straight-line code,
//     L_bb read as LDS.128 broadcasts), then the rank-32 trailing update of v2 / v3 (strided 4 x 4 register tiles);
//   * three block-wide barriers per panel instead of 32 + 2;
//   * inv(L) as in v3: the four 32 x 32 diagonal inverses by forward substitution (one thread per column), the off-diagonal
//     blocks from  X_qp = -X_qq sum_r L_qr X_rp.
// The CPU suite runs this source on the SIMT emulator against fp64 (tests/test_simt_emu_cpu.py).
#pragma once

namespace cd4 {
constexpr int NB = 128, LS = 132, PW = 32, T4 = 256;
struct Smem4 { float L[NB * LS]; float X[NB * LS]; float inv[NB]; };

// 1 / sqrt(p) to ~1 ulp: hardware approximation + one Newton step,  y <- y + y (1/2 - (p y / 2) y)
__device__ __forceinline__ float rsqrt_nr(float p) {
    const float y = rsqrtf(p);
    const float h = 0.5f * p * y;
    return fmaf(y, fmaf(-h, y, 0.5f), y);
}

__global__ void __launch_bounds__(T4) chol_diag_v4_kernel(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem4 &s = *reinterpret_cast<Smem4 *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *Ab = A + (size_t)k0 * ld + k0;
    float *Bb = Binv + (size_t)k0 * ld + k0;
    for (int id = tid; id < NB * NB; id += T4) {
        const int i = id >> 7, j = id & 127;
        s.L[i * LS + j] = (j <= i) ? Ab[(size_t)i * ld + j] : 0.0f;
        s.X[i * LS + j] = 0.0f;
    }
    __syncthreads();
    for (int o = 0; o < NB; o += PW) {
        // ---- (a) warp 0: Cholesky of the 32 x 32 diagonal block, lane = row, the row's 32 entries in registers ----
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
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float pj = __shfl_sync(0xffffffffu, a[j], j, 32);
                const bool bad = !(pj > 0.0f) || !isfinite(pj);
                if (bad) pj = 1.0f;
                const float iv = rsqrt_nr(pj);
                if (lane == j) {
                    s.inv[o + j] = iv;
                    if (bad) *not_pd = 1;
                }
                // column j of L: the diagonal is sqrt(p) = p / sqrt(p); lanes above the diagonal carry finite junk that no
                // lane below ever reads (the shuffles below only take from lanes k > j)
                const float lj = (lane == j) ? pj * iv : a[j] * iv;
                a[j] = lj;
#pragma unroll
                for (int k = j + 1; k < 32; ++k) {
                    const float ck = __shfl_sync(0xffffffffu, lj, k, 32);
                    a[k] = fmaf(-lj, ck, a[k]);
                }
            }
            float4 *dst = reinterpret_cast<float4 *>(&s.L[(o + lane) * LS + o]);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(4 * q <= lane ? a[4 * q] : 0.0f, 4 * q + 1 <= lane ? a[4 * q + 1] : 0.0f,
                                     4 * q + 2 <= lane ? a[4 * q + 2] : 0.0f, 4 * q + 3 <= lane ? a[4 * q + 3] : 0.0f);
        }
        __syncthreads();
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
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float acc = x[j];
                    const float4 *lrow = reinterpret_cast<const float4 *>(&s.L[(o + j) * LS + o]);
#pragma unroll
                    for (int q = 0; q < (j + 3) / 4; ++q) {
                        const float4 l4 = lrow[q];
                        if (4 * q < j) acc = fmaf(-x[4 * q], l4.x, acc);
                        if (4 * q + 1 < j) acc = fmaf(-x[4 * q + 1], l4.y, acc);
                        if (4 * q + 2 < j) acc = fmaf(-x[4 * q + 2], l4.z, acc);
                        if (4 * q + 3 < j) acc = fmaf(-x[4 * q + 3], l4.w, acc);
                    }
                    x[j] = acc * s.inv[o + j];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) row[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            }
            __syncthreads();
            // ---- (c) rank-32 update of the trailing part of L (strided 4 x 4 tiles, see chol_diag_v2_kernel) ----
            const int nt = R / 4;
            for (int t = tid; t < nt * nt; t += T4) {
                const int ti = t / nt, tc = t - ti * nt;
                const int i0 = lo + ti, c0 = lo + tc;
                float acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0f;
#pragma unroll 2
                for (int k = o; k < lo; k += 4) {
                    float4 av[4], bv[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) av[x] = *reinterpret_cast<const float4 *>(&s.L[(i0 + nt * x) * LS + k]);
#pragma unroll
                    for (int y = 0; y < 4; ++y) bv[y] = *reinterpret_cast<const float4 *>(&s.L[(c0 + nt * y) * LS + k]);
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
                        const int i = i0 + nt * x, c = c0 + nt * y;
                        if (c <= i) s.L[i * LS + c] -= acc[x][y];
                    }
            }
            __syncthreads();
        }
    }
    // ---- X = inv(L): the four diagonal 32 x 32 blocks by forward substitution, one thread per column, registers only
    //      (straight-line code over all 32 rows; the entries above the column's diagonal come out as exact zeros) ----
    if (tid < NB) {
        const int o = tid & ~31, cc = tid & 31;
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float acc = (i == cc) ? 1.0f : 0.0f;
            const float4 *lrow = reinterpret_cast<const float4 *>(&s.L[(o + i) * LS + o]);
#pragma unroll
            for (int q = 0; q < (i + 3) / 4; ++q) {
                const float4 l4 = lrow[q];
                if (4 * q < i) acc = fmaf(-l4.x, x[4 * q], acc);
                if (4 * q + 1 < i) acc = fmaf(-l4.y, x[4 * q + 1], acc);
                if (4 * q + 2 < i) acc = fmaf(-l4.z, x[4 * q + 2], acc);
                if (4 * q + 3 < i) acc = fmaf(-l4.w, x[4 * q + 3], acc);
            }
            x[i] = (i >= cc) ? acc * s.inv[o + i] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) s.X[(o + i) * LS + o + cc] = x[i];
    }
    __syncthreads();
    // ---- off-diagonal blocks, block diagonal d = 1..3:  T = sum_{r=p}^{q-1} L_qr X_rp  (scratch: the unused upper block
    //      (p, q) of X),  X_qp = -X_qq T ----
    for (int d = 1; d < 4; ++d) {
        const int nblk = 4 - d;
        for (int it = tid; it < nblk * 256; it += T4) {
            const int p = it >> 8, q = p + d, i = (it >> 3) & 31, cg = it & 7;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int k = 32 * p; k < 32 * q; k += 4) {
                const float4 l4 = *reinterpret_cast<const float4 *>(&s.L[(32 * q + i) * LS + k]);
                const float lk[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 x4 = *reinterpret_cast<const float4 *>(&s.X[(k + kk) * LS + 32 * p + 4 * cg]);
                    acc[0] = fmaf(lk[kk], x4.x, acc[0]); acc[1] = fmaf(lk[kk], x4.y, acc[1]);
                    acc[2] = fmaf(lk[kk], x4.z, acc[2]); acc[3] = fmaf(lk[kk], x4.w, acc[3]);
                }
            }
            *reinterpret_cast<float4 *>(&s.X[(32 * p + i) * LS + 32 * q + 4 * cg]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        for (int it = tid; it < nblk * 256; it += T4) {
            const int p = it >> 8, q = p + d, i = (it >> 3) & 31, cg = it & 7;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int k = 0; k < 32; k += 4) {
                const float4 l4 = *reinterpret_cast<const float4 *>(&s.X[(32 * q + i) * LS + 32 * q + k]);   // X_qq, lower
                const float lk[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(&s.X[(32 * p + k + kk) * LS + 32 * q + 4 * cg]);
                    acc[0] = fmaf(lk[kk], t4.x, acc[0]); acc[1] = fmaf(lk[kk], t4.y, acc[1]);
                    acc[2] = fmaf(lk[kk], t4.z, acc[2]); acc[3] = fmaf(lk[kk], t4.w, acc[3]);
                }
            }
            *reinterpret_cast<float4 *>(&s.X[(32 * q + i) * LS + 32 * p + 4 * cg]) = make_float4(-acc[0], -acc[1], -acc[2], -acc[3]);
        }
        __syncthreads();
    }
    for (int id = tid; id < NB * NB; id += T4) {
        const int i = id >> 7, j = id & 127;
        if (j <= i) Ab[(size_t)i * ld + j] = s.L[i * LS + j];
        Bb[(size_t)i * ld + j] = (j <= i) ? s.X[i * LS + j] : 0.0f;
        if (BinvT != nullptr) BinvT[(size_t)(k0 + i) * ld + k0 + j] = (i <= j) ? s.X[j * LS + i] : 0.0f;
    }
}
}  // namespace cd4
