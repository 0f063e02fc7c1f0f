// chol_diag_v3.cuh -- register-resident variant (GQ_DIAG_V2=3; superseded as the default by chol_diag_v4.cuh) of the (128 x 128) diagonal-block kernel of gq_prepare
// (csrc/linalg.cu: chol_diag_v2_kernel, 88.7 us on B200, issue-bound): 70.4 us on B200 (round 2), meets the same accuracy bounds
// (tests/test_gpu_parity.py); also run on the SIMT emulator of the CPU suite (tests/test_simt_emu_cpu.py).
// Same contract as chol_diag_v2_kernel: factor A[k0:k0+128, k0:k0+128] = L L^T in place (lower), inv(L) to Binv, inv(L)^T to
// BinvT (may be null).
//   * inside a 32-column panel every thread keeps ITS 16 panel entries of one row in registers (256 threads = 128 rows x 2
//     halves); per column only the column itself goes through shared memory (double-buffered, ONE barrier per column), and
//     the j-loop is unrolled so that the register indices are compile-time (reused by the 4 panels);
//   * inv(L) is not carried through the sweep: the four 32 x 32 diagonal blocks are inverted by forward substitution (one
//     thread per column), the off-diagonal blocks follow from X_qp = -X_qq * sum_r L_qr X_rp, block diagonal by block
//     diagonal, with LDS.128 along k and 4 outputs per thread;
//   * the rank-32 trailing update of L is v2's (conflict-free strided 4 x 4 tiles).
#pragma once

namespace cd3 {
constexpr int NB = 128, LS = 132, PW = 32, T3 = 256;
struct Smem3 { float L[NB * LS]; float X[NB * LS]; float col[2][NB]; float inv[NB]; float d[NB]; };

__global__ void __launch_bounds__(T3) chol_diag_v3_kernel(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem3 &s = *reinterpret_cast<Smem3 *>(raw);
    const int tid = threadIdx.x;
    float *Ab = A + (size_t)k0 * ld + k0;
    float *Bb = Binv + (size_t)k0 * ld + k0;
    for (int id = tid; id < NB * NB; id += T3) {
        const int i = id >> 7, j = id & 127;
        s.L[i * LS + j] = (j <= i) ? Ab[(size_t)i * ld + j] : 0.0f;
        s.X[i * LS + j] = 0.0f;
    }
    __syncthreads();
    const int r = tid & 127, hf = tid >> 7;
    for (int base = 0; base < NB; base += PW) {
        // ---- (a) column sweep of the panel, the thread's 16 entries of row r in registers ----
        float a[16];
        {
            const float4 *src = reinterpret_cast<const float4 *>(&s.L[r * LS + base + 16 * hf]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = src[q];
                a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
            }
        }
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const int col = base + j;
            float *cb = s.col[j & 1];
            if (hf == (j >> 4)) cb[r] = a[j & 15];              // column `col`, unscaled (rows above the panel: never read)
            __syncthreads();
            float piv = cb[col];
            const bool bad = !(piv > 0.0f) || !isfinite(piv);
            if (bad) piv = 1.0f;
            const float dj = sqrtf(piv), iv = 1.0f / dj;
            if (tid == 0) {
                s.d[col] = dj;
                s.inv[col] = iv;
                if (bad) *not_pd = 1;
            }
            if (r > col) {
                const float li = cb[r] * iv;
                const float4 *lc4 = reinterpret_cast<const float4 *>(cb + base + 16 * hf);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = lc4[q];
                    const float lc[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 4 * q + e;                // own column index; global column gc = base + 16 hf + c
                        if (16 * hf + c > j && base + 16 * hf + c <= r) a[c] = fmaf(-li, lc[e] * iv, a[c]);
                    }
                }
            }
        }
        if (r >= base) {
            float4 *dst = reinterpret_cast<float4 *>(&s.L[r * LS + base + 16 * hf]);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
        }
        __syncthreads();
        // ---- scale the panel's columns: L[:, c] / l_cc, diagonal = l_cc ----
        for (int id = tid; id < NB * PW; id += T3) {
            const int i = id >> 5, c = base + (id & 31);
            if (i > c) s.L[i * LS + c] *= s.inv[c];
            else if (i == c) s.L[i * LS + c] = s.d[c];
        }
        __syncthreads();
        // ---- (b) rank-32 update of the trailing part of L (strided 4 x 4 tiles, see chol_diag_v2_kernel) ----
        const int lo = base + PW, R = NB - lo;
        if (R > 0) {
            const int nt = R / 4;
            for (int t = tid; t < nt * nt; t += T3) {
                const int ti = t / nt, tc = t - ti * nt;
                const int i0 = lo + ti, c0 = lo + tc;
                float acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0f;
#pragma unroll 2
                for (int k = base; k < lo; k += 4) {
                    float4 av[4], bv[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) av[x] = *reinterpret_cast<const float4 *>(&s.L[(i0 + nt * x) * LS + k]);
#pragma unroll
                    for (int y = 0; y < 4; ++y) bv[y] = *reinterpret_cast<const float4 *>(&s.L[(c0 + nt * y) * LS + k]);
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y)
                            acc[x][y] += av[x].x * bv[y].x + av[x].y * bv[y].y + av[x].z * bv[y].z + av[x].w * bv[y].w;
                }
                float v[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) v[x][y] = s.L[(i0 + nt * x) * LS + c0 + nt * y];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) s.L[(i0 + nt * x) * LS + c0 + nt * y] = v[x][y] - acc[x][y];
            }
            __syncthreads();
        }
    }
    // ---- X = inv(L): the four diagonal 32 x 32 blocks by forward substitution, one thread per column ----
    if (tid < NB) {
        const int o = tid & ~31, cc = tid & 31;
        s.X[(o + cc) * LS + o + cc] = s.inv[o + cc];
        for (int i = cc + 1; i < 32; ++i) {
            float acc = 0.0f;
            for (int k = cc; k < i; ++k) acc = fmaf(s.L[(o + i) * LS + o + k], s.X[(o + k) * LS + o + cc], acc);
            s.X[(o + i) * LS + o + cc] = -acc * s.inv[o + i];
        }
    }
    __syncthreads();
    // ---- off-diagonal blocks, block diagonal d = 1..3:  T = sum_{r=p}^{q-1} L_qr X_rp  (scratch: the unused upper block
    //      (p, q) of X),  X_qp = -X_qq T ----
    for (int d = 1; d < 4; ++d) {
        const int nblk = 4 - d;
        for (int it = tid; it < nblk * 256; it += T3) {
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
        for (int it = tid; it < nblk * 256; it += T3) {
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
    for (int id = tid; id < NB * NB; id += T3) {
        const int i = id >> 7, j = id & 127;
        if (j <= i) Ab[(size_t)i * ld + j] = s.L[i * LS + j];
        Bb[(size_t)i * ld + j] = (j <= i) ? s.X[i * LS + j] : 0.0f;
        if (BinvT != nullptr) BinvT[(size_t)(k0 + i) * ld + k0 + j] = (i <= j) ? s.X[j * LS + i] : 0.0f;
    }
}
}  // namespace cd3
