// chol_diag_v2.cuh -- the shipped diagonal-block kernel of gq_prepare (csrc/linalg.cu), in a header of its own so that the CPU
// suite can run it on the SIMT emulator (tests/test_simt_emu_cpu.py).  Included INSIDE linalg.cu's anonymous namespace, after
// `constexpr int NB = 128;` -- it is not a stand-alone header.
// ---------------------------------------------------------------------------------------------
// chol_diag_v2_kernel -- same contract as chol_diag_kernel (factor the 128 x 128 diagonal block, invert the factor),
// restructured so that the per-column critical path is ONE barrier + pivot + a short update:
//   * 32-column panels: inside a panel the rank-1 updates touch only the panel's columns of L (all rows below) and the
//     panel's own rows of X; everything right of / below the panel gets ONE rank-32 update per panel from 4 x 4 register
//     tiles (LDS.128 along k) instead of 32 rank-1 sweeps through shared memory;
//   * the scaling of column `col` (L[:, col] / l_cc, X[col, :] / l_cc) is folded into the update's operands and applied
//     to the stored values once per panel, which removes the second barrier of every column;
//   * all loops are rolled (a 128 x 128 factorisation executes each instruction once per launch: straight-line code is
//     instruction-fetch bound, profiles/r01_gptq_kernel_notes.md).
// Selected by GQ_DIAG_V2 (default on); GQ_DIAG_V2=0 falls back to chol_diag_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int DT2 = 512;
constexpr int LS = 132;    // shared-memory row stride in floats (multiple of 4: rows are read with LDS.128)
constexpr int PW = 32;     // panel width
struct DiagSmem2 { float L[NB * LS]; float X[NB * LS]; float inv[NB]; float d[NB]; };

__global__ void __launch_bounds__(DT2) chol_diag_v2_kernel(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    extern __shared__ __align__(16) uint8_t raw[];
    DiagSmem2 &s = *reinterpret_cast<DiagSmem2 *>(raw);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    constexpr int NW = DT2 / 32;
    float *Ab = A + (size_t)k0 * ld + k0;
    float *Bb = Binv + (size_t)k0 * ld + k0;
    for (int id = tid; id < NB * NB; id += DT2) {
        const int i = id >> 7, j = id & 127;
        s.L[i * LS + j] = (j <= i) ? Ab[(size_t)i * ld + j] : 0.0f;
        s.X[i * LS + j] = (i == j) ? 1.0f : 0.0f;
    }
    for (int base = 0; base < NB; base += PW) {
        // ---- (a) column sweep inside the panel ----
#pragma unroll 1
        for (int j = 0; j < PW; ++j) {
            const int col = base + j;
            __syncthreads();                                   // updates of column col-1 (or the previous panel) are visible
            float piv = s.L[col * LS + col];
            const bool bad = !(piv > 0.0f) || !isfinite(piv);
            if (bad) piv = 1.0f;
            const float dj = __fsqrt_rn(piv), iv = __frcp_rn(dj);
            if (tid == 0) {
                s.d[col] = dj;
                s.inv[col] = iv;
                if (bad) *not_pd = 1;
            }
            // Warps 0..7 update L, warps 8..15 update X (every warp executes one region's instructions only: with 16
            // warps the sweep is issue-bound, not latency-bound, if each warp walks through predicated-off code).
            // All loads of a trip are issued before its first store: the compiler cannot move a shared-memory load
            // across a store that might alias, and a load -> FMA -> store chain per element would serialise the sweep.
            if (w < NW / 2) {
                // L: rows below col, the panel's columns right of col:  L[i][c] -= l_i,col * l_c,col   (operands scaled on the fly)
                const int c = base + lane;
                if (c > col) {
                    const float lc = __fmul_rn(s.L[c * LS + col], iv);
#pragma unroll 1
                    for (int i0 = col + 1 + w; i0 < NB; i0 += 2 * NW) {      // rows i0, i0+8, i0+16, i0+24
                        float li[4], t[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + (NW / 2) * u;
                            if (i < NB && c <= i) { li[u] = s.L[i * LS + col]; t[u] = s.L[i * LS + c]; }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + (NW / 2) * u;
                            if (i < NB && c <= i) s.L[i * LS + c] = __fmaf_rn(-__fmul_rn(li[u], iv), lc, t[u]);
                        }
                    }
                }
            } else {
                // X: the panel's rows below col:  X[i][0..col] -= l_i,col * x_col,:
                const int w2 = w - NW / 2, nq = (col >> 5) + 1;
                float xs[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = lane + 32 * q;
                    xs[q] = (q < nq && cc <= col) ? __fmul_rn(s.X[col * LS + cc], iv) : 0.0f;
                }
#pragma unroll 1
                for (int i0 = col + 1 + w2; i0 < base + PW; i0 += NW) {        // rows i0, i0+8
                    float xi[2][4], li[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int i = i0 + (NW / 2) * r;
                        if (i < base + PW) {
                            li[r] = s.L[i * LS + col];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int cc = lane + 32 * q;
                                if (q < nq && cc <= col) xi[r][q] = s.X[i * LS + cc];
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int i = i0 + (NW / 2) * r;
                        if (i < base + PW) {
                            const float l = __fmul_rn(li[r], iv);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int cc = lane + 32 * q;
                                if (q < nq && cc <= col) s.X[i * LS + cc] = __fmaf_rn(-l, xs[q], xi[r][q]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- scale the panel's stored values: L[:, col] and X[col, :] by 1 / l_col,col ----
        for (int id = tid; id < NB * PW; id += DT2) {
            const int i = id >> 5, c = base + (id & 31);
            if (i > c) s.L[i * LS + c] = __fmul_rn(s.L[i * LS + c], s.inv[c]);
            else if (i == c) s.L[i * LS + c] = s.d[c];
        }
        for (int id = tid; id < PW * NB; id += DT2) {
            const int r = base + (id >> 7), cc = id & 127;
            if (cc <= r) s.X[r * LS + cc] = __fmul_rn(s.X[r * LS + cc], s.inv[r]);
        }
        __syncthreads();
        // ---- (b) rank-32 updates with the finished panel (4 x 4 register tiles) ----
        // A tile takes every nt-th row (and, for L, every nt-th column) of the trailing part, so that the lanes of a warp
        // (consecutive tc) read CONSECUTIVE rows with LDS.128: the row stride of 132 floats spreads eight consecutive rows
        // over all 32 banks.  (Blocked 4 x 4 tiles put the lanes 4 rows = 16 banks apart: a 16-way conflict on every load.)
        // L tiles near the diagonal also produce entries above it; the upper part of L is scratch and never read.
        const int lo = base + PW, R = NB - lo;
        if (R > 0) {
            const int nt = R / 4, nLt = nt * nt, nXc = lo / 4, nXt = nt * nXc;
            for (int t = tid; t < nLt + nXt; t += DT2) {
                const bool isX = t >= nLt;
                int ti, tc;
                if (!isX) { ti = t / nt; tc = t - ti * nt; }
                else { const int u = t - nLt; ti = u / nXc; tc = u - ti * nXc; }
                const int i0 = lo + ti;                       // rows i0 + nt * r
                const int c0 = isX ? 4 * tc : lo + tc;        // X: columns c0 .. c0+3;  L: columns c0 + nt * q
                float acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[r][q] = 0.0f;
#pragma unroll 2
                for (int k = base; k < lo; k += 4) {
                    float4 a[4], b[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4 *>(&s.L[(i0 + nt * r) * LS + k]);
                    if (!isX) {      // b[q] = L[c0 + nt*q][k..k+3]
#pragma unroll
                        for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4 *>(&s.L[(c0 + nt * q) * LS + k]);
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                acc[r][q] = __fmaf_rn(a[r].x, b[q].x, acc[r][q]);
                                acc[r][q] = __fmaf_rn(a[r].y, b[q].y, acc[r][q]);
                                acc[r][q] = __fmaf_rn(a[r].z, b[q].z, acc[r][q]);
                                acc[r][q] = __fmaf_rn(a[r].w, b[q].w, acc[r][q]);
                            }
                    } else {         // b[kk] = X[k+kk][c0..c0+3]
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4 *>(&s.X[(k + kk) * LS + c0]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            acc[r][0] = __fmaf_rn(a[r].x, b[0].x, acc[r][0]); acc[r][1] = __fmaf_rn(a[r].x, b[0].y, acc[r][1]);
                            acc[r][2] = __fmaf_rn(a[r].x, b[0].z, acc[r][2]); acc[r][3] = __fmaf_rn(a[r].x, b[0].w, acc[r][3]);
                            acc[r][0] = __fmaf_rn(a[r].y, b[1].x, acc[r][0]); acc[r][1] = __fmaf_rn(a[r].y, b[1].y, acc[r][1]);
                            acc[r][2] = __fmaf_rn(a[r].y, b[1].z, acc[r][2]); acc[r][3] = __fmaf_rn(a[r].y, b[1].w, acc[r][3]);
                            acc[r][0] = __fmaf_rn(a[r].z, b[2].x, acc[r][0]); acc[r][1] = __fmaf_rn(a[r].z, b[2].y, acc[r][1]);
                            acc[r][2] = __fmaf_rn(a[r].z, b[2].z, acc[r][2]); acc[r][3] = __fmaf_rn(a[r].z, b[2].w, acc[r][3]);
                            acc[r][0] = __fmaf_rn(a[r].w, b[3].x, acc[r][0]); acc[r][1] = __fmaf_rn(a[r].w, b[3].y, acc[r][1]);
                            acc[r][2] = __fmaf_rn(a[r].w, b[3].z, acc[r][2]); acc[r][3] = __fmaf_rn(a[r].w, b[3].w, acc[r][3]);
                        }
                    }
                }
                if (isX) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        float4 v = *reinterpret_cast<float4 *>(&s.X[(i0 + nt * r) * LS + c0]);
                        v.x -= acc[r][0]; v.y -= acc[r][1]; v.z -= acc[r][2]; v.w -= acc[r][3];
                        *reinterpret_cast<float4 *>(&s.X[(i0 + nt * r) * LS + c0]) = v;
                    }
                } else {
                    float v[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[r][q] = s.L[(i0 + nt * r) * LS + c0 + nt * q];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int q = 0; q < 4; ++q) s.L[(i0 + nt * r) * LS + c0 + nt * q] = v[r][q] - acc[r][q];
                }
            }
        }
        // the next panel's first __syncthreads() publishes these updates
    }
    __syncthreads();
    for (int id = tid; id < NB * NB; id += DT2) {
        const int i = id >> 7, j = id & 127;
        if (j <= i) Ab[(size_t)i * ld + j] = s.L[i * LS + j];
        Bb[(size_t)i * ld + j] = (j <= i) ? s.X[i * LS + j] : 0.0f;
        if (BinvT != nullptr) BinvT[(size_t)(k0 + i) * ld + k0 + j] = (i <= j) ? s.X[j * LS + i] : 0.0f;   // inv(L_kk)^T, upper
    }
}

