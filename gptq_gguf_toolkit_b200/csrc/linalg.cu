// linalg.cu -- the dense linear algebra either side of the column loop:
//   gq_hessian_update  (replaces GPTQ.update's H.addmm_, quant/gptq/src/gptq.py:108-112)
//   gq_pre_step        (replaces quantization_pre_step's dead-channel fix, gptq.py:134-141)
//   gq_prepare         (replaces GPTQ._prepare + inv_sym, gptq.py:305-324, linalg_utils.py:8-12)
//
// gq_prepare does NOT mirror the reference's potrf -> potri -> potrf chain (4n^3/3 flops).  With J the
// index reversal, H = R R^T (R upper)  <=>  J H J = L L^T (L lower, L = J R J), and
// chol(inv(H), upper) = R^-1 = J L^-1 J by uniqueness of the Cholesky factor.  So: one blocked Cholesky of
// the reversed matrix plus one blocked triangular inverse (2n^3/3 flops), then a reversed copy.
// The triangular inverse is a log-depth pairwise merge  inv([[A,0],[C,B]]) = [[Ai,0],[-Bi C Ai, Bi]],
// whose work is all GEMM (zero tiles of the triangular factors are skipped).
#include <cstdlib>
#include <vector>

#include "chol_diag_v3.cuh"
#include "chol_diag_v4.cuh"
#include "gemm_f16x3.cuh"
#include "gemm_tf32.cuh"
#include "sgemm.cuh"

namespace {

constexpr int NB = 128;   // panel width
#ifndef GQ_PREPARE_F16_DEFAULT
#define GQ_PREPARE_F16_DEFAULT true
#endif

// ---------------------------------------------------------------------------------------------
// masks and damping
// ---------------------------------------------------------------------------------------------
// W[:, j] = 0 where H[j][j] == 0 (gptq.py:134,141).  Runs BEFORE fix_dead_diag (stream order).
__global__ void zero_dead_cols_kernel(const float *__restrict__ H, float *W, int d_row, int d_col) {
    const long n = (long)d_row * d_col;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int j = (int)(i % d_col);
        if (H[(size_t)j * d_col + j] == 0.0f) W[i] = 0.0f;
    }
}
__global__ void fix_dead_diag_kernel(float *H, int d_col) {   // gptq.py:135
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d_col && H[(size_t)j * d_col + j] == 0.0f) H[(size_t)j * d_col + j] = 1.0f;
}

// nz[j] |= any(W[:, j] != 0)   (gptq.py:308)
__global__ void col_nonzero_kernel(const float *__restrict__ W, int d_row, int d_col, int rows_per_block, int *nz) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d_col) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(d_row, r0 + rows_per_block);
    int any = 0;
    for (int r = r0; r < r1; ++r) any |= (W[(size_t)r * d_col + j] != 0.0f);
    if (any) nz[j] = 1;
}
// H[z,:] = 0, H[:,z] = 0, H[z,z] = 1 for all-zero weight columns z (gptq.py:311-313)
__global__ void mask_zero_cols_kernel(float *H, int n, const int *__restrict__ nz) {
    const long total = (long)n * n;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / n), c = (int)(i % n);
        if (!nz[r] || !nz[c]) H[i] = (r == c) ? 1.0f : 0.0f;
    }
}
// H[i,i] += rel_damp * mean(diag H)  (gptq.py:315-316); one CTA, fixed summation order, double accumulator.
__global__ void __launch_bounds__(1024) damp_kernel(float *H, int n, float rel_damp) {
    __shared__ double part[1024];
    __shared__ float damp_s;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += (double)H[(size_t)i * n + i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) damp_s = __fmul_rn(rel_damp, (float)(part[0] / (double)n));
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024) H[(size_t)i * n + i] = __fadd_rn(H[(size_t)i * n + i], damp_s);
}

// ---------------------------------------------------------------------------------------------
// Cholesky chain
// ---------------------------------------------------------------------------------------------
// A[i][j] = H[n-1-i][n-1-j]   (lower triangle is what the factorisation reads)
__global__ void reverse_copy_kernel(const float *__restrict__ H, float *A, int n) {
    const long total = (long)n * n;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        A[i] = H[total - 1 - i];
}

// NOTE (round 1): three rewrites of this kernel were measured and discarded (profiles/r01_gptq_kernel_notes.md): a
// two-level blocked version with a register/shuffle 32 x 32 factorisation (130-138 us: shuffles under `if (warp == 0)`
// compile to WARPSYNC.COLLECTIVE sequences; run redundantly by all warps it executes 0.5 M warp instructions), and a
// one-warp shared-memory version with rolled loops (200 us: every step pays the full shared-memory round trip).
// This column sweep takes ~129 us per launch and remains the critical path of the chain at n = 4096.
// Factor the (128 x 128) diagonal block at A[k0:k0+128, k0:k0+128] in shared memory: L_kk (written back to A)
// and inv(L_kk) (written to Binv, the level-0 blocks of the triangular inverse; upper part zeroed).
constexpr int DT = 1024;
struct DiagSmem { float L[NB][NB + 1]; float X[NB][NB + 1]; float d[NB]; };
__global__ void __launch_bounds__(DT) chol_diag_kernel(float *A, float *Binv, float *BinvT, long ld, int k0, int *not_pd) {
    extern __shared__ __align__(16) uint8_t raw[];
    DiagSmem &s = *reinterpret_cast<DiagSmem *>(raw);
    const int tid = threadIdx.x;
    float *Ab = A + (size_t)k0 * ld + k0;
    float *Bb = Binv + (size_t)k0 * ld + k0;
    for (int id = tid; id < NB * NB; id += DT) {
        const int i = id >> 7, j = id & 127;
        s.L[i][j] = (j <= i) ? Ab[(size_t)i * ld + j] : 0.0f;
    }
    for (int id = tid; id < NB * NB; id += DT) s.X[id >> 7][id & 127] = ((id >> 7) == (id & 127)) ? 1.0f : 0.0f;
    __syncthreads();
    // Column sweep, two barriers per column.  Right-looking Cholesky of L fused with the forward elimination
    // that turns X = I into inv(L):  X_j /= l_jj, then X_i -= l_ij X_j for i > j.
    for (int j = 0; j < NB; ++j) {
        __syncthreads();                                   // trailing updates of column j-1 are visible
        float piv = s.L[j][j];                             // stays un-rooted in smem; the root goes to s.d[j]
        const bool bad = !(piv > 0.0f) || !isfinite(piv);
        if (bad) piv = 1.0f;
        const float dj = __fsqrt_rn(piv), inv = __frcp_rn(dj);
        if (tid == 0) {
            s.d[j] = dj;
            if (bad) *not_pd = 1;
        }
        for (int i = j + 1 + tid; i < NB; i += DT) s.L[i][j] = __fmul_rn(s.L[i][j], inv);
        for (int c = tid; c <= j; c += DT) s.X[j][c] = __fmul_rn(s.X[j][c], inv);
        __syncthreads();
        // Rows below j: columns c <= j belong to X (X_i -= l_ij X_j), columns j < c <= i to L (L_ic -= l_ij l_cj).
        // A thread owns one column and every RG-th row; its column factor is loaded once, and the rows are
        // processed four at a time with all loads ahead of the stores (the chain is latency bound otherwise).
        {
            constexpr int RG = DT / NB;
            const int c = tid & (NB - 1), rg = tid >> 7;
            const bool isX = (c <= j);
            const float colv = isX ? s.X[j][c] : s.L[c][j];
            float (*T)[NB + 1] = isX ? s.X : s.L;
            for (int i0 = j + 1 + rg; i0 < NB; i0 += 4 * RG) {
                float lij[4], t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * RG;
                    if (i < NB) { lij[u] = s.L[i][j]; t[u] = T[i][c]; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * RG;
                    if (i < NB && (isX || c <= i)) T[i][c] = __fmaf_rn(-lij[u], colv, t[u]);
                }
            }
        }
    }
    __syncthreads();
    for (int id = tid; id < NB * NB; id += DT) {
        const int i = id >> 7, j = id & 127;
        if (j <= i) Ab[(size_t)i * ld + j] = (i == j) ? s.d[i] : s.L[i][j];
        Bb[(size_t)i * ld + j] = s.X[i][j];
        if (BinvT != nullptr) BinvT[(size_t)(k0 + i) * ld + k0 + j] = s.X[j][i];   // inv(L_kk)^T, upper triangular
    }
}

#include "chol_diag_v2.cuh"      // chol_diag_v2_kernel (the default diagonal-block kernel)

// dst[b][c][r] = src[b][r][c]  (rows x cols block of a strided matrix -> cols x rows block of another), 32x32 tiles
__global__ void __launch_bounds__(256) transpose_block_kernel(const float *__restrict__ src, long ld_src, long bs_src,
                                                              float *__restrict__ dst, long ld_dst, long bs_dst, int rows, int cols) {
    __shared__ float t[32][33];
    const float *S = src + (long)blockIdx.z * bs_src;
    float *D = dst + (long)blockIdx.z * bs_dst;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) t[r][tx] = S[(long)(r0 + r) * ld_src + c0 + tx];
    __syncthreads();
    for (int c = ty; c < 32; c += 8) D[(long)(c0 + c) * ld_dst + r0 + tx] = t[tx][c];
}

// U[i][j] = Linv[n-1-i][n-1-j] for j >= i, 0 below; identity if the factorisation failed (gptq.py:321-323).
__global__ void finish_u_kernel(const float *__restrict__ Linv, float *U, int n, const int *__restrict__ not_pd) {
    const long total = (long)n * n;
    const int bad = *not_pd;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / n), j = (int)(idx % n);
        float v;
        if (bad) v = (i == j) ? 1.0f : 0.0f;
        else v = (j >= i) ? Linv[total - 1 - idx] : 0.0f;
        U[idx] = v;
    }
}
__global__ void copy_flag_kernel(const int *src, int *dst) { *dst = *src; }

int ew_grid(long n) {
    long g = (n + 255) / 256;
    const long cap = 148L * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

sg::Args gemm_args(const float *A, long a_rs, long a_cs, const float *B, long b_rs, long b_cs, float *C, long ldc,
                   int M, int N, int K, float alpha, float beta, int tile_mode, int k_mode) {
    sg::Args a;
    a.A = A; a.B = B; a.C = C; a.a_rs = a_rs; a.a_cs = a_cs; a.b_rs = b_rs; a.b_cs = b_cs; a.ldc = ldc;
    a.a_batch = a.b_batch = a.c_batch = 0;
    a.M = M; a.N = N; a.K = K; a.alpha = alpha; a.beta = beta; a.tile_mode = tile_mode; a.k_mode = k_mode;
    a.in_dtype = GQ_F32;
    return a;
}

}  // namespace

// =============================================================================================
// tcgen05 path (hessian_tc.cu)
size_t gq_hessian_tc_workspace_bytes(long n_tok, int d_col);
bool gq_hessian_tc_supported(long n_tok, int d_col, int x_dtype);
int gq_hessian_tc(float *H, const void *X, long n_tok, int d_col, int x_dtype, float beta, float alpha, void *workspace,
                  size_t ws_bytes, cudaStream_t st);

extern "C" size_t gq_hessian_workspace_bytes(long n_tok, int d_col, int x_dtype) {
    return gq_hessian_tc_supported(n_tok, d_col, x_dtype) ? gq_hessian_tc_workspace_bytes(n_tok, d_col) : 0;
}

extern "C" int gq_hessian_update(float *H, const void *X, long n_tok, int d_col, int x_dtype, float beta, float alpha,
                                 void *workspace, size_t ws_bytes, gq_stream_t stream) {
    GQ_REQUIRE(H && X, "gq_hessian_update: null pointer");
    GQ_REQUIRE(d_col > 0 && d_col % 128 == 0, "gq_hessian_update: d_col=%d must be a positive multiple of 128", d_col);
    GQ_REQUIRE(n_tok > 0 && n_tok < (1L << 31), "gq_hessian_update: bad n_tok %ld", n_tok);
    GQ_REQUIRE(x_dtype >= GQ_F32 && x_dtype <= GQ_BF16, "gq_hessian_update: bad x_dtype %d", x_dtype);
    GQ_REQUIRE(((uintptr_t)H | (uintptr_t)X) % 16 == 0, "gq_hessian_update: H and X must be 16-byte aligned");
    if (gq_hessian_tc_supported(n_tok, d_col, x_dtype))     // 16-bit activations: tcgen05 tensor cores
        return gq_hessian_tc(H, X, n_tok, d_col, x_dtype, beta, alpha, workspace, ws_bytes, (cudaStream_t)stream);
    // fp32 activations (or d_col not a multiple of 256): fp32 SIMT tiles
    sg::Args a;
    a.A = X; a.B = X; a.C = H;
    a.a_rs = 1; a.a_cs = d_col;      // A(m,k) = X[k][m]
    a.b_rs = d_col; a.b_cs = 1;      // B(k,n) = X[k][n]
    a.ldc = d_col; a.a_batch = a.b_batch = a.c_batch = 0;
    a.M = d_col; a.N = d_col; a.K = (int)n_tok;
    a.alpha = alpha; a.beta = beta; a.tile_mode = sg::TM_UPPER_MIRROR; a.k_mode = sg::KM_FULL; a.in_dtype = x_dtype;
    return sg::launch(a, 1, (cudaStream_t)stream);
}

extern "C" int gq_pre_step(float *H, float *W, int d_row, int d_col, gq_stream_t stream) {
    GQ_REQUIRE(H && W && d_row > 0 && d_col > 0, "gq_pre_step: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    zero_dead_cols_kernel<<<ew_grid((long)d_row * d_col), 256, 0, st>>>(H, W, d_row, d_col);
    fix_dead_diag_kernel<<<(d_col + 255) / 256, 256, 0, st>>>(H, d_col);
    gq_count_launches(2);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

// workspace: A (n*n) | Linv (n*n) | Linv^T (n*n) | nz (n ints) + flag | 3xTF32 operand splits
namespace {
size_t split_ws_bytes(size_t n) { return (size_t)(1.25 * (double)n * (double)n) * sizeof(float) + (n * 512 + 1024) * sizeof(float) + 8192; }
// Steps per trailing update of the blocked Cholesky (GQ_PREPARE_GROUP = 1, 2, 4 or 8; read on every call).  With G > 1 the
// factorisation is left-looking inside a group of G block columns (a block column receives the group's earlier columns in ONE
// skinny GEMM right before it is factored) and right-looking between groups (ONE rank-128 G update of everything behind the
// group): the trailing matrix -- 411 MB on average at n = 14336, far beyond L2 -- is read and written n/(128 G) times instead
// of n/128 times, which is what bounded the rank-128 updates (32 flop per byte of C).
int prepare_group() {
    const char *e = getenv("GQ_PREPARE_GROUP");
    const int v = e ? atoi(e) : 4;
    return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 4;
}
int prepare_diag_variant() {      // GQ_DIAG_V2: 0 = chol_diag_kernel, 1 = chol_diag_v2_kernel, 3 = register-resident chol_diag_v3.cuh,
    const char *e = getenv("GQ_DIAG_V2");      // 4 (default) = two-level chol_diag_v4.cuh; read on every call (tests / micro-benchmarks)
    return (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : (e && e[0] == '3') ? 3 : 4;
}
// GQ_PREPARE_GEMM=tf32|f16 (read on every call: tests and micro-benchmarks flip it): which tcgen05 GEMM the chain uses --
// 3xTF32 (gemm_tf32.cu) or split-fp16 (gemm_f16x3.cu: half the operand bytes, twice the MMA rate, same 22-bit operands).
bool prepare_use_f16() {
    const char *e = getenv("GQ_PREPARE_GEMM");
    return e ? (e[0] == 'f') : GQ_PREPARE_F16_DEFAULT;
}
bool prepare_use_simt() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("GQ_PREPARE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
}  // namespace

extern "C" size_t gq_prepare_workspace_bytes(int d_col) {
    const size_t n = (size_t)d_col;
    return 3 * n * n * sizeof(float) + (n + 64) * sizeof(int) + 1024 + split_ws_bytes(n);
}

extern "C" int gq_prepare(float *H, const float *W, int d_row, int d_col, float rel_damp, float *U_out, void *workspace,
                          size_t ws_bytes, int *not_pd_flag, gq_stream_t stream) {
    GQ_REQUIRE(H && W && U_out && workspace, "gq_prepare: null pointer");
    GQ_REQUIRE(d_row > 0 && d_col > 0 && d_col % NB == 0, "gq_prepare: d_col=%d must be a positive multiple of 128", d_col);
    GQ_REQUIRE(((uintptr_t)H | (uintptr_t)U_out | (uintptr_t)workspace) % 16 == 0, "gq_prepare: H, U_out, workspace must be 16-byte aligned");
    if (ws_bytes < gq_prepare_workspace_bytes(d_col)) {
        gq_set_error("gq_prepare: workspace %zu < %zu bytes", ws_bytes, gq_prepare_workspace_bytes(d_col));
        return GQ_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = d_col;
    const long ld = n;
    const bool simt = prepare_use_simt(), use_f16 = prepare_use_f16();
    float *A = (float *)workspace;
    float *Li = A + (size_t)n * n;
    float *LiT = Li + (size_t)n * n;
    int *nz = (int *)(LiT + (size_t)n * n);
    int *flag = nz + n;
    void *sws = (void *)(((uintptr_t)(nz + n + 64) + 1023) & ~(uintptr_t)1023);
    const size_t sws_bytes = split_ws_bytes((size_t)n);

    // --- masks + damping (gptq.py:308-316) ---
    GQ_CHECK_CUDA(cudaMemsetAsync(nz, 0, (size_t)(n + 64) * sizeof(int), st));
    {
        const int rpb = 256;
        dim3 g((n + 255) / 256, (d_row + rpb - 1) / rpb);
        col_nonzero_kernel<<<g, 256, 0, st>>>(W, d_row, n, rpb, nz);
    }
    mask_zero_cols_kernel<<<ew_grid((long)n * n), 256, 0, st>>>(H, n, nz);
    damp_kernel<<<1, 1024, 0, st>>>(H, n, rel_damp);
    gq_count_launches(3);

    // --- L = chol(J H J); inv(L_kk) and its transpose for every diagonal block ---
    reverse_copy_kernel<<<ew_grid((long)n * n), 256, 0, st>>>(H, A, n);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaMemsetAsync(Li, 0, (size_t)n * n * sizeof(float), st));
    if (!simt) GQ_CHECK_CUDA(cudaMemsetAsync(LiT, 0, (size_t)n * n * sizeof(float), st));
    GQ_CHECK_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DiagSmem)));
    GQ_CHECK_CUDA(cudaFuncSetAttribute(chol_diag_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DiagSmem2)));
    const int diag_variant = prepare_diag_variant();
    if (diag_variant == 3)
        GQ_CHECK_CUDA(cudaFuncSetAttribute(cd3::chol_diag_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cd3::Smem3)));
    if (diag_variant == 4)
        GQ_CHECK_CUDA(cudaFuncSetAttribute(cd4::chol_diag_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cd4::Smem4)));
    auto tc_gemm = [&](const float *Ap, const float *Bp, float *Cp, int M, int N, int K, int batch, long ab, long bb, long cb,
                       float alpha, float beta, int tile_mode, int k_mode, bool same) {
        tg::GemmArgs g;
        g.A = Ap; g.lda = ld; g.a_batch = ab; g.B = Bp; g.ldb = ld; g.b_batch = bb; g.C = Cp; g.ldc = ld; g.c_batch = cb;
        g.M = M; g.N = N; g.K = K; g.batch = batch; g.alpha = alpha; g.beta = beta; g.tile_mode = tile_mode; g.k_mode = k_mode;
        g.same_ab = same;
        return use_f16 ? th::gemm_f16x3_nt(g, sws, sws_bytes, st) : tg::gemm_tf32x3_nt(g, sws, sws_bytes, st);
    };
    const int G = simt ? 1 : prepare_group();
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int g0 = (k0 / NB) / G * G * NB;                 // first column of this step's group
        int rc = GQ_OK;
        if (k0 > g0) {
            // left-looking inside the group: block column k0 (diagonal block included) receives the group's earlier columns
            //   A[k0:, k0:k0+128] -= L[k0:, g0:k0] L[k0:k0+128, g0:k0]^T
            const float *Lg = A + (size_t)k0 * ld + g0;
            // (B = the first 128 rows of A: with the fp16 back-end the operand is split once)
            rc = tc_gemm(Lg, Lg, A + (size_t)k0 * ld + k0, n - k0, NB, k0 - g0, 1, 0, 0, 0, -1.0f, 1.0f, tg::TM_FULL, tg::KM_FULL, use_f16);
            if (rc) return rc;
        }
        if (diag_variant == 1) chol_diag_v2_kernel<<<1, DT2, sizeof(DiagSmem2), st>>>(A, Li, simt ? nullptr : LiT, ld, k0, flag);
        else if (diag_variant == 3) cd3::chol_diag_v3_kernel<<<1, cd3::T3, sizeof(cd3::Smem3), st>>>(A, Li, simt ? nullptr : LiT, ld, k0, flag);
        else if (diag_variant == 4) cd4::chol_diag_v4_kernel<<<1, cd4::T4, sizeof(cd4::Smem4), st>>>(A, Li, simt ? nullptr : LiT, ld, k0, flag);
        else chol_diag_kernel<<<1, DT, sizeof(DiagSmem), st>>>(A, Li, simt ? nullptr : LiT, ld, k0, flag);
        gq_count_launches(1);
        const int rem = n - k0 - NB;
        if (rem > 0) {
            float *P = A + (size_t)(k0 + NB) * ld + k0;          // panel (rem x 128)
            const float *Lkk_inv = Li + (size_t)k0 * ld + k0;     // inv(L_kk), row-major lower
            float *T = A + (size_t)(k0 + NB) * ld + (k0 + NB);    // trailing matrix
            if (simt) {
                // P <- P * inv(L_kk)^T : A(m,k) = P[m][k], B(k,n) = Lkk_inv[n][k]
                rc = sg::launch(gemm_args(P, ld, 1, Lkk_inv, 1, ld, P, ld, rem, NB, NB, 1.0f, 0.0f, sg::TM_FULL, sg::KM_FULL), 1, st);
                if (rc) return rc;
                rc = sg::launch(gemm_args(P, ld, 1, P, 1, ld, T, ld, rem, rem, NB, -1.0f, 1.0f, sg::TM_LOWER, sg::KM_FULL), 1, st);
            } else {
                // both GEMMs are NT (K-contiguous operands); operands are copied (split) before C is written => in place is safe
                rc = tc_gemm(P, Lkk_inv, P, rem, NB, NB, 1, 0, 0, 0, 1.0f, 0.0f, tg::TM_FULL, tg::KM_FULL, false);
                if (rc) return rc;
                if (k0 + NB == g0 + G * NB) {
                    // the group is complete: everything behind it gets the group's G block columns at once
                    //   T -= L[ge:, g0:ge] L[ge:, g0:ge]^T      (lower tiles)
                    const float *Lg = A + (size_t)(k0 + NB) * ld + g0;
                    rc = tc_gemm(Lg, Lg, T, rem, rem, k0 + NB - g0, 1, 0, 0, 0, -1.0f, 1.0f, tg::TM_LOWER, tg::KM_FULL, true);
                }
            }
            if (rc) return rc;
        }
    }

    // --- X = inv(L) by pairwise merging of diagonal blocks:  X21 = -X22 * L21 * X11.  U_out is the scratch. ---
    // Tensor path keeps Y = X^T as well so that every product is NT:
    //   T^T[a][b] = sum_k Y11[a][k] L21[b][k]   (Y11 upper: k >= a)      X21[b][a] = -sum_k X22[b][k] T^T[a][k]   (X22 lower: k <= b)
    {
        int nblk = n / NB;
        int *start = new int[nblk + 1];
        for (int i = 0; i <= nblk; ++i) start[i] = i * NB;
        int rc = GQ_OK;
        while (nblk > 1 && rc == GQ_OK) {
            int out = 0;
            if (simt) {
                for (int b = 0; b + 1 < nblk && rc == GQ_OK; b += 2) {
                    const int r1 = start[b], r2 = start[b + 1], r3 = start[b + 2];
                    const int s1 = r2 - r1, s2 = r3 - r2;
                    const float *C = A + (size_t)r2 * ld + r1;
                    const float *Ai = Li + (size_t)r1 * ld + r1;
                    const float *Bi = Li + (size_t)r2 * ld + r2;
                    float *Tm = U_out + (size_t)r2 * ld + r1;
                    float *O = Li + (size_t)r2 * ld + r1;
                    rc = sg::launch(gemm_args(C, ld, 1, Ai, ld, 1, Tm, ld, s2, s1, s1, 1.0f, 0.0f, sg::TM_FULL, sg::KM_FROM_N), 1, st);
                    if (rc == GQ_OK)
                        rc = sg::launch(gemm_args(Bi, ld, 1, Tm, ld, 1, O, ld, s2, s1, s2, -1.0f, 0.0f, sg::TM_FULL, sg::KM_TO_M), 1, st);
                }
            } else {
                // uniform pairs of this level go out as one batched launch, an odd-sized tail pair separately
                const int npairs = nblk / 2;
                int b = 0;
                while (b < npairs && rc == GQ_OK) {
                    const int r1 = start[2 * b], r2 = start[2 * b + 1], r3 = start[2 * b + 2];
                    const int s1 = r2 - r1, s2 = r3 - r2;
                    int cnt = 1;
                    while (b + cnt < npairs && start[2 * (b + cnt) + 1] - start[2 * (b + cnt)] == s1 &&
                           start[2 * (b + cnt) + 2] - start[2 * (b + cnt) + 1] == s2)
                        ++cnt;
                    const long bstride = (long)(s1 + s2) * (ld + 1);          // next pair sits one (s1+s2) block down the diagonal
                    const float *L21 = A + (size_t)r2 * ld + r1;             // (s2 x s1)
                    const float *Y11 = LiT + (size_t)r1 * ld + r1;           // (s1 x s1) upper
                    const float *X22 = Li + (size_t)r2 * ld + r2;            // (s2 x s2) lower
                    float *Tt = U_out + (size_t)r1 * ld + r2;                // (s1 x s2) scratch
                    float *X21 = Li + (size_t)r2 * ld + r1;                  // (s2 x s1)
                    float *Y12 = LiT + (size_t)r1 * ld + r2;                 // (s1 x s2)
                    rc = tc_gemm(Y11, L21, Tt, s1, s2, s1, cnt, bstride, bstride, bstride, 1.0f, 0.0f, tg::TM_FULL, tg::KM_FROM_M, false);
                    if (rc == GQ_OK)
                        rc = tc_gemm(X22, Tt, X21, s2, s1, s2, cnt, bstride, bstride, bstride, -1.0f, 0.0f, tg::TM_FULL, tg::KM_TO_M, false);
                    if (rc == GQ_OK) {
                        dim3 g(s1 / 32, s2 / 32, cnt);
                        transpose_block_kernel<<<g, 256, 0, st>>>(X21, ld, bstride, Y12, ld, bstride, s2, s1);
                        gq_count_launches(1);
                    }
                    b += cnt;
                }
            }
            for (int b = 0; b + 1 < nblk; b += 2) start[out++] = start[b];
            if (nblk & 1) start[out++] = start[nblk - 1];
            start[out] = n;
            nblk = out;
        }
        delete[] start;
        if (rc != GQ_OK) return rc;
    }

    finish_u_kernel<<<ew_grid((long)n * n), 256, 0, st>>>(Li, U_out, n, flag);
    gq_count_launches(not_pd_flag ? 2 : 1);
    if (not_pd_flag) copy_flag_kernel<<<1, 1, 0, st>>>(flag, not_pd_flag);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
