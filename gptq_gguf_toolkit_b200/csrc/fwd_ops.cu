// fwd_ops.cu -- fused element-wise pieces of the calibration block forwards.
//
// The two forward passes per transformer block are outside the reference's hot path (they are the caller's
// model.forward), but once the hot path is fast they are 60 % of the wall-clock, and in HF's eager Llama about a third
// of THAT is memory-bound element-wise kernels: RMSNorm is 6 launches (to fp32, pow, mean, add+rsqrt, mul, to bf16, mul),
// the rotary embedding 10 (mul, slice, neg, cat, mul, add for q and k), SiLU*up 2.  Each entry point below does one of
// them in ONE pass over HBM with HF's own rounding points:
//   gq_fwd_silu_mul : out = rn16( float(rn16(silu(float g))) * float(u) )                                   bit-identical
//   gq_fwd_rope     : out = rn16( float(rn16(x*cos)) + float(rn16(rotate_half(x)*sin)) )                    bit-identical
//   gq_fwd_rmsnorm  : out = rn16( float(w) * float(rn16(float(x) * rsqrtf(mean(x^2) + eps))) )
//                     (the fp32 mean is summed in a different order than torch's reduction: last-bit differences of the
//                      variance, i.e. at most rare 1-ulp differences of the 16-bit result)
// 16-bit activations only (bf16 / fp16); the host wrapper (fused_forward.py) checks each of them against the module it
// replaces before using it and keeps HF's implementation otherwise.
#include "common.cuh"
#include "../../include/gq_fwd.h"

namespace {

template <int DT> struct H16;
template <> struct H16<GQ_BF16> {
    using T = __nv_bfloat16;
    static __device__ __forceinline__ float up(T v) { return __bfloat162float(v); }
    static __device__ __forceinline__ T down(float v) { return __float2bfloat16_rn(v); }
};
template <> struct H16<GQ_F16> {
    using T = __half;
    static __device__ __forceinline__ float up(T v) { return __half2float(v); }
    static __device__ __forceinline__ T down(float v) { return __float2half_rn(v); }
};

template <int DT> union Vec8 {
    uint4 u;
    typename H16<DT>::T h[8];
};

// ---- RMSNorm: one CTA of 256 threads per row, rows of up to 256*8*4 = 8192 elements kept in registers ----------
template <int DT>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const void *__restrict__ x, const void *__restrict__ w, void *__restrict__ out,
                                                      int dim, float eps, float inv_dim) {
    using T = typename H16<DT>::T;
    const T *xr = (const T *)x + (size_t)blockIdx.x * dim;
    T *orow = (T *)out + (size_t)blockIdx.x * dim;
    const int nv = dim / 8;
    Vec8<DT> v[4];
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + 256 * i;
        if (idx < nv) {
            v[i].u = reinterpret_cast<const uint4 *>(xr)[idx];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float f = H16<DT>::up(v[i].h[e]);
                ss = __fmaf_rn(f, f, ss);
            }
        }
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float r = rsqrtf(__fadd_rn(__fmul_rn(tot, inv_dim), eps));       // torch mean = sum * (1/N)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + 256 * i;
        if (idx < nv) {
            Vec8<DT> wv, o;
            wv.u = reinterpret_cast<const uint4 *>(w)[idx];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const T n = H16<DT>::down(__fmul_rn(H16<DT>::up(v[i].h[e]), r));
                o.h[e] = H16<DT>::down(__fmul_rn(H16<DT>::up(wv.h[e]), H16<DT>::up(n)));
            }
            reinterpret_cast<uint4 *>(orow)[idx] = o.u;
        }
    }
}

// ---- SiLU(gate) * up ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) silu_mul_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ u, uint4 *__restrict__ out,
                                                       long n8) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        Vec8<DT> a, b, o;
        a.u = g[i];
        b.u = u[i];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float x = H16<DT>::up(a.h[e]);
            const float s = __fdiv_rn(x, __fadd_rn(1.0f, expf(-x)));            // torch: x / (1 + exp(-x)) in fp32
            o.h[e] = H16<DT>::down(__fmul_rn(H16<DT>::up(H16<DT>::down(s)), H16<DT>::up(b.h[e])));
        }
        out[i] = o.u;
    }
}

// ---- rotary embedding: x (B, H, L, hd) with element strides (sb, sh, sl, 1); cos / sin (Bc, L, hd) contiguous ------
template <int DT>
__global__ void __launch_bounds__(256) rope_kernel(const void *__restrict__ x, void *__restrict__ out, const void *__restrict__ cs,
                                                   const void *__restrict__ sn, int B, int H, int L, int hd, long sb, long sh, long sl,
                                                   long osb, long osh, long osl, int cos_batched) {
    using T = typename H16<DT>::T;
    const int half = hd / 2, hv = half / 8;          // vectors of 8 per half
    const long total = (long)B * H * L * hv;
    for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long)gridDim.x * blockDim.x) {
        const int v = (int)(id % hv);
        long t = id / hv;
        const int h = (int)(t % H);
        t /= H;
        const int l = (int)(t % L);
        const int b = (int)(t / L);
        const T *xp = (const T *)x + b * sb + h * sh + l * sl;
        T *op = (T *)out + b * osb + h * osh + l * osl;
        const T *cp = (const T *)cs + ((size_t)(cos_batched ? b : 0) * L + l) * hd;
        const T *sp = (const T *)sn + ((size_t)(cos_batched ? b : 0) * L + l) * hd;
        Vec8<DT> x1, x2, c1, c2, s1, s2, o1, o2;
        x1.u = *reinterpret_cast<const uint4 *>(xp + 8 * v);
        x2.u = *reinterpret_cast<const uint4 *>(xp + half + 8 * v);
        c1.u = *reinterpret_cast<const uint4 *>(cp + 8 * v);
        c2.u = *reinterpret_cast<const uint4 *>(cp + half + 8 * v);
        s1.u = *reinterpret_cast<const uint4 *>(sp + 8 * v);
        s2.u = *reinterpret_cast<const uint4 *>(sp + half + 8 * v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = H16<DT>::up(x1.h[e]), bq = H16<DT>::up(x2.h[e]);
            // first half: x1*cos + (-x2)*sin ; second half: x2*cos + x1*sin ; every product and the sum rounded to 16 bit
            const T p1 = H16<DT>::down(__fmul_rn(a, H16<DT>::up(c1.h[e])));
            const T q1 = H16<DT>::down(__fmul_rn(-bq, H16<DT>::up(s1.h[e])));
            o1.h[e] = H16<DT>::down(__fadd_rn(H16<DT>::up(p1), H16<DT>::up(q1)));
            const T p2 = H16<DT>::down(__fmul_rn(bq, H16<DT>::up(c2.h[e])));
            const T q2 = H16<DT>::down(__fmul_rn(a, H16<DT>::up(s2.h[e])));
            o2.h[e] = H16<DT>::down(__fadd_rn(H16<DT>::up(p2), H16<DT>::up(q2)));
        }
        *reinterpret_cast<uint4 *>(op + 8 * v) = o1.u;
        *reinterpret_cast<uint4 *>(op + half + 8 * v) = o2.u;
    }
}

int grid_for(long n) {
    long g = (n + 255) / 256;
    const long cap = 148L * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" GQ_API int gq_fwd_rmsnorm(const void *x, const void *w, void *out, long rows, int dim, float eps, int dtype,
                                     gq_stream_t stream) {
    GQ_REQUIRE(x && w && out && rows > 0, "gq_fwd_rmsnorm: bad arguments");
    GQ_REQUIRE(dtype == GQ_BF16 || dtype == GQ_F16, "gq_fwd_rmsnorm: 16-bit activations only");
    GQ_REQUIRE(dim > 0 && dim % 8 == 0 && dim <= 8192, "gq_fwd_rmsnorm: dim=%d must be a multiple of 8, at most 8192", dim);
    GQ_REQUIRE(((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) % 16 == 0, "gq_fwd_rmsnorm: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GQ_BF16) rmsnorm_kernel<GQ_BF16><<<(unsigned)rows, 256, 0, st>>>(x, w, out, dim, eps, (float)(1.0 / (double)dim));
    else rmsnorm_kernel<GQ_F16><<<(unsigned)rows, 256, 0, st>>>(x, w, out, dim, eps, (float)(1.0 / (double)dim));
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" GQ_API int gq_fwd_silu_mul(const void *gate, const void *up, void *out, long n, int dtype, gq_stream_t stream) {
    GQ_REQUIRE(gate && up && out && n > 0 && n % 8 == 0, "gq_fwd_silu_mul: bad arguments (n must be a multiple of 8)");
    GQ_REQUIRE(dtype == GQ_BF16 || dtype == GQ_F16, "gq_fwd_silu_mul: 16-bit activations only");
    GQ_REQUIRE(((uintptr_t)gate | (uintptr_t)up | (uintptr_t)out) % 16 == 0, "gq_fwd_silu_mul: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n / 8);
    if (dtype == GQ_BF16) silu_mul_kernel<GQ_BF16><<<g, 256, 0, st>>>((const uint4 *)gate, (const uint4 *)up, (uint4 *)out, n / 8);
    else silu_mul_kernel<GQ_F16><<<g, 256, 0, st>>>((const uint4 *)gate, (const uint4 *)up, (uint4 *)out, n / 8);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

extern "C" GQ_API int gq_fwd_rope(const void *x, void *out, const void *cos, const void *sin, int B, int H, int L, int hd, long sb,
                                  long sh, long sl, long osb, long osh, long osl, int cos_batched, int dtype, gq_stream_t stream) {
    GQ_REQUIRE(x && out && cos && sin && B > 0 && H > 0 && L > 0, "gq_fwd_rope: bad arguments");
    GQ_REQUIRE(dtype == GQ_BF16 || dtype == GQ_F16, "gq_fwd_rope: 16-bit activations only");
    GQ_REQUIRE(hd > 0 && hd % 16 == 0, "gq_fwd_rope: head_dim=%d must be a multiple of 16", hd);
    GQ_REQUIRE(((uintptr_t)x | (uintptr_t)out | (uintptr_t)cos | (uintptr_t)sin) % 16 == 0 && (sb | sh | sl | osb | osh | osl) % 8 == 0,
               "gq_fwd_rope: pointers must be 16-byte aligned and strides multiples of 8 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for((long)B * H * L * (hd / 16));
    if (dtype == GQ_BF16) rope_kernel<GQ_BF16><<<g, 256, 0, st>>>(x, out, cos, sin, B, H, L, hd, sb, sh, sl, osb, osh, osl, cos_batched);
    else rope_kernel<GQ_F16><<<g, 256, 0, st>>>(x, out, cos, sin, B, H, L, hd, sb, sh, sl, osb, osh, osl, cos_batched);
    gq_count_launches(1);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
