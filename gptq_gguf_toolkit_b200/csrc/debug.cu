// debug.cu -- test hooks (exported, NOT part of the reference-facing API in include/gq.h).
#include "common.cuh"
#include "f32x2.cuh"
#include "kquant.cuh"

namespace {
__global__ void divby_kernel(const float *a, const float *b, float *out, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = DivBy::make(b[i]).div(a[i]);
}
// out[i] = a[i] - fl(e[i] * u[i]) on both halves of a packed pair (f2_mul_nofuse + f2_sub): must be two roundings
__global__ void mulsub2_kernel(const float *a, const float *e, const float *u, float *out, long n, f2_t nz2) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < n; i += (long)gridDim.x * blockDim.x) {
        const f2_t r = f2_sub(f2_pack(a[2 * i], a[2 * i + 1]),
                              f2_mul_nofuse(f2_pack(e[2 * i], e[2 * i + 1]), f2_pack(u[2 * i], u[2 * i + 1]), nz2));
        f2_unpack(r, out[2 * i], out[2 * i + 1]);
    }
}
// counts the bit patterns u in [first, last] (both signs, NaNs skipped) for which kq_rint_clamp differs from clamp(rintf(.)) as a value
__global__ void rint_clamp_check_kernel(unsigned int first, unsigned int last, float lo, float hi, unsigned long long *bad) {
    unsigned long long mine = 0;
    for (unsigned long long u = (unsigned long long)first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= last;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        if ((u & 0x7FFFFFFFull) > 0x7F800000ull) continue;
#pragma unroll
        for (unsigned int sign = 0; sign < 2; ++sign) {
            const float v = __uint_as_float((unsigned int)u | (sign << 31));
            const float a = kq_rint_clamp(v, lo, hi), b = clampf(rintf(v), lo, hi);
            if (!(a == b)) ++mine;
        }
    }
    if (mine) atomicAdd(bad, mine);
}
__global__ void sq_u8_check_kernel(unsigned long long *bad) {
    const int l = threadIdx.x;
    int m = 0;
    if (l <= 3) m += kq_sq_u8<3>((float)l) != (float)((l * l) & 255);
    if (l <= 15) m += kq_sq_u8<15>((float)l) != (float)((l * l) & 255);
    if (l <= 31) m += kq_sq_u8<31>((float)l) != (float)((l * l) & 255);
    m += kq_sq_u8<255>((float)l) != (float)((l * l) & 255);
    if (m) atomicAdd(bad, (unsigned long long)m);
}
}  // namespace

// *bad (device, zeroed by the caller) += number of float bit patterns in [first, last] (both signs) where the conversion-free
// kq_rint_clamp (kquant.cuh) and clamp(rintf(v), lo, hi) give different values, plus the mismatches of kq_sq_u8 over all codes
extern "C" GQ_API int gq_debug_rint_clamp(unsigned int first, unsigned int last, float lo, float hi, unsigned long long *bad,
                                          gq_stream_t stream) {
    rint_clamp_check_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(first, last, lo, hi, bad);
    sq_u8_check_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(bad);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}

// out[i] = DivBy::make(b[i]).div(a[i])  -- must equal the IEEE quotient a[i] / b[i] bit for bit
extern "C" GQ_API int gq_debug_divby(const float *a, const float *b, float *out, long n, gq_stream_t stream) {
    divby_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
extern "C" GQ_API int gq_debug_mulsub2(const float *a, const float *e, const float *u, float *out, long n, gq_stream_t stream) {
    mulsub2_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(a, e, u, out, n, F2_NEG_ZERO2);
    GQ_CHECK_CUDA(cudaGetLastError());
    return GQ_OK;
}
