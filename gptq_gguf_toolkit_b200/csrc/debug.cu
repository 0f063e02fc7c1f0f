// debug.cu -- test hooks (exported, NOT part of the reference-facing API in include/gq.h).
#include "common.cuh"
#include "f32x2.cuh"

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
}  // namespace

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
