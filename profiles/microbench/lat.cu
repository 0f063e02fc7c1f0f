// Latency / throughput probes for the instructions on the column loop's dependent chain (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gptq_gguf_toolkit_b200/csrc/f32x2.cuh"

#define N 2048
// clock reads are tied to the value chain through asm operands so that the compiler cannot move the loop out
__device__ __forceinline__ long long tick(float &x) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t), "+f"(x) :: "memory");
    return t;
}
template <class F> __device__ long long chain(F f, float &x) {
    long long t0 = tick(x);
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = f(x);
    long long t1 = tick(x);
    return t1 - t0;
}
__global__ void probe(float *out, long long *cyc, float a, float b, f2_t nz2, int q) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = a + i * 1e-3f;
    __syncthreads();
    float x = a + threadIdx.x * 1e-3f, sink = 0.0f;
    int k = 0;
    cyc[k++] = chain([&](float v) { return __fadd_rn(v, b); }, x);                               // 0 FADD
    cyc[k++] = chain([&](float v) { return __fmaf_rn(v, b, a); }, x);                            // 1 FFMA
    cyc[k++] = chain([&](float v) { return __shfl_sync(0xffffffffu, v, q, 4) + b; }, x);             // 2 SHFL + FADD
    cyc[k++] = chain([&](float v) { return rintf(v) + b; }, x);                                      // 3 FRND + FADD
    cyc[k++] = chain([&](float v) { return fminf(fmaxf(v, 0.0f), 15.0f) + b; }, x);              // 4 FMNMX x2 + FADD
    sink += x; x = a + 1e-30f * sink;
    cyc[k++] = chain([&](float v) { return __fdiv_rn(v, b) + a; }, x);                           // 5 fdiv + FADD
    sink += x; x = a + 1e-30f * sink;
    DivBy d = DivBy::make(b);
    cyc[k++] = chain([&](float v) { return d.div(v) + a; }, x);                                  // 6 DivBy.div + FADD
    sink += x; x = a + 1e-30f * sink;
    cyc[k++] = chain([&](float v) { float lo, hi; f2_unpack(f2_fma(f2_pack(v, v), f2_pack(b, b), f2_pack(a, a)), lo, hi); return lo; }, x);  // 7 FFMA2
    sink += x; x = a + 1e-30f * sink;
    cyc[k++] = chain([&](float v) { float lo, hi; f2_unpack(f2_sub(f2_pack(v, v), f2_mul_nofuse(f2_pack(b, b), f2_pack(a, a), nz2)), lo, hi); return hi; }, x);  // 8 FADD2 (mul off chain)
    sink += x; x = a + 1e-30f * sink;
    cyc[k++] = chain([&](float v) { float lo, hi; f2_unpack(f2_sub(f2_pack(a, a), f2_mul_nofuse(f2_pack(v, v), f2_pack(b, b), nz2)), lo, hi); return hi; }, x);  // 9 FFMA2(mul)+FADD2
    sink += x; x = a + 1e-30f * sink;
    cyc[k++] = chain([&](float v) { return sm[((int)v) & 1023] ; }, x);                          // 10 F2I + LDS
    {   // 11: throughput of independent FFMA2 (8 chains) vs 12: independent FFMA (16 chains)
        f2_t acc[8];
        for (int j = 0; j < 8; ++j) acc[j] = f2_pack(a + j, b + j);
        const f2_t bb = f2_pack(b, b);
        long long t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = f2_fma(acc[j], bb, bb);
        long long t1 = clock64();
        cyc[k++] = t1 - t0;
        float s = 0;
        for (int j = 0; j < 8; ++j) { float lo, hi; f2_unpack(acc[j], lo, hi); s += lo + hi; }
        x += s;
    }
    {
        float acc[16];
        for (int j = 0; j < 16; ++j) acc[j] = a + j;
        long long t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __fmaf_rn(acc[j], b, b);
        long long t1 = clock64();
        cyc[k++] = t1 - t0;
        float s = 0;
        for (int j = 0; j < 16; ++j) s += acc[j];
        x += s;
    }
    out[threadIdx.x] = x + sink;
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64 * 8);
    const char *names[] = {"FADD", "FFMA", "SHFL+FADD", "FRND+FADD", "FMNMX2+FADD", "fdiv_rn+FADD", "DivBy+FADD", "FFMA2(pack/unpack)",
                           "FADD2 chain", "FFMA2mul+FADD2 chain", "F2I+LDS", "8 indep FFMA2 (per iter)", "16 indep FFMA (per iter)"};
    for (int threads : {32, 128, 256}) {
        probe<<<1, threads>>>(out, cyc, 1.25f, 0.999f, F2_NEG_ZERO2, 1);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%d\n", threads);
        for (int i = 0; i < 13; ++i) printf("  %-28s %.2f cyc/iter\n", names[i], (double)h[i] / N);
    }
    {   // clock64 ticks per microsecond (is clock64 the boosted SM clock?)
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); probe<<<1, 32>>>(out, cyc, 1.25f, 0.999f, F2_NEG_ZERO2, 1); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long tot = 0; for (int i = 0; i < 13; ++i) tot += h[i];
        printf("sum of probe ticks %lld over %.3f ms kernel => >= %.0f MHz tick rate\n", tot, ms, tot / (ms * 1e3));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
