// How fast can an SM sub-partition issue the rank-k update's FFMA pattern?  (B200, sm_100a)
//   acc[8][4] += e[8] (x) u[4]   per k, operands from registers (A, B) or from shared memory as in the kernel (C, D).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma_tile ffma_tile.cu && ./ffma_tile
#include <cstdio>
#include <cuda_runtime.h>
#define KSTEPS 4096
template <int MODE> __global__ void __launch_bounds__(256, 1) tile(float *out, long long *cyc, const float *gsrc) {
    __shared__ __align__(16) float us[16 * 256];
    __shared__ __align__(16) float es[16 * 32];
    for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) us[i] = gsrc[i];
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) es[i] = gsrc[4096 + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rgrp = lane >> 3, cgrp = lane & 7;
    float acc[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float e[8], u[4];
    for (int i = 0; i < 8; ++i) e[i] = es[i + 8 * rgrp];
    for (int j = 0; j < 4; ++j) u[j] = us[j + 4 * cgrp];
    long long t0 = clock64();
    if (MODE == 0) {          // A: registers, i outer / j inner
#pragma unroll 16
        for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(e[i], u[j], acc[i][j]);
    } else if (MODE == 1) {   // B: registers, j outer / i inner
#pragma unroll 16
        for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][j] = __fmaf_rn(e[i], u[j], acc[i][j]);
    } else {                  // C / D: operands from shared memory exactly like rank_update (3 LDS.128 per k)
        for (int k0 = 0; k0 < KSTEPS; k0 += 16) {
            const float *up = us + 32 * warp + 4 * cgrp, *ep = es + 8 * rgrp;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float4 uu = *reinterpret_cast<const float4 *>(up + k * 256);
                const float4 ea = *reinterpret_cast<const float4 *>(ep + k * 32);
                const float4 eb = *reinterpret_cast<const float4 *>(ep + k * 32 + 4);
                const float ev[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
                const float uv[4] = {uu.x, uu.y, uu.z, uu.w};
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(ev[i], uv[j], acc[i][j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i][j] = __fmaf_rn(ev[i], uv[j], acc[i][j]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, float *out, long long *cyc, float *src) {
    for (int threads : {32, 128, 256}) {
        tile<MODE><<<1, threads>>>(out, cyc, src);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s threads=%3d: %.1f cyc per k-step (32 FFMA/lane)  -> %.2f FFMA/clk/SMSP\n", name, threads, (double)h / KSTEPS,
               32.0 * ((threads + 127) / 128) / ((double)h / KSTEPS));
    }
}
int main() {
    float *out, *src; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096); cudaMalloc(&src, 8192 * 4);
    cudaMemset(src, 0, 8192 * 4);
    run<0>("A regs, i outer j inner", out, cyc, src);
    run<1>("B regs, j outer i inner", out, cyc, src);
    run<2>("C smem operands (3 LDS.128/k), i outer", out, cyc, src);
    run<3>("D smem operands (3 LDS.128/k), j outer", out, cyc, src);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
