// Which part of rank_update's cp.async pipeline costs 128 cycles per k-step?  (B200, sm_100a)
//   one CTA = 256 threads, 32 rows x 256 cols tile, pieces of KP=16 k's: U piece 16 KB (cp.async 16 B), E^T piece 2 KB.
//   flags: 1 = load U, 2 = load E with 4-byte cp.async (transposing), 4 = load E with 16-byte cp.async (row-major),
//          8 = FFMA compute, 16 = U rows are 56 KB apart (d_col 14336) instead of 16 KB
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipe pipe.cu && ./pipe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NT = 256, KP = 16, R = 32;
__device__ __forceinline__ void cp16(void *s, const void *g) { uint32_t a = (uint32_t)__cvta_generic_to_shared(s); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g)); }
__device__ __forceinline__ void cp4(void *s, const void *g) { uint32_t a = (uint32_t)__cvta_generic_to_shared(s); asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(g)); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int S, int FLAGS> __global__ void __launch_bounds__(NT, 1) pipe(const float *U, const float *W, int ld, int K, float *out, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    float *Us = sm, *Es = sm + S * KP * 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rgrp = lane >> 3, cgrp = lane & 7;
    const int r0 = blockIdx.x * R, P = K / KP;
    auto issue = [&](int pc) {
        if (pc < P) {
            const int k0 = KP * pc, st = pc % S;
            float *us = Us + st * KP * 256, *es = Es + st * KP * R;
            if (FLAGS & 1)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int id = tid + NT * m, row = id >> 6, c16 = id & 63;
                    cp16(us + row * 256 + 4 * c16, U + (size_t)(k0 + row) * ld + 4 * c16);
                }
            if (FLAGS & 2)
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int id = tid + NT * m, row = id >> 4, k = id & 15;
                    cp4(es + k * R + row, W + (size_t)(r0 + row) * ld + k0 + k);
                }
            if ((FLAGS & 4) && tid < 128) {
                const int row = tid >> 2, part = tid & 3;
                cp16(es + row * KP + 4 * part, W + (size_t)(r0 + row) * ld + k0 + 4 * part);
            }
        }
        commit();
    };
    float acc[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    long long t0 = clock64();
    for (int s = 0; s < S - 1; ++s) issue(s);
    for (int pc = 0; pc < P; ++pc) {
        waitg<S - 2>();
        __syncthreads();
        issue(pc + S - 1);
        if (FLAGS & 8) {
            const float *up = Us + (pc % S) * KP * 256 + 32 * warp + 4 * cgrp, *ep = Es + (pc % S) * KP * R + 8 * rgrp;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const float4 uu = *reinterpret_cast<const float4 *>(up + k * 256);
                const float4 ea = *reinterpret_cast<const float4 *>(ep + k * R);
                const float4 eb = *reinterpret_cast<const float4 *>(ep + k * R + 4);
                const float ev[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
                const float uv[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(ev[i], uv[j], acc[i][j]);
            }
        }
    }
    waitg<0>();
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * NT + tid] = s + sm[tid];
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int S, int FLAGS> void run(const char *name, const float *U, const float *W, float *out, long long *cyc) {
    const int K = 4096, ld = (FLAGS & 16) ? 14336 : 4096;
    const size_t smem = (size_t)S * (KP * 256 + KP * R) * 4;
    cudaFuncSetAttribute(pipe<S, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {1, 8, 128}) {
        pipe<S, FLAGS><<<grid, NT, smem>>>(U, W, ld, K, out, cyc);
        cudaDeviceSynchronize();
        long long h[128];
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-52s S=%d grid=%3d: %.1f cyc per k-step\n", name, S, grid, (double)mx / K);
    }
}
int main() {
    float *U, *W, *out; long long *cyc;
    cudaMalloc(&U, (size_t)4096 * 14336 * 4); cudaMalloc(&W, (size_t)4096 * 14336 * 4); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
    cudaMemset(U, 0, (size_t)4096 * 14336 * 4); cudaMemset(W, 0, (size_t)4096 * 14336 * 4);
    run<4, 8>("compute only (stale smem)", U, W, out, cyc);
    run<4, 1>("load U only", U, W, out, cyc);
    run<4, 2>("load E (4-byte transposing) only", U, W, out, cyc);
    run<4, 4>("load E (16-byte) only", U, W, out, cyc);
    run<4, 1 | 2>("load U + E4", U, W, out, cyc);
    run<4, 1 | 8>("load U + compute", U, W, out, cyc);
    run<4, 1 | 2 | 8>("load U + E4 + compute (= kernel)", U, W, out, cyc);
    run<4, 1 | 4 | 8>("load U + E16 + compute", U, W, out, cyc);
    run<3, 1 | 2 | 8>("load U + E4 + compute", U, W, out, cyc);
    run<6, 1 | 2 | 8>("load U + E4 + compute", U, W, out, cyc);
    run<4, 1 | 2 | 8 | 16>("load U + E4 + compute, ld=14336", U, W, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
