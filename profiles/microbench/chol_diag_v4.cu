// Stand-alone check + timing of gptq_gguf_toolkit_b200/csrc/chol_diag_v4.cuh, the two-level (warp-shuffle 32 x 32 factorisation +
// substitution panel solve) variant of the (128 x 128) diagonal-block kernel of gq_prepare.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o chol_diag_v4 chol_diag_v4.cu && ./chol_diag_v4
// The host code checks L, inv(L) and inv(L)^T against a double-precision factorisation and times 200 launches.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef SIMT_EMU          // tests/test_simt_emu_cpu.py runs the kernel below on the host (tests/helpers/simt_emu)
#include <cuda_runtime.h>
#endif

#define CD4_PROFILE 1
#include "../../gptq_gguf_toolkit_b200/csrc/chol_diag_v4.cuh"
using namespace cd4;

#ifndef SIMT_EMU
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
    const int n = 384, k0 = 128;          // the block in the middle of a 384 x 384 matrix: exercises ld and k0
    std::vector<double> M((size_t)NB * 3 * NB), H((size_t)NB * NB);
    srand(1);
    for (auto &v : M) v = (rand() / (double)RAND_MAX - 0.5);
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) {
            double acc = (i == j) ? 0.05 : 0.0;
            for (int k = 0; k < 3 * NB; ++k) acc += M[(size_t)i * 3 * NB + k] * M[(size_t)j * 3 * NB + k] / (3.0 * NB);
            H[(size_t)i * NB + j] = acc;
        }
    // double-precision reference: L and X = inv(L)
    std::vector<double> L(H), X((size_t)NB * NB, 0.0);
    for (int j = 0; j < NB; ++j) {
        for (int k = 0; k < j; ++k)
            for (int i = j; i < NB; ++i) L[(size_t)i * NB + j] -= L[(size_t)i * NB + k] * L[(size_t)j * NB + k];
        const double dj = sqrt(L[(size_t)j * NB + j]);
        for (int i = j; i < NB; ++i) L[(size_t)i * NB + j] /= dj;
    }
    for (int c = 0; c < NB; ++c)
        for (int i = c; i < NB; ++i) {
            double acc = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) acc -= L[(size_t)i * NB + k] * X[(size_t)k * NB + c];
            X[(size_t)i * NB + c] = acc / L[(size_t)i * NB + i];
        }
    std::vector<float> hA((size_t)n * n, 0.0f);
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) hA[(size_t)(k0 + i) * n + k0 + j] = (float)H[(size_t)i * NB + j];
    float *dA, *dA0, *dB, *dBT;
    int *dflag;
    CK(cudaMalloc(&dA, sizeof(float) * n * n)); CK(cudaMalloc(&dA0, sizeof(float) * n * n));
    CK(cudaMalloc(&dB, sizeof(float) * n * n)); CK(cudaMalloc(&dBT, sizeof(float) * n * n)); CK(cudaMalloc(&dflag, sizeof(int)));
    CK(cudaMemcpy(dA0, hA.data(), sizeof(float) * n * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(dB, 0, sizeof(float) * n * n)); CK(cudaMemset(dBT, 0, sizeof(float) * n * n)); CK(cudaMemset(dflag, 0, sizeof(int)));
    CK(cudaFuncSetAttribute(chol_diag_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem4)));
    CK(cudaMemcpy(dA, dA0, sizeof(float) * n * n, cudaMemcpyDeviceToDevice));
    chol_diag_v4_kernel<<<1, T4, sizeof(Smem4)>>>(dA, dB, dBT, n, k0, dflag);
    CK(cudaDeviceSynchronize());
    std::vector<float> gL((size_t)n * n), gX((size_t)n * n), gXT((size_t)n * n);
    int flag = 0;
    CK(cudaMemcpy(gL.data(), dA, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gX.data(), dB, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gXT.data(), dBT, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost));
    double eL = 0, eX = 0, eXT = 0, mL = 0, mX = 0;
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) {
            const double l = (j <= i) ? L[(size_t)i * NB + j] : 0.0, x = (j <= i) ? X[(size_t)i * NB + j] : 0.0;
            if (j <= i) eL = fmax(eL, fabs(gL[(size_t)(k0 + i) * n + k0 + j] - l));
            eX = fmax(eX, fabs(gX[(size_t)(k0 + i) * n + k0 + j] - x));
            eXT = fmax(eXT, fabs(gXT[(size_t)(k0 + j) * n + k0 + i] - x));
            mL = fmax(mL, fabs(l)); mX = fmax(mX, fabs(x));
        }
    printf("not_pd %d   max|L - L64| / max|L| = %.2e   max|X - X64| / max|X| = %.2e   transpose %.2e\n", flag, eL / mL, eX / mX, eXT / mX);
    const bool ok = flag == 0 && eL / mL < 1e-4 && eX / mX < 1e-3 && eXT / mX < 1e-3;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 200;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) chol_diag_v4_kernel<<<1, T4, sizeof(Smem4)>>>(dA, dB, dBT, n, k0, dflag);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%s   chol_diag_v4: %.1f us per launch (v3 70.4 us, v2 88.7 us, v1 129 us on B200)\n", ok ? "OK" : "MISMATCH",
           1e3 * ms / reps);
    {
        unsigned long long z[8] = {0}, c[8];
        CK(cudaMemcpyToSymbol(cd4_clk, z, sizeof(z)));
        chol_diag_v4_kernel<<<1, T4, sizeof(Smem4)>>>(dA, dB, dBT, n, k0, dflag);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpyFromSymbol(c, cd4_clk, sizeof(c)));
        const char *nm[7] = {"load", "warp-factor x4", "panel-solve x3", "trailing x3", "diag-inverse", "offdiag-inverse", "store"};
        unsigned long long tot = 0;
        for (int i = 0; i < 7; ++i) tot += c[i];
        printf("phase cycles (thread 0, one launch, total %llu):", tot);
        for (int i = 0; i < 7; ++i) printf("  %s %llu", nm[i], c[i]);
        printf("\n");
    }
    return ok ? 0 : 2;
}
#endif  // SIMT_EMU
