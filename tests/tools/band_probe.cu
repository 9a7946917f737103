// Timing probe for the single-CTA banded LDL^T (coop_ldlt.cuh) and the raw FP64 latencies behind it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=true -I stereovision-slam_b200/csrc tests/tools/band_probe.cu -o tests/tools/band_probe.bin
#include <cstdio>
#include <cstdlib>
#include <cfloat>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "coop_ldlt.cuh"

__global__ void k_lat(double *out, long long *cyc, double seed)
{
    double x = seed, y = seed * 0.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) x = fma(x, 1.0000001, y);
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = 1.0 / (x + 1.5);
    long long t2 = clock64();
    float f = (float)seed;
#pragma unroll 1
    for (int i = 0; i < 1024; i++) f = fmaf(f, 1.0001f, 0.5f);
    long long t3 = clock64();
    out[threadIdx.x] = x + f;
    if (threadIdx.x == 0) { cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = t3 - t2; }
}

__global__ void __launch_bounds__(512, 1) k_band(const double *B0, const double *z0, int n, int hb, double *x, double *sign, long long *cyc, int reps)
{
    extern __shared__ double sm[];
    double *Bb = sm, *z = Bb + (size_t)n * (hb + 1);
    long long tot = 0;
    for (int r = 0; r < reps; r++) {
        for (int t = threadIdx.x; t < n * (hb + 1); t += 512) Bb[t] = B0[t];
        for (int t = threadIdx.x; t < n; t += 512) z[t] = z0[t];
        __syncthreads();
        long long t0 = clock64();
        band_ldlt_solve_cta<512, true>(Bb, z, n, hb, x, sign, cyc + 1);
        tot += clock64() - t0;
    }
    if (threadIdx.x == 0) cyc[0] = tot / reps;
}

int main()
{
    const int n = 300, hb = 41, w = hb + 1;
    std::vector<double> A((size_t)n * n, 0.0), B((size_t)n * w, 0.0), z(n), xref(n);
    srand(1);
    for (int i = 0; i < n; i++) for (int j = std::max(0, i - hb); j <= i; j++) { double v = (rand() / (double)RAND_MAX - 0.5); if (i == j) v = 30 + fabs(v); A[(size_t)i * n + j] = A[(size_t)j * n + i] = v; B[(size_t)i * w + (i - j)] = v; }
    for (int i = 0; i < n; i++) xref[i] = rand() / (double)RAND_MAX;
    for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += A[(size_t)i * n + j] * xref[j]; z[i] = s; }
    double *dB, *dz, *dx, *dsign, *dout; long long *dc;
    cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dz, n * 8); cudaMalloc(&dx, n * 8); cudaMalloc(&dsign, 8); cudaMalloc(&dc, 64); cudaMalloc(&dout, 8 * 64);
    cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dz, z.data(), n * 8, cudaMemcpyHostToDevice);
    const size_t smem = CL_BAND_DOUBLES(n, hb) * 8 + 64;
    cudaFuncSetAttribute(k_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_band<<<1, 512, smem>>>(dB, dz, n, hb, dx, dsign, dc, 5);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc[8]; std::vector<double> x(n); double sign;
    cudaMemcpy(cyc, dc, 40, cudaMemcpyDeviceToHost);
    printf("thread 1 per step: barrier wait %.1f, pivot load %.1f, first update %.1f, pivot+rest %.1f\n", cyc[1] / (double)n, cyc[2] / (double)n, cyc[3] / (double)n, cyc[4] / (double)n); cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(&sign, dsign, 8, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < n; i++) err = fmax(err, fabs(x[i] - xref[i]));
    printf("band n=%d hb=%d: %s cycles/solve %lld (%.1f per step) max err %.3e sign %.0f\n", n, hb, cudaGetErrorString(e), cyc[0], cyc[0] / (double)n, err, sign);
    k_lat<<<1, 32>>>(dout, dc, 1.0);
    cudaDeviceSynchronize();
    cudaMemcpy(cyc, dc, 24, cudaMemcpyDeviceToHost);
    printf("latency: dependent DFMA %.1f cycles, dependent (DADD + 1/x) %.1f cycles, dependent FFMA %.1f cycles\n", cyc[0] / 1024.0, cyc[1] / 256.0, cyc[2] / 1024.0);
    return 0;
}
