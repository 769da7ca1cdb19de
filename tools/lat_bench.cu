// Dependent-chain latencies that bound the in-register Householder panel (one warp, clock64 around 256 dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double x0) {
    double x = x0 + threadIdx.x * 1e-3, y = 1.000001;
    long long t0, t1;
    const int N = 256;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) clk[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = sqrt(x + 1.0);
    t1 = clock64(); if (threadIdx.x == 0) clk[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = 1.0 / (x + 1.0);
    t1 = clock64(); if (threadIdx.x == 0) clk[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 4) & 31);
    t1 = clock64(); if (threadIdx.x == 0) clk[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (threadIdx.x == 0) clk[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = rsqrt(x + 1.0);
    t1 = clock64(); if (threadIdx.x == 0) clk[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + y;
    t1 = clock64(); if (threadIdx.x == 0) clk[6] = t1 - t0;
    // throughput of independent double shuffles (16 per iteration)
    double v[16];
    for (int j = 0; j < 16; ++j) v[j] = x + j;
    t0 = clock64();
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __shfl_sync(0xffffffffu, v[j], (threadIdx.x + 4) & 31);
    }
    t1 = clock64(); if (threadIdx.x == 0) clk[7] = t1 - t0;
    for (int j = 0; j < 16; ++j) x += v[j];
    out[threadIdx.x] = x;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 256); cudaMalloc(&clk, 64);
    k<<<1, 32>>>(out, clk, 1.5); k<<<1, 32>>>(out, clk, 1.5);
    long long h[8]; cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    const char* n[8] = {"dfma", "sqrt(x+1)", "1/(x+1)", "shfl.f64", "shfl_xor.f64 + dadd", "rsqrt(x+1)", "dadd", "16 indep shfl.f64 (per shfl)"};
    for (int i = 0; i < 8; ++i) printf("%-30s %.1f clk\n", n[i], h[i] / 256.0);
    return 0;
}
