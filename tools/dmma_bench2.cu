// Microbenchmark: DMMA m8n8k4 throughput per SM sub-partition as a function of resident warps and independent
// accumulator chains per warp, alone and with scalar DFMAs interleaved (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CHAINS, int NFMA>  // NFMA scalar DFMAs (independent chains) issued per DMMA
__global__ void k(double* out, int iters, long long* clk) {
    double c[CHAINS][2];
    double f[8];
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int i = 0; i < 8; ++i) f[i] = i + threadIdx.x * 1e-3;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < CHAINS; ++u) {
            dmma(c[u][0], c[u][1], a, b);
#pragma unroll
            for (int q = 0; q < NFMA; ++q) f[(u * NFMA + q) & 7] = fma(f[(u * NFMA + q) & 7], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < 8; ++i) s += f[i];
    if (s == -1.0) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int CHAINS, int NFMA> void run(int sms, int warps_per_smsp) {
    double* out; cudaMalloc(&out, 8);
    long long* clk; cudaMalloc(&clk, 8);
    int iters = 4096, grid = sms, block = 32 * 4 * warps_per_smsp;
    k<CHAINS, NFMA><<<grid, block>>>(out, iters, clk);
    k<CHAINS, NFMA><<<grid, block>>>(out, iters, clk);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double dmma_per_smsp = (double)warps_per_smsp * iters * CHAINS;
    printf("warps/SMSP %d chains %d fma/dmma %d : %6.1f clk per DMMA per SMSP (ideal 16 + %d)  pipe busy %.0f%%\n",
           warps_per_smsp, CHAINS, NFMA, h / dmma_per_smsp, 2 * NFMA, 100.0 * (16 + 2 * NFMA) / (h / dmma_per_smsp));
    cudaFree(out); cudaFree(clk);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    int sms = p.multiProcessorCount;
    for (int w : {1, 2, 3, 4, 8}) { run<1, 0>(sms, w); run<2, 0>(sms, w); run<4, 0>(sms, w); run<8, 0>(sms, w); }
    for (int w : {1, 2, 3, 4, 8}) { run<4, 1>(sms, w); run<4, 4>(sms, w); run<4, 8>(sms, w); }
    return 0;
}
