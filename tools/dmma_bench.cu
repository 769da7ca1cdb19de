// Microbenchmark: FP64 tensor-core (mma.sync.m8n8k4.f64 "DMMA") issue rate on sm_100a, alone and mixed with DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>  // 0: DMMA only, 1: DFMA only, 2: both interleaved
__global__ void k(double* out, int iters) {
    double c[8][2];
    double f[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i + threadIdx.x * 1e-3; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE != 1) dmma(c[u][0], c[u][1], a, b);
            if (MODE != 0) { f[u] = fma(f[u], a, b); }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    if (s == -1.0) out[0] = s;
}
template <int MODE> void run(const char* name, int sms) {
    double* out; cudaMalloc(&out, 8);
    int iters = 20000, grid = sms * 4, block = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a); k<MODE><<<grid, block>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double warps = (double)grid * block / 32;
        double mma_flops = (MODE != 1) ? warps * iters * 8.0 * 2 * 8 * 8 * 4 : 0;
        double fma_flops = (MODE != 0) ? (double)grid * block * iters * 8.0 * 2 : 0;
        if (rep == 2) printf("%-28s %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s\n", name, ms, mma_flops / ms / 1e9, fma_flops / ms / 1e9);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("DMMA m8n8k4 only", p.multiProcessorCount);
    run<1>("DFMA only", p.multiProcessorCount);
    run<2>("DMMA + DFMA interleaved", p.multiProcessorCount);
    return 0;
}
