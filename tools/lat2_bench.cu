// Dependent-chain latencies of the building blocks of the Householder panel (one warp, 1 block): clock64 around N ops.
#include <cstdio>
#include "../pymc_experimental_b200/csrc/bkf_sqrt_warp.cuh"
using namespace bkf::sqw;
__global__ void k(double* out, long long* clk, double x0) {
    __shared__ __align__(16) double sm[64];
    const int lane = threadIdx.x;
    double x = x0 + lane * 1e-3, acc = 0;
    long long t0, t1;
    const int N = 128;
    // reflector() chain
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) { Refl R = reflector(x, x * 0.5); x = fabs(R.ts) + 1.0; }
    acc += x; t1 = clock64(); if (lane == 0) clk[0] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    acc += x; t1 = clock64(); if (lane == 0) clk[1] = t1 - t0;
    x = fmin(fabs(x), 3.0) + x0 + lane * 1e-3;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    acc += x; t1 = clock64(); if (lane == 0) clk[2] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = red4_mma(x);
    acc += x; t1 = clock64(); if (lane == 0) clk[3] = t1 - t0;
    x = fmin(fabs(x), 3.0) + x0 + lane * 1e-3;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) { const bool b = __any_sync(0xffffffffu, x > 1e300); x = b ? x + 1.0 : x; }
    acc += x; t1 = clock64(); if (lane == 0) clk[4] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        if ((lane >> 2) == (i & 7)) reinterpret_cast<double2*>(sm)[lane & 3] = make_double2(x, x);
        __syncwarp();
        x = reinterpret_cast<const double2*>(sm)[lane & 3].x;
    }
    acc += x; t1 = clock64(); if (lane == 0) clk[5] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = red_t(x) * 0.25;
    acc += x; t1 = clock64(); if (lane == 0) clk[6] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i * 4) & 31);
    acc += x; t1 = clock64(); if (lane == 0) clk[7] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = (x > 0.5) ? x * 0.999 : 1.0;   // DSETP + select chain
    acc += x; t1 = clock64(); if (lane == 0) clk[8] = t1 - t0;
    t0 = clock64(); asm volatile("" : "+d"(x) :: "memory");
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = copysign(x, -x) * -1.0;
    acc += x; t1 = clock64(); if (lane == 0) clk[9] = t1 - t0;
    out[lane] = x + acc;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 256); cudaMalloc(&clk, 128);
    k<<<1, 32>>>(out, clk, 1.5); k<<<1, 32>>>(out, clk, 1.5);
    long long h[10]; cudaMemcpy(h, clk, 80, cudaMemcpyDeviceToHost);
    const char* n[10] = {"reflector() + fabs + add", "rsqrt.approx.f64", "rcp.approx.f64", "red4_mma (DMMA x ones)", "vote.any + select", "STS(pred) syncwarp LDS.128", "red_t (2 shfl_xor+add) + mul", "shfl.f64 idx", "dsetp+select+dmul", "copysign+dmul"};
    for (int i = 0; i < 10; ++i) printf("%-32s %.1f clk\n", n[i], h[i] / 128.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
