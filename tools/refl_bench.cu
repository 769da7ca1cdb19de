// Microbenchmark of the Householder panel of the SquareRootFilter forward sweep (bkf_sqrt_warp.cuh): cycles per
// reflector of one warp per scheduler (the C4 regime: 512 one-warp CTAs on 148 SMs), variant against variant, and the
// difference of their results.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -o tools/refl_bench tools/refl_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../pymc_experimental_b200/csrc/bkf_sqrt_warp.cuh"

using namespace bkf::sqw;

template <int NQ, int VAR>
__global__ void __launch_bounds__(32) kern(const double* __restrict__ in, double* __restrict__ out, long long* clk, int iters) {
    __shared__ __align__(16) double tq[72 + 128 + 64];
    const int lane = threadIdx.x;
    double2 pa0[NQ];
    for (int q = 0; q < NQ; ++q) pa0[q] = reinterpret_cast<const double2*>(in)[(blockIdx.x * NQ + q) * 32 + lane];
    double2 pa[NQ], Rd = make_double2(0, 0);
    double tauv = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) pa[q] = pa0[q];
        if (it) pa[0].x += Rd.x * 1e-30;  // keep the iterations dependent
        if (VAR == 0) panel_reflectors<NQ>(pa, Rd, tauv, tq, tq + 72, lane);
        else panel_reflectors_la<NQ>(pa, Rd, tauv, tq, tq + 72, lane);
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) clk[blockIdx.x] = t1 - t0;
    double* o = out + (size_t)blockIdx.x * (NQ + 1) * 64 + 64;
    for (int q = 0; q < NQ; ++q) reinterpret_cast<double2*>(o)[q * 32 + lane] = pa[q];
    reinterpret_cast<double2*>(out + (size_t)blockIdx.x * (NQ + 1) * 64)[lane] = Rd;
    if ((lane & 3) == 0) o[NQ * 64 - 8 + (lane >> 2)] += 0.0 * tauv;
}

template <int NQ>
void run(int nblk, int iters, int mode) {
    const size_t nin = (size_t)nblk * NQ * 64, nout = (size_t)nblk * (NQ + 1) * 64;
    std::vector<double> h(nin);
    srand(7 + mode);
    for (auto& x : h) x = (rand() / (double)RAND_MAX - 0.5);
    if (mode == 1)  // nearly dependent columns and zero columns
        for (int b = 0; b < nblk; ++b)
            for (int q = 0; q < NQ; ++q)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2;
                    double* e = &h[((size_t)(b * NQ + q) * 32 + lane) * 2];
                    if (g == 3) { e[0] = 0; e[1] = 0; }
                    if (g == 5) {
                        const double* s = &h[((size_t)(b * NQ + q) * 32 + (lane & 3) + 8) * 2];
                        e[0] = s[0] * (1 + 1e-7 * e[0]);
                        e[1] = s[1] * (1 + 1e-7 * e[1]);
                    }
                }
    double *din, *dout[2];
    long long* dclk;
    cudaMalloc(&din, nin * 8);
    cudaMalloc(&dout[0], nout * 8);
    cudaMalloc(&dout[1], nout * 8);
    cudaMalloc(&dclk, nblk * 8);
    cudaMemcpy(din, h.data(), nin * 8, cudaMemcpyHostToDevice);
    std::vector<double> r[2];
    double cyc[2];
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            if (v == 0) kern<NQ, 0><<<nblk, 32>>>(din, dout[v], dclk, iters);
            else kern<NQ, 2><<<nblk, 32>>>(din, dout[v], dclk, iters);
            cudaDeviceSynchronize();
        }
        std::vector<long long> c(nblk);
        cudaMemcpy(c.data(), dclk, nblk * 8, cudaMemcpyDeviceToHost);
        double s = 0;
        for (auto x : c) s += x;
        cyc[v] = s / nblk / iters / 8.0;
        r[v].resize(nout);
        cudaMemcpy(r[v].data(), dout[v], nout * 8, cudaMemcpyDeviceToHost);
    }
    double worst = 0, scale = 0;
    for (size_t i = 0; i < nout; ++i) {
        const double d = fabs(r[0][i] - r[1][i]);
        if (!(d <= worst)) worst = d;
        if (fabs(r[0][i]) > scale) scale = fabs(r[0][i]);
    }
    printf("NQ=%d mode=%d blocks=%d: base %.1f cycles/reflector, look-ahead %.1f; max |diff| %.3e (scale %.2f)  err=%s\n", NQ, mode, nblk,
           cyc[0], cyc[1], worst, scale, cudaGetErrorString(cudaGetLastError()));
    cudaFree(din); cudaFree(dout[0]); cudaFree(dout[1]); cudaFree(dclk);
}

int main(int argc, char** argv) {
    const int nblk = argc > 1 ? atoi(argv[1]) : 512, iters = argc > 2 ? atoi(argv[2]) : 200;
    for (int mode = 0; mode < 2; ++mode) {
        run<1>(nblk, iters, mode);
        run<2>(nblk, iters, mode);
        run<4>(nblk, iters, mode);
        run<6>(nblk, iters, mode);
    }
    return 0;
}
