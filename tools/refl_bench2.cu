// Elimination experiments on the Householder panel loop (which part of a reflector costs what): variants of
// sqw::panel_reflectors with pieces switched off (results are then wrong; only the time matters).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pymc_experimental_b200/csrc/bkf_sqrt_warp.cuh"
using namespace bkf::sqw;

// E bits: 1 no vote/branch, 2 trivial scalars, 4 no Gram store, 8 no akc shuffle, 16 no smem broadcast (xb = pa),
//         32 red_t instead of red4_mma, 64 no ss shuffle (ss = dot), 128 fully unrolled
template <int NQ, int E>
__device__ __forceinline__ void panel_x(double2 (&pa)[NQ], double2& Rd, double& tauv, double* tq, double* xbuf, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    tauv = 0.0;
    double scalv = 1.0;
#pragma unroll (E & 128 ? 8 : 2)
    for (int k = 0; k < 8; ++k) {
        double2 xb[NQ];
        double alpha;
        if constexpr (E & 16) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) xb[q] = pa[q];
            alpha = pa[0].x;
        } else {
            if (g == k) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) reinterpret_cast<double2*>(xbuf)[q * 4 + t] = pa[q];
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < NQ; ++q) xb[q] = reinterpret_cast<const double2*>(xbuf)[q * 4 + t];
            alpha = xbuf[k];
        }
        double akc;
        if constexpr (E & 8) akc = pa[0].x;
        else akc = __shfl_sync(FULL, (k & 1) ? pa[0].y : pa[0].x, 4 * g + (k >> 1));
        double2 xm = xb[0];
        if (2 * t <= k) xm.x = 0.0;
        if (2 * t + 1 <= k) xm.y = 0.0;
        double d0 = xm.x * pa[0].x, d1 = xm.y * pa[0].y;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            d0 = fma(xb[q].x, pa[q].x, d0);
            d1 = fma(xb[q].y, pa[q].y, d1);
        }
        const double dsum = d0 + d1;
        if constexpr (!(E & 1)) {
            const bool nz = __any_sync(FULL, g == k && dsum != 0.0);
            if (!nz) {
                if (g < k && t == 0) tq[g * 8 + k] = 0.0;
                continue;
            }
        }
        double dot;
        if constexpr (E & 32) dot = red_t(dsum);
        else dot = red4_mma(dsum);
        double ss;
        if constexpr (E & 64) ss = dot;
        else ss = __shfl_sync(FULL, dot, 4 * k);
        Refl R;
        if constexpr (E & 2) {
            const double x2 = fma(alpha, alpha, ss);
            R.beta = x2; R.tau = x2 * 0.5; R.scal = x2 * 0.25; R.ts = x2 * 0.125;
        } else {
            R = reflector(fma(alpha, alpha, ss), alpha);
        }
        const double zz = fma(R.scal, dot, akc);
        if constexpr (!(E & 4)) {
            if (g < k && t == 0) tq[g * 8 + k] = scalv * zz;
        }
        const bool piv = g == k;
        const double w = g > k ? R.ts * zz : 0.0;
        double2 ud = xm;
        if (2 * t == k) ud.x = alpha - R.beta;
        if (2 * t + 1 == k) ud.y = alpha - R.beta;
        {
            const double nx = fma(-w, ud.x, pa[0].x), ny = fma(-w, ud.y, pa[0].y);
            pa[0].x = (piv && 2 * t == k) ? R.beta : nx;
            pa[0].y = (piv && 2 * t + 1 == k) ? R.beta : ny;
        }
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            pa[q].x = fma(-w, xb[q].x, pa[q].x);
            pa[q].y = fma(-w, xb[q].y, pa[q].y);
        }
        if (piv) {
            tauv = R.tau;
            scalv = R.scal;
        }
    }
    const double2 r = pa[0];
    Rd = make_double2(2 * t <= g ? r.x : 0.0, 2 * t + 1 <= g ? r.y : 0.0);
    pa[0].x = 2 * t > g ? r.x * scalv : (2 * t == g ? 1.0 : 0.0);
    pa[0].y = 2 * t + 1 > g ? r.y * scalv : (2 * t + 1 == g ? 1.0 : 0.0);
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
        pa[q].x *= scalv;
        pa[q].y *= scalv;
    }
}

template <int NQ, int E>
__global__ void __launch_bounds__(32) kern(const double* __restrict__ in, double* __restrict__ out, long long* clk, int iters) {
    __shared__ __align__(16) double tq[72 + 64 + 64];
    const int lane = threadIdx.x;
    double2 pa0[NQ];
    for (int q = 0; q < NQ; ++q) pa0[q] = reinterpret_cast<const double2*>(in)[(blockIdx.x * NQ + q) * 32 + lane];
    double2 pa[NQ], Rd = make_double2(0, 0);
    double tauv = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) pa[q] = pa0[q];
        if (it) pa[0].x += Rd.x * 1e-30;
        panel_x<NQ, E>(pa, Rd, tauv, tq, tq + 72, lane);
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) clk[blockIdx.x] = t1 - t0;
    double* o = out + (size_t)blockIdx.x * (NQ + 1) * 64 + 64;
    for (int q = 0; q < NQ; ++q) reinterpret_cast<double2*>(o)[q * 32 + lane] = pa[q];
    reinterpret_cast<double2*>(out + (size_t)blockIdx.x * (NQ + 1) * 64)[lane] = Rd;
    if ((lane & 3) == 0) o[NQ * 64 - 8 + (lane >> 2)] += 0.0 * tauv;
}

template <int NQ, int E>
void run(int nblk, int iters, const char* what) {
    const size_t nin = (size_t)nblk * NQ * 64, nout = (size_t)nblk * (NQ + 1) * 64;
    std::vector<double> h(nin);
    srand(7);
    for (auto& x : h) x = (rand() / (double)RAND_MAX - 0.5);
    double *din, *dout;
    long long* dclk;
    cudaMalloc(&din, nin * 8); cudaMalloc(&dout, nout * 8); cudaMalloc(&dclk, nblk * 8);
    cudaMemcpy(din, h.data(), nin * 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) { kern<NQ, E><<<nblk, 32>>>(din, dout, dclk, iters); cudaDeviceSynchronize(); }
    std::vector<long long> c(nblk);
    cudaMemcpy(c.data(), dclk, nblk * 8, cudaMemcpyDeviceToHost);
    double s = 0;
    for (auto x : c) s += x;
    printf("NQ=%d E=%3d %-46s %.1f cycles/reflector  (%s)\n", NQ, E, what, s / nblk / iters / 8.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(din); cudaFree(dout); cudaFree(dclk);
}

int main(int argc, char** argv) {
    const int nblk = argc > 1 ? atoi(argv[1]) : 512, iters = 200;
#define RUNS(NQ) \
    run<NQ, 0>(nblk, iters, "baseline"); \
    run<NQ, 1>(nblk, iters, "no vote / branch"); \
    run<NQ, 2>(nblk, iters, "trivial scalars"); \
    run<NQ, 4>(nblk, iters, "no Gram store"); \
    run<NQ, 8>(nblk, iters, "no akc shuffle"); \
    run<NQ, 16>(nblk, iters, "no smem broadcast"); \
    run<NQ, 32>(nblk, iters, "red_t instead of DMMA reduce"); \
    run<NQ, 64>(nblk, iters, "no ss shuffle"); \
    run<NQ, 128>(nblk, iters, "fully unrolled"); \
    run<NQ, 1 + 4 + 8 + 64>(nblk, iters, "no vote, Gram, akc, ss"); \
    run<NQ, 1 + 2 + 4 + 8 + 16 + 64>(nblk, iters, "only dots + reduce + update");
    RUNS(6)
    RUNS(2)
    return 0;
}
