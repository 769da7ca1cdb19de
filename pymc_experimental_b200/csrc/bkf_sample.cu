// bkf_sample.cu -- draws from N(mu[b,t], cov[b,t]) for every (series, step): SURVEY.md section 8 row f-3.
//
// Reference: SequenceMvNormal.rv_op (filters/distributions.py:411-443) scans MvNormalSVD (:52-87) over the time axis --
// one multivariate normal per step, the SVD method chosen because smoothed covariances may be rank deficient; it is the
// consumer of the smoother's output in sample_conditional_posterior (core/statespace.py:1119-1150).
//
// A draw is  x = mu + L z  with  L L^T = cov  and z standard normal.  Which square root is used does not change the
// distribution (numpy takes sqrt(s) V^T from the SVD); this kernel takes the lower Cholesky factor with SEMIDEFINITE
// PIVOTS DROPPED -- a pivot below 64 m eps max(diag) zeroes its column, which for a PSD matrix is exact up to that
// tolerance -- so it is as robust to low rank as the SVD at a tenth of the cost.  z is supplied by the caller (the host
// side draws it from its own generator), so the call is deterministic; parity with the reference is distributional.
//
// Mapping: a group of G = 8 / 16 / 32 lanes per matrix, lane i owns row i of the factor in registers; column j needs
// row j of L, broadcast by shuffles (m^2/2 per matrix).  No shared memory.
#include "bkf_common.cuh"
#include "bkf_launch.h"

namespace bkf {
namespace smp {

template <typename R, int G>
__global__ void __launch_bounds__(128) sample_kernel(const R* __restrict__ mu, const R* __restrict__ cov,
                                                    const R* __restrict__ z, R* __restrict__ x, int* __restrict__ status,
                                                    long long N, int n, int m) {
    constexpr int GPW = 32 / G;  // matrices per warp
    const int lane = threadIdx.x & 31, i = lane % G, sub = lane / G;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long mat = warp * GPW + sub;
    const bool live = mat < N;
    const long long mm = live ? mat : 0;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
    const int base = sub * G;
    const R* C = cov + mm * (long long)m * m;
    R c[G], l[G];
    R dmax = 0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        c[j] = (i < m && j <= i) ? C[(long long)i * m + j] : R(0);
        l[j] = 0;
        if (j == i) dmax = c[j];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(gmask, dmax, o));
    const R eps = sizeof(R) == 8 ? R(2.220446049250313e-16) : R(1.1920929e-7f);
    const R tol = R(64 * m) * eps * dmax;                       // pivots below it are rounding residue of a PSD matrix
    const R neg = (sizeof(R) == 8 ? R(1e-8) : R(1e-4)) * dmax;  // a pivot below -neg: the matrix is not PSD
    int bad = 0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        if (j < m) {  // uniform
            R acc = c[j];
#pragma unroll
            for (int k = 0; k < j; ++k) acc = fma(-l[k], __shfl_sync(gmask, l[k], base + j), acc);
            const R d = __shfl_sync(gmask, acc, base + j);
            const bool ok = d > tol;
            if (d < -neg) bad = BKF_STATUS_NOT_PD;
            const R ljj = ok ? sqrt(d) : R(0);
            const R inv = ok ? R(1) / ljj : R(0);
            l[j] = i == j ? ljj : (i > j ? acc * inv : R(0));
        }
    }
    const R zi = (live && i < m) ? z[mm * m + i] : R(0);
    R acc = (live && i < m) ? mu[mm * m + i] : R(0);
#pragma unroll
    for (int k = 0; k < G; ++k)
        if (k < m) acc = fma(l[k], __shfl_sync(gmask, zi, base + k), acc);
    if (live && i < m) x[mm * m + i] = acc;
    if (live && status && bad && i == 0) atomicOr(status + mm / n, bad);
}

template <typename R>
int launch(const R* mu, const R* cov, const R* z, R* x, int* status, long long N, int n, int m, cudaStream_t stream,
           const char** err) {
    if (m > 32) { *err = "bkf_mvn_sample: m > 32 is not supported"; return BKF_ERR_UNSUPPORTED; }
    const int G = m <= 8 ? 8 : (m <= 16 ? 16 : 32);
    const long long warps = (N + (32 / G) - 1) / (32 / G);
    const long long blocks = (warps + 3) / 4;
    if (blocks == 0) return BKF_OK;
    if (G == 8) sample_kernel<R, 8><<<(unsigned)blocks, 128, 0, stream>>>(mu, cov, z, x, status, N, n, m);
    else if (G == 16) sample_kernel<R, 16><<<(unsigned)blocks, 128, 0, stream>>>(mu, cov, z, x, status, N, n, m);
    else sample_kernel<R, 32><<<(unsigned)blocks, 128, 0, stream>>>(mu, cov, z, x, status, N, n, m);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

template int launch<double>(const double*, const double*, const double*, double*, int*, long long, int, int, cudaStream_t,
                            const char**);
template int launch<float>(const float*, const float*, const float*, float*, int*, long long, int, int, cudaStream_t,
                           const char**);

}  // namespace smp
}  // namespace bkf
