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

// 1 / sqrt(d) for a positive, normal pivot: hardware seed + two Newton steps (a third of the IEEE sqrt + divide sequence)
__device__ __forceinline__ double pivot_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
#pragma unroll
    for (int i = 0; i < 2; ++i) r = fma(r, fma(-hx * r, r, 0.5), r);
    return r;
}
__device__ __forceinline__ float pivot_rsqrt(float x) { return rsqrtf(x); }

// Counter-based standard normals (Philox4x32-10 + Box-Muller): variate i of variate array `stream` of a call is a pure
// function of (seed, offset, stream, i).  Counter = [lo, hi of offset + i / 2, stream, 0], key = [lo, hi of seed]; the four
// 32-bit outputs give two uniforms in (0, 1) with 53 random bits each; variate i is element i & 1 of the Box-Muller pair.
// Restated for the parity tests in oracle/philox.py.
__device__ __forceinline__ double philox_normal(const Rng rng, const unsigned stream, const unsigned long long i) {
    const unsigned long long ctr = rng.offset + (i >> 1);
    unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = stream, c3 = 0u;
    unsigned k0 = (unsigned)rng.seed, k1 = (unsigned)(rng.seed >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const double u1 = ((double)((((unsigned long long)c0 << 32) | c1) >> 11) + 0.5) * 0x1.0p-53;
    const double u2 = ((double)((((unsigned long long)c2 << 32) | c3) >> 11) + 0.5) * 0x1.0p-53;
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    return rad * ((i & 1) ? sn : cs);
}
template <typename R>
__device__ __forceinline__ R variate(const R* __restrict__ z, const Rng rng, const unsigned stream, const unsigned long long i) {
    return z ? z[i] : (R)philox_normal(rng, stream, i);
}

template <typename R, int G>
__global__ void __launch_bounds__(128) sample_kernel(const R* __restrict__ mu, const R* __restrict__ cov,
                                                    const R* __restrict__ z, const Rng rng, R* __restrict__ x,
                                                    int* __restrict__ status, long long N, int n, int m) {
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
            const R inv = ok ? pivot_rsqrt(d) : R(0);
            l[j] = i >= j ? acc * inv : R(0);  // the diagonal: d / sqrt(d)
        }
    }
    const R zi = (live && i < m) ? variate<R>(z, rng, 0u, (unsigned long long)(mm * m + i)) : R(0);
    R acc = (live && i < m) ? mu[mm * m + i] : R(0);
#pragma unroll
    for (int k = 0; k < G; ++k)
        if (k < m) acc = fma(l[k], __shfl_sync(gmask, zi, base + k), acc);
    if (live && i < m) x[mm * m + i] = acc;
    if (live && status && bad && i == 0) atomicOr(status + mm / n, bad);
}

template <typename R>
int launch(const R* mu, const R* cov, const R* z, Rng rng, R* x, int* status, long long N, int n, int m,
           cudaStream_t stream, const char** err) {
    if (m > 32) { *err = "bkf_mvn_sample: m > 32 is not supported"; return BKF_ERR_UNSUPPORTED; }
    const int G = m <= 8 ? 8 : (m <= 16 ? 16 : 32);
    const long long warps = (N + (32 / G) - 1) / (32 / G);
    const long long blocks = (warps + 3) / 4;
    if (blocks == 0) return BKF_OK;
    if (G == 8) sample_kernel<R, 8><<<(unsigned)blocks, 128, 0, stream>>>(mu, cov, z, rng, x, status, N, n, m);
    else if (G == 16) sample_kernel<R, 16><<<(unsigned)blocks, 128, 0, stream>>>(mu, cov, z, rng, x, status, N, n, m);
    else sample_kernel<R, 32><<<(unsigned)blocks, 128, 0, stream>>>(mu, cov, z, rng, x, status, N, n, m);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

// ----------------------------------------------------------------------------------------------------
// Unconditional simulation of the state-space model: LinearGaussianStateSpace.rv_op (filters/distributions.py:181-293).
//     x_0 ~ N(a0, P0),  y_0 ~ N(Z x_0, H)           (:262-263 -- no d in the first observation mean, as the reference)
//     x_t = c + T x_{t-1} + R eps_t,  eps_t ~ N(0, Q);   y_t = d + Z x_t + eta_t,  eta_t ~ N(0, H)      (:247-254)
// with every normal draw written as (factor) x (caller-supplied standard normal variates), factors = Cholesky with
// dropped semidefinite pivots as above.  One warp per series, matrices in the warp's shared-memory arena, lane i owns
// row i of the matrix-vector products.
template <typename R>
__device__ int chol_psd_warp(R* A, int k, int lane) {
    R dmax = 0;
    for (int i = 0; i < k; ++i) dmax = fmax(dmax, A[i * k + i]);
    const R eps = sizeof(R) == 8 ? R(2.220446049250313e-16) : R(1.1920929e-7f);
    const R tol = R(64 * k) * eps * dmax, neg = (sizeof(R) == 8 ? R(1e-8) : R(1e-4)) * dmax;
    int bad = 0;
    for (int j = 0; j < k; ++j) {
        R d = A[j * k + j];
        for (int q = 0; q < j; ++q) d = fma(-A[j * k + q], A[j * k + q], d);
        const bool ok = d > tol;
        if (d < -neg) bad = BKF_STATUS_NOT_PD;
        const R ljj = ok ? sqrt(d) : R(0), inv = ok ? R(1) / ljj : R(0);
        __syncwarp();
        for (int i = j + 1 + lane; i < k; i += 32) {
            R tv = A[i * k + j];
            for (int q = 0; q < j; ++q) tv = fma(-A[i * k + q], A[j * k + q], tv);
            A[i * k + j] = tv * inv;
        }
        if (lane == 0) A[j * k + j] = ljj;
        for (int i = lane; i < j; i += 32) A[i * k + j] = 0;
        __syncwarp();
    }
    return bad;
}

__host__ __device__ inline size_t sim_arena_elems(int m, int p, int r) {
    size_t tot = /*L0,T*/ 2 * (size_t)m * m + /*Z*/ (size_t)p * m + /*RL*/ (size_t)m * r + /*LH*/ (size_t)p * p +
                 /*LQ*/ (size_t)r * r + /*a,an,c*/ 3 * (size_t)m + /*d*/ (size_t)p + /*zs*/ (size_t)r + /*zo*/ (size_t)p;
    return (tot + 1) & ~(size_t)1;
}

template <typename R>
__global__ void simulate_kernel(KArgs<R> g, const R* __restrict__ z_x0, const R* __restrict__ z_y0,
                                const R* __restrict__ z_state, const R* __restrict__ z_obs, const Rng rng,
                                R* __restrict__ xs,
                                R* __restrict__ ys, int wpc, size_t arena) {
    extern __shared__ double smem_raw[];
    R* s_ = reinterpret_cast<R*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * wpc + warp;
    if (b >= g.B) return;
    const int m = g.m, p = g.p, r = g.r, n = g.n;
    R* L0 = s_ + (size_t)warp * arena;
    R *Tm = L0 + m * m, *Zm = Tm + m * m, *RL = Zm + p * m, *LH = RL + m * r, *LQ = LH + p * p, *a = LQ + r * r,
      *an = a + m, *cv = an + m, *dv = cv + m, *zs = dv + p, *zo = zs + r;
    int st = 0;
    // parameters of step t (sequences of length n, distributions.py:213-218); first: everything, else only what carries
    // a time axis -- with the factors of H_t, Q_t and R_t L_Q,t redone
    auto load_step = [&](int t, bool first) {
        auto want = [&](int which) { return first || ((g.tv_mask >> which) & 1u); };
        auto ld = [&](R* dst, const R* src, int which, size_t sz) {
            const R* q = param_ptr_t(src, g.shared_mask, g.tv_mask, which, b, t, n, sz);
            for (int i = lane; i < (int)sz; i += 32) dst[i] = q[i];
        };
        if (want(BKF_PARAM_T)) ld(Tm, g.T, BKF_PARAM_T, (size_t)m * m);
        if (want(BKF_PARAM_Z)) ld(Zm, g.Z, BKF_PARAM_Z, (size_t)p * m);
        if (want(BKF_PARAM_C)) ld(cv, g.c, BKF_PARAM_C, (size_t)m);
        if (want(BKF_PARAM_D)) ld(dv, g.d, BKF_PARAM_D, (size_t)p);
        if (want(BKF_PARAM_H)) ld(LH, g.H, BKF_PARAM_H, (size_t)p * p);
        if (want(BKF_PARAM_Q)) ld(LQ, g.Q, BKF_PARAM_Q, (size_t)r * r);
        __syncwarp();
        if (want(BKF_PARAM_H)) st |= chol_psd_warp(LH, p, lane);
        if (want(BKF_PARAM_Q)) st |= chol_psd_warp(LQ, r, lane);
        if (want(BKF_PARAM_Q) || want(BKF_PARAM_R)) {
            const R* Rg = param_ptr_t(g.Rm, g.shared_mask, g.tv_mask, BKF_PARAM_R, b, t, n, (size_t)m * r);
            for (int idx = lane; idx < m * r; idx += 32) {
                const int i = idx / r, k = idx - i * r;
                R acc = 0;
                for (int l = k; l < r; ++l) acc = fma(Rg[i * r + l], LQ[l * r + k], acc);
                RL[idx] = acc;
            }
        }
        __syncwarp();
    };
    {
        const R* q = param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)m * m);
        for (int i = lane; i < m * m; i += 32) L0[i] = q[i];
        __syncwarp();
        st |= chol_psd_warp(L0, m, lane);
    }
    load_step(0, true);
    const R* a0 = param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)m);
    R* xb = xs + (size_t)b * (n + 1) * m;
    R* yb = ys + (size_t)b * (n + 1) * p;
    // x_0 = a0 + L0 z ;  y_0 = Z x_0 + LH z'
    for (int i = lane; i < m; i += 32) an[i] = variate<R>(z_x0, rng, 1u, (unsigned long long)b * m + i);
    for (int i = lane; i < p; i += 32) zo[i] = variate<R>(z_y0, rng, 2u, (unsigned long long)b * p + i);
    __syncwarp();
    for (int i = lane; i < m; i += 32) {
        R acc = a0[i];
        for (int k = 0; k <= i; ++k) acc = fma(L0[i * m + k], an[k], acc);
        a[i] = acc;
        xb[i] = acc;
    }
    __syncwarp();
    for (int o = lane; o < p; o += 32) {
        R acc = 0;
        for (int j = 0; j < m; ++j) acc = fma(Zm[o * m + j], a[j], acc);
        for (int k = 0; k <= o; ++k) acc = fma(LH[o * p + k], zo[k], acc);
        yb[o] = acc;
    }
    for (int t = 0; t < n; ++t) {
        __syncwarp();
        if (g.tv_mask && t > 0) load_step(t, false);
        for (int i = lane; i < r; i += 32) zs[i] = variate<R>(z_state, rng, 3u, ((unsigned long long)b * n + t) * r + i);
        for (int i = lane; i < p; i += 32) zo[i] = variate<R>(z_obs, rng, 4u, ((unsigned long long)b * n + t) * p + i);
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            R acc = cv[i];
            for (int j = 0; j < m; ++j) acc = fma(Tm[i * m + j], a[j], acc);
            for (int k = 0; k < r; ++k) acc = fma(RL[i * r + k], zs[k], acc);
            an[i] = acc;
            xb[(size_t)(t + 1) * m + i] = acc;
        }
        __syncwarp();
        for (int o = lane; o < p; o += 32) {
            R acc = dv[o];
            for (int j = 0; j < m; ++j) acc = fma(Zm[o * m + j], an[j], acc);
            for (int k = 0; k <= o; ++k) acc = fma(LH[o * p + k], zo[k], acc);
            yb[(size_t)(t + 1) * p + o] = acc;
        }
        for (int i = lane; i < m; i += 32) a[i] = an[i];
    }
    if (lane == 0 && g.status) g.status[b] = st;
}

template <typename R>
int launch_simulate(const KArgs<R>& g, const R* z_x0, const R* z_y0, const R* z_state, const R* z_obs, Rng rng, R* xs,
                    R* ys, cudaStream_t stream, const char** err) {
    const size_t arena = sim_arena_elems(g.m, g.p, g.r), arena_bytes = arena * sizeof(R);
    if (arena_bytes > 227 * 1024) { *err = "problem too large for the simulation kernel (m, p, r)"; return BKF_ERR_UNSUPPORTED; }
    int wpc = 4;
    while (wpc > 1 && arena_bytes * wpc > 96 * 1024) wpc >>= 1;
    const size_t smem = arena_bytes * wpc;
    auto kern = simulate_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<(g.B + wpc - 1) / wpc, 32 * wpc, smem, stream>>>(g, z_x0, z_y0, z_state, z_obs, rng, xs, ys, wpc, arena);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

template int launch_simulate<double>(const KArgs<double>&, const double*, const double*, const double*, const double*,
                                     Rng, double*, double*, cudaStream_t, const char**);
template int launch_simulate<float>(const KArgs<float>&, const float*, const float*, const float*, const float*, Rng,
                                    float*, float*, cudaStream_t, const char**);

template int launch<double>(const double*, const double*, const double*, Rng, double*, int*, long long, int, int,
                            cudaStream_t, const char**);
template int launch<float>(const float*, const float*, const float*, Rng, float*, int*, long long, int, int, cudaStream_t,
                           const char**);

}  // namespace smp
}  // namespace bkf
