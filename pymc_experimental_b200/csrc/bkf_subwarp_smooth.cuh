// bkf_subwarp_smooth.cuh -- RTS smoother in the sub-warp column layout of bkf_subwarp.cuh (5 <= m <= 16, f64).
//
// kalman_smoother.py:107-126 per step (t = n-2 .. 0), with (a, P) = filtered (a_t|t, P_t|t):
//   a_hat = T a (no intercept, quirk Q8);  P_hat = sym(T P T^T) + sym(R Q R^T) + eps I
//   G = (pinv(P_hat) T P)^T;  a_s = a + G (a_s+ - a_hat);  P_s = P + sym(G (P_s+ - P_hat) G^T) + eps I + eps0 I
//
// pinv policy: P_hat is symmetric positive definite after the jitter, so pinv(P_hat) equals
// the inverse whenever no singular value falls under NumPy's 1e-15 cut-off.  This kernel solves P_hat G^T = T P
// by Cholesky (G^T = P_hat^{-1} (T P) because P, P_hat are symmetric).  When a pivot is not positive or
// min(pivot)/max(pivot) < 1e-13 -- the regime where the cut-off could act -- the series is flagged
// (BKF_STATUS_SMOOTHER_ILL) and the host re-runs the call through the generic kernel's eigen-decomposition pinv.
#pragma once
#include "bkf_subwarp.cuh"

namespace bkf {
namespace sw {

template <int M>
struct SmSmem {
    using C = Cfg<M>;
    static constexpr int T_cm = 0;               // T column-major
    static constexpr int A = T_cm + C::MAT;      // scratch matrix (X, D, ...)
    static constexpr int Bm = A + C::MAT;        // scratch matrix (Y, V^T, ...)
    static constexpr int RQR = Bm + C::MAT;      // sym(R Q R^T) columns
    static constexpr int L = RQR + C::MAT;       // Cholesky factor of P_hat, column-major
    static constexpr int IN0 = L + C::MAT;       // filtered record [a | P], double buffered
    static constexpr int IN1 = IN0 + C::DENSE_PAD;
    static constexpr int OUT = IN1 + C::DENSE_PAD;
    static constexpr int VA = OUT + C::DENSE_PAD;
    static constexpr int VB = VA + C::VEC;
    static constexpr int VD = VB + C::VEC;       // reciprocal Cholesky diagonal
    static constexpr int TOTAL = ((VD + C::VEC + 15) & ~15) + 16 / C::SPW;
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}

template <int M>
__global__ void __launch_bounds__(64) smoother_kernel(KArgs<double> g, double jitter0, int* ill_list) {
    using C = Cfg<M>;
    using S = SmSmem<M>;
    constexpr int LPS = C::LPS, LD = C::LD;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPS, j = lane % LPS;
    const int series = (blockIdx.x * (blockDim.x >> 5) + warp) * C::SPW + sub;
    const bool valid = series < g.B;
    const int b = valid ? series : g.B - 1;
    double* sm = smem + (size_t)((warp * C::SPW + sub)) * S::TOTAL;
    for (int i = j; i < S::TOTAL; i += LPS) sm[i] = 0.0;  // idle lanes (M < LPS) read rows nobody writes: see bkf_subwarp.cuh
    __syncwarp();
    const int n = g.n;
    const bool act = j < M;

    const double* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)M * M);
    const double* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * g.r);
    const double* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)g.r * g.r);
    double Trow[M];
#pragma unroll
    for (int k = 0; k < M; ++k) Trow[k] = act ? Tg[j * M + k] : 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) sm[S::T_cm + k * LD + j] = Trow[k];
    {
        const int r = g.r;
#pragma unroll 1
        for (int i = 0; i < M; ++i) {
            double s_ij = 0, s_ji = 0;
            if (act) {
                for (int k = 0; k < r; ++k) {
                    double rq_i = 0, rq_j = 0;
                    for (int l = 0; l < r; ++l) {
                        rq_i = fma(Rg[i * r + l], Qg[l * r + k], rq_i);
                        rq_j = fma(Rg[j * r + l], Qg[l * r + k], rq_j);
                    }
                    s_ij = fma(rq_i, Rg[j * r + k], s_ij);
                    s_ji = fma(rq_j, Rg[i * r + k], s_ji);
                }
            }
            sm[S::RQR + j * LD + i] = 0.5 * (s_ij + s_ji);
        }
    }
    // zero the padding rows/columns of the scratch matrices once (inactive lanes never produce NaNs)
#pragma unroll
    for (int i = 0; i < LPS; ++i) {
        sm[S::A + j * LD + i] = 0.0;
        sm[S::Bm + j * LD + i] = 0.0;
        sm[S::L + j * LD + i] = 0.0;
    }
    __syncwarp();

    const double* fa = g.sm_filt_a + (size_t)b * n * M;
    const double* fP = g.sm_filt_P + (size_t)b * n * M * M;
    auto prefetch = [&](int t, int buf) {
        double* dst = sm + (buf ? S::IN1 : S::IN0);
        const double* pa = fa + (size_t)t * M;
        const double* pP = fP + (size_t)t * M * M;
        for (int q = j; q < M; q += LPS) cp_async8(dst + q, pa + q);
        for (int q = j; q < M * M; q += LPS) cp_async8(dst + M + q, pP + q);
        cp_async_commit();
    };
    auto emit = [&](int t, double as_j, const double (&Ps)[M]) {
        double* o = sm + S::OUT;
        if (act) {
            o[j] = as_j;
#pragma unroll
            for (int i = 0; i < M; ++i) o[M + j * M + i] = Ps[i];  // symmetric: row j == column j
        }
        __syncwarp();
        if (valid) {
            if (g.sm_a)
                for (int q = j; q < M; q += LPS) g.sm_a[((size_t)b * n + t) * M + q] = o[q];
            if (g.sm_P)
                for (int q = j; q < M * M; q += LPS) g.sm_P[((size_t)b * n + t) * M * M + q] = o[M + q];
        }
        __syncwarp();
    };

    // t = n-1: smoothed == filtered
    prefetch(n - 1, (n - 1) & 1);
    cp_async_wait<0>();
    __syncwarp();
    double as_j, Ps[M];
    {
        const double* in = sm + (((n - 1) & 1) ? S::IN1 : S::IN0);
        as_j = act ? in[j] : 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) Ps[i] = act ? in[M + i * M + j] : 0.0;  // column j, no symmetry assumed here
        if (valid) {
            if (g.sm_a && act) g.sm_a[((size_t)b * n + n - 1) * M + j] = as_j;
            if (g.sm_P && act)
#pragma unroll
                for (int i = 0; i < M; ++i) g.sm_P[((size_t)b * n + n - 1) * M * M + i * M + j] = Ps[i];
        }
    }
    int ill = 0;
    if (n >= 2) prefetch(n - 2, (n - 2) & 1);
    for (int t = n - 2; t >= 0; --t) {
        cp_async_wait<0>();
        __syncwarp();
        if (t > 0) prefetch(t - 1, (t - 1) & 1);
        const double* in = sm + ((t & 1) ? S::IN1 : S::IN0);
        const double a_j = act ? in[j] : 0.0;
        double Pc[M];
#pragma unroll
        for (int i = 0; i < M; ++i) Pc[i] = act ? in[M + j * M + i] : 0.0;  // filtered covariances are symmetric
        // w = a_s+ - T a
        double av[M];
        load_vec<M>(in, av);
        const double w_j = as_j - dot<M>(Trow, av);
        if (act) sm[S::VA + j] = w_j;
        // X = T P ; Y = X T^T ; P_hat = sym(Y) + RQR + eps I
        double X[M];
        prod<M, LD>(sm + S::T_cm, Pc, X);
        store_vec<M>(sm + S::A + j * LD, X);
        __syncwarp();
        double Y[M];
        prod<M, LD>(sm + S::A, Trow, Y);
        store_vec<M>(sm + S::Bm + j * LD, Y);
        __syncwarp();
        double Ph[M];
        {
            double rq[M];
            load_vec<M>(sm + S::RQR + j * LD, rq);
#pragma unroll
            for (int i = 0; i < M; ++i) {
                Ph[i] = 0.5 * (Y[i] + sm[S::Bm + i * LD + j]) + rq[i];
                if (i == j) Ph[i] += g.jitter;
            }
        }
        // ---- Cholesky of P_hat: lane j owns column j of the trailing matrix ----
        double Lc[M];
#pragma unroll
        for (int i = 0; i < M; ++i) Lc[i] = Ph[i];
        double dmin = 1e300, dmax = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
            if (j == k) {
                const double piv = Lc[k];
                const double dk = sqrt(piv);
                const double rk = 1.0 / dk;
                sm[S::VD + k] = rk;
                sm[S::VB + 0] = piv;
#pragma unroll
                for (int i = 0; i < M; ++i) Lc[i] = (i == k) ? dk : (i > k ? Lc[i] * rk : 0.0);
                store_vec<M>(sm + S::L + k * LD, Lc);
            }
            __syncwarp();
            const double piv = sm[S::VB + 0];
            if (!(piv > 0.0)) ill = 1;
            dmin = fmin(dmin, piv);
            dmax = fmax(dmax, piv);
            if (j > k) {
                double colk[M];
                load_vec<M>(sm + S::L + k * LD, colk);
                const double lkj = sm[S::L + k * LD + j];
#pragma unroll
                for (int i = k + 1; i < M; ++i) Lc[i] = fma(-colk[i], lkj, Lc[i]);
            }
            __syncwarp();
        }
        if (dmin < 1e-13 * dmax) ill = 1;
        // ---- solve P_hat G^T = X : lane j solves for column j of G^T (= row j of G) ----
        double rd[M];
        load_vec<M>(sm + S::VD, rd);
        double Gr[M];
#pragma unroll
        for (int i = 0; i < M; ++i) Gr[i] = X[i];
#pragma unroll
        for (int q = 0; q < M; ++q) {  // forward substitution, column oriented
            double colq[M];
            load_vec<M>(sm + S::L + q * LD, colq);
            Gr[q] *= rd[q];
#pragma unroll
            for (int i = q + 1; i < M; ++i) Gr[i] = fma(-colq[i], Gr[q], Gr[i]);
        }
#pragma unroll
        for (int q = M - 1; q >= 0; --q) {  // backward substitution with L^T
            double colq[M];
            load_vec<M>(sm + S::L + q * LD, colq);
            double s = Gr[q];
#pragma unroll
            for (int i = q + 1; i < M; ++i) s = fma(-colq[i], Gr[i], s);
            Gr[q] = s * rd[q];
        }
        // ---- a_s = a + G w ----
        {
            double wv[M];
            load_vec<M>(sm + S::VA, wv);
            as_j = a_j + dot<M>(Gr, wv);
        }
        // ---- P_s = P + sym(G D G^T) + eps I + eps0 I,  D = P_s+ - P_hat ----
        double D[M];
#pragma unroll
        for (int i = 0; i < M; ++i) D[i] = Ps[i] - Ph[i];
        __syncwarp();
        store_vec<M>(sm + S::A + j * LD, D);  // D column-major (symmetric)
        __syncwarp();
        double V[M];  // V = D G^T, column j
        prod<M, LD>(sm + S::A, Gr, V);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < M; ++i) sm[S::Bm + i * LD + j] = V[i];  // store V^T column-major: S[k*LD+i] = V[k,i]
        __syncwarp();
        double Mt[M];  // (G D G^T)[j, :]
        prod<M, LD>(sm + S::Bm, Gr, Mt);
        __syncwarp();
        store_vec<M>(sm + S::A + j * LD, Mt);  // S[j*LD + i] = M2[j,i]
        __syncwarp();
#pragma unroll
        for (int i = 0; i < M; ++i) {
            double v = Pc[i] + 0.5 * (sm[S::A + i * LD + j] + Mt[i]);  // 0.5 (M2[i,j] + M2[j,i])
            if (i == j) v = (v + g.jitter) + jitter0;
            Ps[i] = act ? v : 0.0;
        }
        __syncwarp();
        emit(t, as_j, Ps);
    }
    if (ill && valid && j == 0) {
        if (g.status) g.status[b] |= BKF_STATUS_SMOOTHER_ILL;
        ill_list[1 + atomicAdd(ill_list, 1)] = b;  // ill_list[0]: how many, then which
    }
}

template <int M>
int launch_smoother(const KArgs<double>& g, double jitter0, int* ill_list, cudaStream_t stream, const char** err) {
    using C = Cfg<M>;
    const int warps = 2;
    const int series_per_cta = warps * C::SPW;
    const int grid = (g.B + series_per_cta - 1) / series_per_cta;
    const size_t smem = (size_t)series_per_cta * SmSmem<M>::TOTAL * sizeof(double);
    auto kern = smoother_kernel<M>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<grid, 32 * warps, smem, stream>>>(g, jitter0, ill_list);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace sw
}  // namespace bkf
