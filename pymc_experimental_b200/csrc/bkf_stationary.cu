// bkf_stationary.cu -- stationary initialisation of a batch of state-space models:
//
//     a0 = (I - T)^{-1} c                         models/SARIMAX.py:369-381, VARMAX.py:379-393 (pt.linalg.solve)
//     P0 = solve_discrete_lyapunov(T, R Q R^T)    same lines; ETS.py:453-466   (P0 = T P0 T^T + R Q R^T)
//
// and the reverse-mode pull-back of both (PyTensor's SolveDiscreteLyapunov L_op: Xbar = lyap(A^T, dX),
// Abar = Xbar A X^T + Xbar^T A X, Qbar = Xbar; Solve L_op for a0).
//
// Method: both are Neumann series in T (the model is stationary exactly when they converge),
//     P0 = sum_k T^k (R Q R^T) (T^k)^T ,   a0 = sum_k T^k c ,
// summed by doubling -- X <- X + A X A^T, s <- s + A s, A <- A A -- so 2^j terms cost j steps of three small products;
// the same loop with T^T gives the adjoint series.  One warp per series, matrices in the warp's shared-memory arena.
// The reference's direct (Kronecker) / bilinear solvers compute the same unique solution; a series whose powers have not
// died out after 60 doublings (spectral radius >= 1: not stationary) is flagged BKF_STATUS_NOT_STATIONARY.
#include "bkf_common.cuh"
#include "bkf_launch.h"

namespace bkf {
namespace stat {

// C[M,N] (row-major) = / += sum_k A(i,k) B(k,j),  A(i,k) = A[i*as0 + k*as1],  B(k,j) = B[k*bs0 + j*bs1]
template <bool ADD, typename R>
__device__ __forceinline__ void mm(R* __restrict__ C, const R* __restrict__ A, int as0, int as1,
                                   const R* __restrict__ Bm, int bs0, int bs1, int M, int N, int K, int lane) {
    for (int idx = lane; idx < M * N; idx += 32) {
        const int i = idx / N, j = idx - i * N;
        R acc = 0;
        for (int k = 0; k < K; ++k) acc = fma(A[i * as0 + k * as1], Bm[k * bs0 + j * bs1], acc);
        C[idx] = ADD ? C[idx] + acc : acc;
    }
    __syncwarp();
}

// X <- sum_k A^k X (A^k)^T and s <- sum_k A^k s by doubling; A is destroyed.  W, A2: m*m scratch; t: m scratch.
// Returns false when the powers of A have not decayed after 60 doublings.
template <typename R>
__device__ bool doubling(R* A, R* X, R* s, R* W, R* A2, R* t, int m, int lane) {
    const R tiny = sizeof(R) == 8 ? R(1e-40) : R(1e-20);
    for (int it = 0; it < 60; ++it) {
        R nrm = 0;
        for (int idx = lane; idx < m * m; idx += 32) nrm = fma(A[idx], A[idx], nrm);
        nrm = warp_sum(nrm);
        if (!(nrm > tiny)) return isfinite(nrm);  // A_j ~ 0: every remaining term is below rounding
        mm<false>(W, A, m, 1, X, m, 1, m, m, m, lane);   // W = A X
        mm<true>(X, W, m, 1, A, 1, m, m, m, m, lane);    // X += W A^T
        for (int i = lane; i < m; i += 32) {              // s += A s
            R acc = 0;
            for (int k = 0; k < m; ++k) acc = fma(A[i * m + k], s[k], acc);
            t[i] = acc;
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) s[i] += t[i];
        mm<false>(A2, A, m, 1, A, m, 1, m, m, m, lane);  // A <- A A
        for (int idx = lane; idx < m * m; idx += 32) A[idx] = A2[idx];
        __syncwarp();
    }
    return false;
}

__host__ __device__ inline size_t arena_elems(int m) { return 4 * (size_t)m * m + 2 * (size_t)m; }

template <typename R>
__global__ void forward_kernel(KArgs<R> g, int wpc) {
    extern __shared__ double smem_raw[];
    R* s_ = reinterpret_cast<R*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * wpc + warp;
    if (b >= g.B) return;
    const int m = g.m, r = g.r;
    R* A = s_ + (size_t)warp * arena_elems(m);
    R *X = A + m * m, *W = X + m * m, *A2 = W + m * m, *sv = A2 + m * m, *t = sv + m;
    const R* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
    const R* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)m * r);
    const R* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    const R* cg = param_ptr(g.c, g.shared_mask, BKF_PARAM_C, b, (size_t)m);
    for (int idx = lane; idx < m * m; idx += 32) A[idx] = Tg[idx];
    for (int i = lane; i < m; i += 32) sv[i] = cg[i];
    // X = R Q R^T  (pt.linalg.matrix_dot(R, Q, R.T): no symmetrisation)
    for (int idx = lane; idx < m * r; idx += 32) {
        const int i = idx / r, k = idx - i * r;
        R acc = 0;
        for (int l = 0; l < r; ++l) acc = fma(Rg[i * r + l], Qg[l * r + k], acc);
        W[idx] = acc;  // (R Q)[i,k]
    }
    __syncwarp();
    for (int idx = lane; idx < m * m; idx += 32) {
        const int i = idx / m, j = idx - i * m;
        R acc = 0;
        for (int k = 0; k < r; ++k) acc = fma(W[i * r + k], Rg[j * r + k], acc);
        X[idx] = acc;
    }
    __syncwarp();
    const bool ok = doubling(A, X, sv, W, A2, t, m, lane);
    for (int idx = lane; idx < m * m; idx += 32) g.sm_P[(size_t)b * m * m + idx] = X[idx];
    for (int i = lane; i < m; i += 32) g.sm_a[(size_t)b * m + i] = sv[i];
    if (lane == 0 && g.status) g.status[b] = ok ? 0 : BKF_STATUS_NOT_STATIONARY;
}

// inputs: T, R, Q, the forward results a0 (sm_filt_a) and P0 (sm_filt_P), cotangents a0bar (a0) and P0bar (P0);
// outputs: g_c, g_T, g_R, g_Q
template <typename R>
__global__ void backward_kernel(KArgs<R> g, int wpc) {
    extern __shared__ double smem_raw[];
    R* s_ = reinterpret_cast<R*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * wpc + warp;
    if (b >= g.B) return;
    const int m = g.m, r = g.r;
    R* A = s_ + (size_t)warp * arena_elems(m);
    R *S = A + m * m, *W = S + m * m, *A2 = W + m * m, *cb = A2 + m * m, *t = cb + m;
    const R* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
    const R* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)m * r);
    const R* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    const R* a0 = g.sm_filt_a + (size_t)b * m;
    const R* P0 = g.sm_filt_P + (size_t)b * m * m;
    const R* abar = g.a0 ? g.a0 + (size_t)b * m : nullptr;
    const R* Pbar = g.P0 ? g.P0 + (size_t)b * m * m : nullptr;
    for (int idx = lane; idx < m * m; idx += 32) {
        A[idx] = Tg[(idx % m) * m + idx / m];  // A = T^T
        S[idx] = Pbar ? Pbar[idx] : R(0);
    }
    for (int i = lane; i < m; i += 32) cb[i] = abar ? abar[i] : R(0);
    __syncwarp();
    doubling(A, S, cb, W, A2, t, m, lane);  // S = sum (T^T)^k Pbar T^k ;  cb = (I - T)^{-T} a0bar
    if (g.g_c)
        for (int i = lane; i < m; i += 32) g.g_c[(size_t)b * m + i] = cb[i];
    // Tbar = S T P0^T + S^T T P0 + cb a0^T
    if (g.g_T) {
        mm<false>(W, S, m, 1, Tg, m, 1, m, m, m, lane);      // W = S T
        mm<false>(A2, W, m, 1, P0, 1, m, m, m, m, lane);     // A2 = (S T) P0^T
        mm<false>(W, S, 1, m, Tg, m, 1, m, m, m, lane);      // W = S^T T
        mm<true>(A2, W, m, 1, P0, m, 1, m, m, m, lane);      // A2 += (S^T T) P0
        for (int idx = lane; idx < m * m; idx += 32)
            g.g_T[(size_t)b * m * m + idx] = fma(cb[idx / m], a0[idx % m], A2[idx]);
        __syncwarp();
    }
    // Qtilde_bar = S:  Rbar = S R Q^T + S^T R Q ;  Qbar = R^T S R
    if (g.g_R || g.g_Q) {
        mm<false>(W, S, m, 1, Rg, r, 1, m, r, m, lane);      // W[:m*r] = S R
        mm<false>(A2, S, 1, m, Rg, r, 1, m, r, m, lane);     // A2[:m*r] = S^T R
        if (g.g_R)
            for (int idx = lane; idx < m * r; idx += 32) {
                const int i = idx / r, j = idx - i * r;
                R acc = 0;
                for (int k = 0; k < r; ++k) acc = fma(W[i * r + k], Qg[j * r + k], fma(A2[i * r + k], Qg[k * r + j], acc));
                g.g_R[(size_t)b * m * r + idx] = acc;
            }
        if (g.g_Q)
            for (int idx = lane; idx < r * r; idx += 32) {
                const int i = idx / r, j = idx - i * r;
                R acc = 0;
                for (int k = 0; k < m; ++k) acc = fma(Rg[k * r + i], W[k * r + j], acc);
                g.g_Q[(size_t)b * r * r + idx] = acc;
            }
    }
}

template <typename R>
int launch(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err) {
    const size_t arena_bytes = arena_elems(g.m) * sizeof(R);
    if (arena_bytes > 227 * 1024 || g.r > g.m) {
        *err = "problem too large for the stationary-initialisation kernels (shared memory) or r > m";
        return BKF_ERR_UNSUPPORTED;
    }
    int wpc = 4;
    while (wpc > 1 && arena_bytes * wpc > 96 * 1024) wpc >>= 1;
    const size_t smem = arena_bytes * wpc;
    auto kern = backward ? backward_kernel<R> : forward_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<(g.B + wpc - 1) / wpc, 32 * wpc, smem, stream>>>(g, wpc);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

template int launch<double>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<float>(const KArgs<float>&, bool, cudaStream_t, const char**);

}  // namespace stat
}  // namespace bkf
