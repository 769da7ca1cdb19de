// bkf_sqrt.cu -- SquareRootFilter (kalman_filter.py:620-750): forward and reverse-mode adjoint, one CTA per series.
//
// Forward (per step): update via the R factor of the (p+m)x(p+m) matrix A = [[Hc^T, 0], [(Zm Pc)^T, Pc^T]]
// (kalman_filter.py:675-683), predict via the R factor of M = [T Pcf, R Qc]^T (:644-646).  Householder QR follows
// LAPACK dgeqr2/dlarfg (beta = -sign(alpha) ||x||, H = I for a zero sub-column) so that the signs of the factors --
// which the reference's log-likelihood depends on for p > 1 -- match numpy.linalg.qr.
//
// Adjoint: for A = Q Rf and an upper-triangular cotangent Rbar,  Abar = A Psi,
//     Psi = Rf^{-1} copyltu(Rf Rbar^T) Rf^{-T},   copyltu(X) = tril(X) + tril(X,-1)^T
// (the Q-free form of the standard QR pull-back; it is what autograd through the reference's code yields).
// cholesky(H), cholesky(Q) are time-invariant: their factor cotangents are summed over t and pulled back once per
// series with the lower-triangular convention of PyTensor's Cholesky.L_op (tril(S + S^T) - diag(S)).
// Validated against torch autograd through the reference source (tests/golden, oracle/kalman_torch.py).
#include "bkf_common.cuh"
#include "bkf_launch.h"

namespace bkf {
namespace sq {

enum { SET = 0, ADD = 1, SUB = 2 };

struct Th {
    int tid, nt;
    bool fast;  // float64 and m, p, r multiples of 8: the O(n^3) primitives run on FP64 tensor-core tiles (DMMA)
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// DMMA-tiled  C (op)= A B  on shared-memory operands with arbitrary strides: A(i,k) = A[i*as0 + k*as1],
// B(k,j) = B[k*bs0 + j*bs1], C(i,j) = C[i*cs0 + j*cs1].  M, N multiples of 8, K a multiple of 4.  The warps of the
// CTA split the 8 x 8 output tiles; two accumulator chains per tile.  BMASK = 1 keeps only B(k,j) with k >= j
// (the upper-triangular cotangent view of qr_pullback_psi).  In-place use (C aliasing B) is safe when K == 8 and
// M == 8: a warp loads both k-slices of its own tile column before the warp-synchronous mma and stores after it.
template <int MODE, int BMASK>
__device__ void mm_dmma(double* C, int cs0, int cs1, const double* A, int as0, int as1, const double* Bm, int bs0,
                        int bs1, int M, int N, int K, Th th) {
    const int warp = th.tid >> 5, nw = th.nt >> 5, lane = th.tid & 31, g = lane >> 2, t = lane & 3;
    const int tn = N >> 3, tiles = (M >> 3) * tn;
    for (int tile = warp; tile < tiles; tile += nw) {
        const int ti = tile / tn, tj = tile - ti * tn;
        const double* a = A + (8 * ti + g) * as0 + t * as1;
        const double* b = Bm + t * bs0 + (8 * tj + g) * bs1;
        const int jcol = 8 * tj + g;
        double c0 = 0, c1 = 0, e0 = 0, e1 = 0;
        int k0 = 0;
        for (; k0 + 8 <= K; k0 += 8) {
            double b0 = b[k0 * bs0], b1 = b[(k0 + 4) * bs0];
            if (BMASK) {
                if (k0 + t < jcol) b0 = 0;
                if (k0 + 4 + t < jcol) b1 = 0;
            }
            dmma(c0, c1, a[k0 * as1], b0);
            dmma(e0, e1, a[(k0 + 4) * as1], b1);
        }
        if (k0 < K) {
            double b0 = b[k0 * bs0];
            if (BMASK && k0 + t < jcol) b0 = 0;
            dmma(c0, c1, a[k0 * as1], b0);
        }
        c0 += e0;
        c1 += e1;
        double* c = C + (8 * ti + g) * cs0 + (8 * tj + 2 * t) * cs1;
        if (MODE == SET) { c[0] = c0; c[cs1] = c1; }
        else if (MODE == ADD) { c[0] += c0; c[cs1] += c1; }
        else { c[0] -= c0; c[cs1] -= c1; }
    }
    __syncthreads();
}

template <int MODE, typename R>
__device__ __forceinline__ void mm(R* __restrict__ C, int ldc, const R* __restrict__ A, int as0, int as1,
                                   const R* __restrict__ Bm, int bs0, int bs1, int M, int N, int K, Th th) {
    if constexpr (sizeof(R) == 8) {
        if (th.fast && !(M & 7) && !(N & 7) && !(K & 3)) {
            mm_dmma<MODE, 0>(C, ldc, 1, A, as0, as1, Bm, bs0, bs1, M, N, K, th);
            return;
        }
    }
    for (int idx = th.tid; idx < M * N; idx += th.nt) {
        int i = idx / N, j = idx - i * N;
        const R* a = A + i * as0;
        const R* b = Bm + j * bs1;
        R acc = 0;
        for (int k = 0; k < K; ++k) acc = fma(a[k * as1], b[k * bs0], acc);
        R* c = C + i * ldc + j;
        if (MODE == SET) *c = acc;
        else if (MODE == ADD) *c += acc;
        else *c -= acc;
    }
    __syncthreads();
}

template <typename R>
__device__ __forceinline__ void bcopy(R* dst, const R* src, int n, Th th) {
    for (int i = th.tid; i < n; i += th.nt) dst[i] = src[i];
    __syncthreads();
}
template <typename R>
__device__ __forceinline__ void bzero(R* dst, int n, Th th) {
    for (int i = th.tid; i < n; i += th.nt) dst[i] = 0;
    __syncthreads();
}

template <typename R>
__device__ __forceinline__ R block_sum(R v, R* red, Th th) {
    v = warp_sum(v);
    if ((th.tid & 31) == 0) red[th.tid >> 5] = v;
    __syncthreads();
    R s = 0;
    for (int w = 0; w < (th.nt + 31) / 32; ++w) s += red[w];
    __syncthreads();
    return s;
}

// In-place R factor of W[rows, cols] (row-major, ld = cols); entries below the diagonal are zeroed.
template <typename R>
__device__ void householder_r(R* W, int rows, int cols, R* red, Th th) {
    const int kmax = rows < cols ? rows : cols;
    for (int j = 0; j < kmax; ++j) {
        R ss = 0;
        for (int i = j + 1 + th.tid; i < rows; i += th.nt) ss = fma(W[i * cols + j], W[i * cols + j], ss);
        ss = block_sum(ss, red, th);
        const R alpha = W[j * cols + j];
        __syncthreads();
        if (ss == R(0)) continue;  // dlarfg: tau = 0, H = I
        const R nrm = sqrt(fma(alpha, alpha, ss));
        const R beta = alpha >= R(0) ? -nrm : nrm;
        const R tau = (beta - alpha) / beta;
        const R scal = R(1) / (alpha - beta);
        for (int i = j + 1 + th.tid; i < rows; i += th.nt) W[i * cols + j] *= scal;
        __syncthreads();
        for (int c = j + 1 + th.tid; c < cols; c += th.nt) {
            R s = W[j * cols + c];
            for (int i = j + 1; i < rows; ++i) s = fma(W[i * cols + j], W[i * cols + c], s);
            s *= tau;
            W[j * cols + c] -= s;
            for (int i = j + 1; i < rows; ++i) W[i * cols + c] = fma(-s, W[i * cols + j], W[i * cols + c]);
        }
        __syncthreads();
        if (th.tid == 0) W[j * cols + j] = beta;
        for (int i = j + 1 + th.tid; i < rows; i += th.nt) W[i * cols + j] = 0;
        __syncthreads();
    }
}

// Sums of eight per-lane values over the warp, all eight totals returned in every lane.  Instead of eight 5-stage
// butterflies (40 double shuffles) the first three stages exchange HALF of the values each (4 + 2 + 1 shuffles), two more
// finish the one value a lane is left with, and eight broadcasts hand the totals out: 17 double shuffles, 9 adds.
__device__ __forceinline__ void warp_sum8(double (&v)[8], int lane) {
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
    double w[4], x[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = h4 ? v[j] : v[j + 4], keep = h4 ? v[j + 4] : v[j];
        w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double send = h3 ? w[j] : w[j + 2], keep = h3 ? w[j + 2] : w[j];
        x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    double y = (h2 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, h2 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    // lane holds the total of value 4*bit4 + 2*bit3 + bit2 = (lane >> 2) & 7
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __shfl_sync(0xffffffffu, y, 4 * j);
}

// Blocked Householder R factor (panel width 8, compact-WY trailing updates on DMMA tiles); rows, cols multiples of 8,
// rows <= 64.  Same reflectors as householder_r (LAPACK dgeqr2/dlarfg sign rules), applied block-wise as dgeqrf does:
//   panel:    one warp, lane-per-row (rows l and l + 32), the 8 panel columns in registers, warp reductions
//   T factor: dlarft (forward, columnwise):  T[i][i] = tau_i,  T[0:i, i] = -tau_i T[0:i,0:i] V[:,0:i]^T v_i
//   trailing: A <- A - V T^T (V^T A)   (three mm_dmma calls)
// hv: rows*8 (V with unit diagonal), ht: 64 (T), hw: 8*cols (V^T A).
__device__ void householder_r_fast(double* W, int rows, int cols, double* hv, double* ht, double* hw, Th th) {
    const int lane = th.tid & 31, warp = th.tid >> 5;
    const int kmaxp = (rows < cols ? rows : cols) >> 3;
    for (int jp = 0; jp < kmaxp; ++jp) {
        const int c0 = jp * 8, prow = rows - c0;
        if (warp == 0) {
            // local rows: slot 0 -> lane, slot 1 -> lane + 32 (relative to c0)
            double a[2][8];
            const bool has1 = lane + 32 < prow;
            const bool has0 = lane < prow;
            {
                const double2* r0 = reinterpret_cast<const double2*>(W + (c0 + lane) * cols + c0);
                const double2* r1 = reinterpret_cast<const double2*>(W + (c0 + lane + 32) * cols + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 u0 = has0 ? r0[q] : make_double2(0.0, 0.0);
                    const double2 u1 = has1 ? r1[q] : make_double2(0.0, 0.0);
                    a[0][2 * q] = u0.x; a[0][2 * q + 1] = u0.y;
                    a[1][2 * q] = u1.x; a[1][2 * q + 1] = u1.y;
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // x = a[k+1:, k] (unscaled).  ONE window of independent warp reductions per reflector:
                //   ss = x.x,  d[c] = x . a[k+1:, c] (c > k),  e[i] = x . v_i[k+1:] (i < k, for the T factor)
                const double x0 = lane > k ? a[0][k] : 0.0, x1 = a[1][k];
                double d[8];  // c > k: trailing panel column;  c == k: ss;  c < k: column c holds v_c below its diagonal
#pragma unroll
                for (int c = 0; c < 8; ++c) d[c] = fma(x1, a[1][c], x0 * a[0][c]);
                warp_sum8(d, lane);
                const double ss = d[k];
                const double(&e)[8] = d;
                const double alpha = __shfl_sync(0xffffffffu, a[0][k], k);
                if (ss == 0.0) {  // dlarfg: tau = 0, H = I
                    if (lane > k) a[0][k] = 0.0;
                    a[1][k] = 0.0;
                    if (lane < 8) ht[lane * 8 + k] = 0.0;
                    __syncwarp();
                    continue;
                }
                const double nrm = sqrt(fma(alpha, alpha, ss));
                const double beta = alpha >= 0.0 ? -nrm : nrm;
                const double tau = (beta - alpha) / beta;
                const double scal = 1.0 / (alpha - beta);
                // v (unit at local row k); rows above the pivot do not take part
                const double v0 = lane > k ? x0 * scal : (lane == k ? 1.0 : 0.0);
                const double v1 = x1 * scal;
#pragma unroll
                for (int c = k + 1; c < 8; ++c) {
                    const double akc = __shfl_sync(0xffffffffu, a[0][c], k);
                    const double w = tau * fma(scal, d[c], akc);  // tau v^T a_c
                    a[0][c] = fma(-w, v0, a[0][c]);
                    a[1][c] = fma(-w, v1, a[1][c]);
                }
                // dlarft column k:  z_i = v_i^T v_k = v_i[k] + scal e_i ;  T[0:k,k] = -tau T[0:k,0:k] z ;  T[k][k] = tau
                double z[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i] = (i < k) ? fma(scal, e[i], __shfl_sync(0xffffffffu, a[0][i], k)) : 0.0;
                if (lane < 8) {  // lane i owns row i of T (kept in shared memory: ht)
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < 8; ++l)
                        if (l < k) acc = fma((l >= lane) ? ht[lane * 8 + l] : 0.0, z[l], acc);
                    ht[lane * 8 + k] = lane < k ? -tau * acc : (lane == k ? tau : 0.0);
                }
                __syncwarp();
                if (lane == k) a[0][k] = beta;
                else if (lane > k) a[0][k] = v0;
                a[1][k] = v1;
            }
            // write the panel back: R on and above the diagonal, zeros below; the reflectors go to hv (unit diagonal)
            {
                double2* r0 = reinterpret_cast<double2*>(W + (c0 + lane) * cols + c0);
                double2* r1 = reinterpret_cast<double2*>(W + (c0 + lane + 32) * cols + c0);
                double2* h0 = reinterpret_cast<double2*>(hv + lane * 8);
                double2* h1 = reinterpret_cast<double2*>(hv + (lane + 32) * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k0 = 2 * q, k1 = 2 * q + 1;
                    if (has0) {
                        r0[q] = make_double2(lane <= k0 ? a[0][k0] : 0.0, lane <= k1 ? a[0][k1] : 0.0);
                        h0[q] = make_double2(lane > k0 ? a[0][k0] : (lane == k0 ? 1.0 : 0.0),
                                             lane > k1 ? a[0][k1] : (lane == k1 ? 1.0 : 0.0));
                    }
                    if (has1) {
                        r1[q] = make_double2(0.0, 0.0);
                        h1[q] = make_double2(a[1][k0], a[1][k1]);
                    }
                }
            }
        }
        __syncthreads();
        const int nt_cols = cols - c0 - 8;
        if (nt_cols > 0) {
            double* At = W + c0 * cols + c0 + 8;
            mm_dmma<SET, 0>(hw, nt_cols, 1, hv, 1, 8, At, cols, 1, 8, nt_cols, prow, th);      // V^T A
            mm_dmma<SET, 0>(hw, nt_cols, 1, ht, 1, 8, hw, nt_cols, 1, 8, nt_cols, 8, th);     // T^T (V^T A), in place
            mm_dmma<SUB, 0>(At, cols, 1, hv, 8, 1, hw, nt_cols, 1, prow, nt_cols, 8, th);     // A -= V (...)
        }
    }
}

template <typename R>
__device__ __forceinline__ void qr_r(R* W, int rows, int cols, R* red, R* hv, R* ht, R* hw, Th th) {
    if constexpr (sizeof(R) == 8) {
        if (th.fast && !(rows & 7) && !(cols & 7) && rows <= 64) {
            householder_r_fast(W, rows, cols, hv, ht, hw, th);
            return;
        }
    }
    householder_r(W, rows, cols, red, th);
}

// X <- U^{-1} X on DMMA tiles (n, ncols multiples of 8): the 8 x 8 diagonal tiles are inverted once (one warp each,
// back-substitution by 8 lanes), then block back-substitution  X_I <- Uinv_I (X_I - sum_{J>I} U_IJ X_J).
// uinv: (n/8)*64 doubles of scratch.
__device__ void solve_upper_fast(const double* U, int ldu, int n, double* X, int xs0, int xs1, int ncols, double* uinv,
                                 Th th) {
    const int nb = n >> 3, lane = th.tid & 31, warp = th.tid >> 5, nw = th.nt >> 5;
    for (int I = warp; I < nb; I += nw) {
        if (lane < 8) {
            const double* D = U + (8 * I) * ldu + 8 * I;
            double x[8];
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                double tt = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 7; q > i; --q) tt = fma(-D[i * ldu + q], x[q], tt);
                x[i] = tt / D[i * ldu + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) uinv[I * 64 + i * 8 + lane] = x[i];
        }
    }
    __syncthreads();
    for (int I = nb - 1; I >= 0; --I) {
        double* XI = X + (8 * I) * xs0;
        if (I < nb - 1)
            mm_dmma<SUB, 0>(XI, xs0, xs1, U + (8 * I) * ldu + 8 * (I + 1), ldu, 1, X + 8 * (I + 1) * xs0, xs0, xs1, 8, ncols,
                            8 * (nb - 1 - I), th);
        mm_dmma<SET, 0>(XI, xs0, xs1, uinv + I * 64, 8, 1, XI, xs0, xs1, 8, ncols, 8, th);
    }
}

// X <- U^{-1} X for upper-triangular U[n,n] (ld = ldu); X element (i, c) at X[i*xs0 + c*xs1], ncols columns.
template <typename R>
__device__ void solve_upper(const R* U, int ldu, int n, R* X, int xs0, int xs1, int ncols, Th th) {
    for (int c = th.tid; c < ncols; c += th.nt) {
        for (int i = n - 1; i >= 0; --i) {
            R t = X[i * xs0 + c * xs1];
            for (int q = i + 1; q < n; ++q) t = fma(-U[i * ldu + q], X[q * xs0 + c * xs1], t);
            X[i * xs0 + c * xs1] = t / U[i * ldu + i];
        }
    }
    __syncthreads();
}

// Psi = U^{-1} copyltu(U Rbar^T) U^{-T}  (n x n, symmetric), with Rbar(i,j) = Rb[i*rs0 + j*rs1] for i <= j, 0 below.
// tmp: n*n scratch.
template <typename R>
__device__ void qr_pullback_psi(R* Psi, const R* U, int ldu, int n, const R* Rb, int rs0, int rs1, R* tmp, R* uinv,
                                Th th) {
    if constexpr (sizeof(R) == 8) {
        if (th.fast && !(n & 7)) {
            // U is stored with exact zeros below its diagonal, Rbar is masked to k >= j inside the product
            mm_dmma<SET, 1>(tmp, n, 1, U, ldu, 1, Rb, rs1, rs0, n, n, n, th);
            for (int idx = th.tid; idx < n * n; idx += th.nt) {
                int i = idx / n, j = idx - i * n;
                Psi[idx] = (i >= j) ? tmp[idx] : tmp[j * n + i];
            }
            __syncthreads();
            solve_upper_fast(U, ldu, n, Psi, n, 1, n, uinv, th);
            solve_upper_fast(U, ldu, n, Psi, 1, n, n, uinv, th);
            return;
        }
    }
    // tmp = U Rbar^T : tmp[i,j] = sum_{k >= max(i,j)} U[i,k] Rbar[j,k]
    for (int idx = th.tid; idx < n * n; idx += th.nt) {
        int i = idx / n, j = idx - i * n;
        R acc = 0;
        for (int k = (i > j ? i : j); k < n; ++k) acc = fma(U[i * ldu + k], Rb[j * rs0 + k * rs1], acc);
        tmp[idx] = acc;
    }
    __syncthreads();
    for (int idx = th.tid; idx < n * n; idx += th.nt) {
        int i = idx / n, j = idx - i * n;
        Psi[idx] = (i >= j) ? tmp[idx] : tmp[j * n + i];
    }
    __syncthreads();
    solve_upper(U, ldu, n, Psi, n, 1, n, th);  // X = U^{-1} Sym
    solve_upper(U, ldu, n, Psi, 1, n, n, th);  // Psi = U^{-1} X^T  (in place on the transposed view; symmetric)
}

// In-place lower Cholesky by thread 0 (tiny p x p / r x r, once per series).
template <typename R>
__device__ bool chol_lower(R* A, int k, Th th) {
    __shared__ int ok_s;
    if (th.tid == 0) {
        int ok = 1;
        for (int j = 0; j < k; ++j) {
            R s = A[j * k + j];
            for (int q = 0; q < j; ++q) s -= A[j * k + q] * A[j * k + q];
            if (!(s > 0)) { ok = 0; s = R(NAN); }
            R ljj = sqrt(s);
            A[j * k + j] = ljj;
            for (int i = j + 1; i < k; ++i) {
                R t = A[i * k + j];
                for (int q = 0; q < j; ++q) t -= A[i * k + q] * A[j * k + q];
                A[i * k + j] = t / ljj;
            }
            for (int i = 0; i < j; ++i) A[i * k + j] = 0;
        }
        ok_s = ok;
    }
    __syncthreads();
    return ok_s != 0;
}

// Pull a lower-triangular factor cotangent Lbar back through S = L L^T, PyTensor lower convention, by thread 0.
// Sbar = tril(2 X) - diag(X),  X = 0.5 L^{-T} (P + P^T) L^{-1},  P = Phi(L^T Lbar).  out may alias nothing; tmp: 2*k*k.
template <typename R>
__device__ void chol_pullback(R* out, const R* L, const R* Lbar, int k, R* tmp, Th th) {
    if (th.tid == 0) {
        R* P = tmp;
        R* X = tmp + k * k;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                R acc = 0;  // (L^T Lbar)[i,j] = sum_q L[q,i] Lbar[q,j], Lbar lower
                for (int q = (i > j ? i : j); q < k; ++q) acc = fma(L[q * k + i], Lbar[q * k + j], acc);
                P[i * k + j] = (i > j) ? acc : (i == j ? R(0.5) * acc : R(0));
            }
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) X[i * k + j] = P[i * k + j] + P[j * k + i];
        // X <- L^{-T} X : solve L^T Y = X column by column (upper-triangular system)
        for (int c = 0; c < k; ++c)
            for (int i = k - 1; i >= 0; --i) {
                R t = X[i * k + c];
                for (int q = i + 1; q < k; ++q) t -= L[q * k + i] * X[q * k + c];
                X[i * k + c] = t / L[i * k + i];
            }
        // X <- X L^{-1} : solve Y L = X row by row
        for (int rw = 0; rw < k; ++rw)
            for (int j = k - 1; j >= 0; --j) {
                R t = X[rw * k + j];
                for (int q = j + 1; q < k; ++q) t -= X[rw * k + q] * L[q * k + j];
                X[rw * k + j] = t / L[j * k + j];
            }
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                const R s = R(0.5) * X[i * k + j];
                out[i * k + j] = (i > j) ? R(2) * s : (i == j ? s : R(0));
            }
    }
    __syncthreads();
}

// Doubles per (series, step) record of the forward sweep kept for the adjoint:
// [a (m) | Pc (m x m) | upper triangle of Rf ((p+m) x (p+m)) | upper triangle of Rp (m x m)]
__host__ __device__ inline size_t traj_record_elems(int m, int p) {
    const size_t N1 = (size_t)p + m;
    return (size_t)m + (size_t)m * m + N1 * (N1 + 1) / 2 + (size_t)m * (m + 1) / 2;
}

// Upper triangle of the leading n x n block of W (row-major, ld) <-> packed rows (row i holds n - i entries).
template <typename R>
__device__ void store_upper(R* __restrict__ dst, const R* __restrict__ W, int ld, int n, Th th) {
    for (int idx = th.tid; idx < n * n; idx += th.nt) {
        int i = idx / n, j = idx - i * n;
        if (j >= i) dst[i * n - (i * (i - 1)) / 2 + (j - i)] = W[i * ld + j];
    }
}
template <typename R>
__device__ void load_upper(R* __restrict__ W, int ld, int rows, int n, const R* __restrict__ src, Th th) {
    for (int idx = th.tid; idx < rows * n; idx += th.nt) {
        int i = idx / n, j = idx - i * n;
        W[i * ld + j] = (i < n && j >= i) ? src[i * n - (i * (i - 1)) / 2 + (j - i)] : R(0);
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------------------
template <typename R>
struct Arena {
    R *T, *Z, *Hc, *Qc, *RQc, *c, *d;        // parameters (factors hoisted: H, Q are time-invariant)
    R *a, *Pc;                               // carried prediction (factor)
    R *Zm, *w, *ym, *v, *sv, *inner, *af, *Pcf;
    R *A0, *Rf, *M0, *Rp;                    // update / predict QR inputs and factors
    R *red;                                  // block reduction scratch (32)
    R *hv, *ht, *hw, *uinv;                  // blocked Householder / tile-solve scratch (fast path)
    // backward only
    R *abar, *Pbar, *S1, *Psi, *Abar, *gT, *gZ, *gc, *gd, *Psum, *Hcbar, *afbar, *Pcfbar, *U1, *vbar, *sbar, *t1;
};

// The forward sweep keeps only what it reads again: the update array is factored in place (A0 aliases Rf), so is the
// predict array (M0 aliases Rp), and Pcf overwrites Pc -- 72 KB at the C4 shape, three CTAs per SM.  The adjoint sweep
// needs the unfactored arrays next to the factors and its own accumulators (one CTA per SM at C4).
__host__ __device__ inline size_t arena_elems(int m, int p, int r, bool bwd) {
    size_t mm_ = (size_t)m * m, pm = (size_t)p * m, pp = (size_t)p * p, rr = (size_t)r * r, mr = (size_t)m * r;
    size_t n1 = (size_t)(p + m) * (p + m), n2 = (size_t)(m + r) * m, sc = 8 * (size_t)(p + m + r);
    size_t common = /*T,Pc*/ 2 * mm_ + /*Z,Zm*/ 2 * pm + pp + rr + mr + /*c,a,af*/ 3 * (size_t)m +
                    /*d,w,ym,v,sv,inner*/ 6 * (size_t)p + /*Rf*/ n1 + /*Rp*/ n2 + 32 + /*hv,hw*/ 2 * sc + /*ht*/ 64;
    size_t bw = /*Pcf*/ mm_ + /*A0*/ n1 + /*M0*/ n2 + /*Pbar,gT,Psum,Pcfbar,U1*/ 5 * mm_ + /*S1,Psi,Abar*/ 3 * n1 + pm +
                /*abar,gc,afbar*/ 3 * (size_t)m + /*gd,vbar,sbar,t1*/ 4 * (size_t)p + pp + /*uinv*/ sc;
    size_t tot = common + (bwd ? bw : 0);
    return (tot + 1) & ~(size_t)1;
}

template <typename R>
__device__ void carve(Arena<R>& A, R* s, int m, int p, int r, bool bwd) {
    size_t mm_ = (size_t)m * m, pm = (size_t)p * m, pp = (size_t)p * p, rr = (size_t)r * r, mr = (size_t)m * r;
    size_t n1 = (size_t)(p + m) * (p + m), n2 = (size_t)(m + r) * m, sc = 8 * (size_t)(p + m + r);
    auto take = [&](size_t n) { R* q = s; s += n; return q; };
    A.T = take(mm_); A.Pc = take(mm_); A.Z = take(pm); A.Zm = take(pm); A.Hc = take(pp);
    A.Qc = take(rr); A.RQc = take(mr); A.c = take(m); A.a = take(m); A.af = take(m);
    A.d = take(p); A.w = take(p); A.ym = take(p); A.v = take(p); A.sv = take(p); A.inner = take(p);
    A.Rf = take(n1); A.Rp = take(n2); A.red = take(32);
    A.hv = take(sc); A.hw = take(sc); A.ht = take(64);
    if (!bwd) {
        A.A0 = A.Rf; A.M0 = A.Rp; A.Pcf = A.Pc; A.uinv = nullptr;
    } else {
        A.Pcf = take(mm_); A.A0 = take(n1); A.M0 = take(n2);
        A.Pbar = take(mm_); A.gT = take(mm_); A.Psum = take(mm_); A.Pcfbar = take(mm_); A.U1 = take(mm_);
        A.S1 = take(n1); A.Psi = take(n1); A.Abar = take(n1); A.gZ = take(pm);
        A.abar = take(m); A.gc = take(m); A.afbar = take(m);
        A.gd = take(p); A.vbar = take(p); A.sbar = take(p); A.t1 = take(p); A.Hcbar = take(pp);
        A.uinv = take(sc);
    }
}

// Per-series parameters shared by forward and backward.  first: load everything (step t of the time-varying ones) and
// hoist the time-invariant factors; otherwise only what carries a time axis is refreshed for step t
// (filters/utilities.py:8-47).  Returns status bits;
// *Hzero_io = all(H_t == 0) (:666).
template <typename R>
__device__ int load_params(const KArgs<R>& g, Arena<R>& A, int b, int t, bool first, bool* Hzero_io, Th th) {
    const int m = g.m, p = g.p, r = g.r, n = g.n;
    auto want = [&](int which) { return first || ((g.tv_mask >> which) & 1u); };
    auto src = [&](const R* base, int which, size_t sz) {
        return param_ptr_t(base, g.shared_mask, g.tv_mask, which, b, t, n, sz);
    };
    if (want(BKF_PARAM_T)) bcopy(A.T, src(g.T, BKF_PARAM_T, (size_t)m * m), m * m, th);
    if (want(BKF_PARAM_Z)) bcopy(A.Z, src(g.Z, BKF_PARAM_Z, (size_t)p * m), p * m, th);
    if (want(BKF_PARAM_C)) bcopy(A.c, src(g.c, BKF_PARAM_C, (size_t)m), m, th);
    if (want(BKF_PARAM_D)) bcopy(A.d, src(g.d, BKF_PARAM_D, (size_t)p), p, th);
    int st = 0;
    if (want(BKF_PARAM_H)) {
        bcopy(A.Hc, src(g.H, BKF_PARAM_H, (size_t)p * p), p * p, th);
        bool Hzero = true;
        for (int i = 0; i < p * p; ++i) Hzero = Hzero && (A.Hc[i] == R(0));
        __syncthreads();
        if (!Hzero && !chol_lower(A.Hc, p, th)) st |= BKF_STATUS_NOT_PD;
        *Hzero_io = Hzero;
    }
    if (want(BKF_PARAM_Q)) {
        bcopy(A.Qc, src(g.Q, BKF_PARAM_Q, (size_t)r * r), r * r, th);
        if (!chol_lower(A.Qc, r, th)) st |= BKF_STATUS_NOT_PD;
    }
    if (want(BKF_PARAM_Q) || want(BKF_PARAM_R))
        mm<SET>(A.RQc, r, src(g.Rm, BKF_PARAM_R, (size_t)m * r), r, 1, A.Qc, r, 1, m, r, r, th);
    return st;
}

// One forward step from (A.a, A.Pc) and y_t.  Leaves A0 (QR input), Rf (its R factor), v, sv, inner, af, Pcf,
// M0, Rp in the arena and returns ll.  flag = all observations missing; partial = some (the reference's
// cholesky(W H) fails there: NaN is propagated).
template <typename R>
__device__ R step(const KArgs<R>& g, Arena<R>& A, const R* yt, bool Hzero, bool* flag_out, const R* Rf_stored,
                  Th th) {
    const int m = g.m, p = g.p, r = g.r, N1 = p + m;
    for (int o = th.tid; o < p; o += th.nt) {
        R yv = yt[o];
        bool miss = is_missing(yv, g.fill, true);
        A.w[o] = miss ? R(0) : R(1);
        A.ym[o] = miss ? R(0) : yv;
    }
    __syncthreads();
    R nobs = 0;
    for (int o = 0; o < p; ++o) nobs += A.w[o];
    const bool flag = (nobs == R(0));
    const bool partial = !flag && (nobs != R(p));
    for (int idx = th.tid; idx < p * m; idx += th.nt) A.Zm[idx] = A.w[idx / m] * A.Z[idx];
    __syncthreads();
    for (int o = th.tid; o < p; o += th.nt) {
        R acc = 0;
        for (int j = 0; j < m; ++j) acc = fma(A.Zm[o * m + j], A.a[j], acc);
        A.sv[o] = acc + A.d[o];  // y_hat (kept in sv until the solves overwrite it)
        A.v[o] = A.ym[o] - (acc + A.d[o]);
    }
    for (int idx = th.tid; idx < N1 * N1; idx += th.nt) {
        int i = idx / N1, j = idx - i * N1;
        R val;
        if (i < p) {
            val = (j < p) ? ((flag || Hzero) ? R(0) : (partial ? R(NAN) : A.Hc[j * p + i])) : R(0);
        } else if (j < p) {
            if (th.fast) continue;  // (Zm Pc)^T block: DMMA product below
            R acc = 0;
            for (int q = 0; q < m; ++q) acc = fma(A.Zm[j * m + q], A.Pc[q * m + (i - p)], acc);
            val = acc;
        } else {
            val = A.Pc[(j - p) * m + (i - p)];
        }
        A.A0[idx] = val;
        A.Rf[idx] = val;
    }
    __syncthreads();
    if constexpr (sizeof(R) == 8) {
        if (th.fast) {  // Rf[p+i, j] = sum_q Pc[q, i] Zm[j, q]
            mm_dmma<SET, 0>(A.Rf + p * N1, N1, 1, A.Pc, 1, m, A.Zm, 1, m, m, p, m, th);
            for (int idx = th.tid; idx < m * p; idx += th.nt) {
                int i = idx / p, j = idx - i * p;
                A.A0[(p + i) * N1 + j] = A.Rf[(p + i) * N1 + j];
            }
            __syncthreads();
        }
    }
    if (Rf_stored) load_upper(A.Rf, N1, N1, N1, Rf_stored, th);  // the adjoint sweep reuses the forward factor
    else qr_r(A.Rf, N1, N1, A.red, A.hv, A.ht, A.hw, th);
    *flag_out = flag;
    return R(0);
}

// Small triangular solves of the update (p <= 32 right-hand-side entries): lane q keeps x[q] in a register, the pivot
// value travels by shuffle and the reciprocal diagonal is formed once, so an elimination step costs one shuffle and one
// FMA instead of a division and a shared-memory round trip.  U = Rf[:p,:p] (upper, leading dimension ld); Fc = U^T.
//   LOWER:  x <- Fc^{-1} x   (forward substitution,  Fc[q,i] = U[i,q])
//   UPPER:  x <- U^{-1} x    (back substitution)
template <bool LOWER, typename R>
__device__ __forceinline__ R warp_trisolve(const R* __restrict__ U, int ld, int p, R x, int lane) {
    const R rdiag = lane < p ? R(1) / U[lane * ld + lane] : R(0);
    if (LOWER) {
        for (int i = 0; i < p; ++i) {
            const R xi = __shfl_sync(0xffffffffu, x, i) * __shfl_sync(0xffffffffu, rdiag, i);
            if (lane == i) x = xi;
            else if (lane > i && lane < p) x = fma(-U[i * ld + lane], xi, x);
        }
    } else {
        for (int i = p - 1; i >= 0; --i) {
            const R xi = __shfl_sync(0xffffffffu, x, i) * __shfl_sync(0xffffffffu, rdiag, i);
            if (lane == i) x = xi;
            else if (lane < i) x = fma(-U[lane * ld + i], xi, x);
        }
    }
    return x;
}

template <typename R>
__device__ R finish_update(const KArgs<R>& g, Arena<R>& A, bool flag, bool need_ll, Th th) {
    const int m = g.m, p = g.p, N1 = p + m;
    R ll = 0;
    if (!flag) {
        if (th.tid < 32) {  // s = Fc^{-1} v ; inner = Fc^{-1} s   with Fc[i,q] = Rf[q,i]   (:686-691)
            // warp 0, column-oriented forward substitution: x_i /= Fc[i,i];  x_q -= Fc[q,i] x_i (q > i) across lanes
            const int lane = th.tid;
            auto solve_fc = [&](R* x) {
                for (int i = 0; i < p; ++i) {
                    __syncwarp();
                    const R xi = x[i] / A.Rf[i * N1 + i];
                    __syncwarp();
                    if (lane == 0) x[i] = xi;
                    for (int q = i + 1 + lane; q < p; q += 32) x[q] = fma(-A.Rf[i * N1 + q], xi, x[q]);
                }
                __syncwarp();
            };
            if (p <= 32) {
                R x = lane < p ? A.v[lane] : R(0);
                x = warp_trisolve<true>(A.Rf, N1, p, x, lane);
                if (lane < p) A.sv[lane] = x;
                x = warp_trisolve<true>(A.Rf, N1, p, x, lane);
                if (lane < p) A.inner[lane] = x;
            } else {
                for (int i = lane; i < p; i += 32) A.sv[i] = A.v[i];
                solve_fc(A.sv);
                for (int i = lane; i < p; i += 32) A.inner[i] = A.sv[i];
                solve_fc(A.inner);
            }
        }
        __syncthreads();
        R quad = 0, logdet = 0;
        if (need_ll) {
            for (int o = 0; o < p; ++o) quad = fma(A.v[o], A.inner[o], quad);
            for (int o = th.tid; o < p; o += th.nt) logdet += log(fabs(A.Rf[o * N1 + o]));
            logdet = block_sum(logdet, A.red, th);
        }
        ll = R(-0.5) * (R(p) * (R(kLog2Pi) + R(2) * logdet) + quad);  // :696 (quirk Q2)
        for (int i = th.tid; i < m; i += th.nt) {
            R acc = 0;
            for (int o = 0; o < p; ++o) acc = fma(A.Rf[o * N1 + (p + i)], A.sv[o], acc);
            A.af[i] = A.a[i] + acc;
        }
        for (int idx = th.tid; idx < m * m; idx += th.nt) {
            int i = idx / m, j = idx - i * m;
            A.Pcf[idx] = A.Rf[(p + j) * N1 + (p + i)] + (i == j ? g.jitter : R(0));  // jitter on the factor (Q4)
        }
    } else {
        for (int i = th.tid; i < m; i += th.nt) A.af[i] = A.a[i];
        for (int idx = th.tid; idx < m * m; idx += th.nt) A.Pcf[idx] = A.Pc[idx] + ((idx / m) == (idx % m) ? g.jitter : R(0));
    }
    __syncthreads();
    return ll;
}

// predict: M0 = [T Pcf, R Qc]^T ((m+r) x m), Rp = its R factor; next (a, Pc) written to (a_next, Pc_next).
template <typename R>
__device__ void predict(const KArgs<R>& g, Arena<R>& A, R* a_next, R* Pc_next, const R* Rp_stored, Th th) {
    const int m = g.m, r = g.r;
    bool tp_done = false;
    if constexpr (sizeof(R) == 8) {
        if (th.fast) {  // M0[i, j] = sum_q Pcf[q, i] T[j, q]   (i, j < m)
            mm_dmma<SET, 0>(A.M0, m, 1, A.Pcf, 1, m, A.T, 1, m, m, m, m, th);
            tp_done = true;
        }
    }
    for (int idx = th.tid; idx < (m + r) * m; idx += th.nt) {
        int i = idx / m, j = idx - i * m;
        R val;
        if (i < m) {
            if (tp_done) {
                A.Rp[idx] = A.M0[idx];
                continue;
            }
            R acc = 0;
            for (int q = 0; q < m; ++q) acc = fma(A.T[j * m + q], A.Pcf[q * m + i], acc);
            val = acc;
        } else {
            val = A.RQc[j * r + (i - m)];
        }
        A.M0[idx] = val;
        A.Rp[idx] = val;
    }
    __syncthreads();
    if (Rp_stored) load_upper(A.Rp, m, m + r, m, Rp_stored, th);
    else qr_r(A.Rp, m + r, m, A.red, A.hv, A.ht, A.hw, th);
    if (a_next) {
        for (int i = th.tid; i < m; i += th.nt) {
            R acc = 0;
            for (int q = 0; q < m; ++q) acc = fma(A.T[i * m + q], A.af[q], acc);
            a_next[i] = acc + A.c[i];
        }
    }
    if (Pc_next)
        for (int idx = th.tid; idx < m * m; idx += th.nt) Pc_next[idx] = A.Rp[(idx % m) * m + idx / m];
    __syncthreads();
}

// NT = 128: three CTAs per SM at the C4 shape (72 KB each); NT = 256: two CTAs with twice the warps -- the launcher
// picks whichever finishes the batch in less (waves x time per wave).
template <typename R, int NT>
__global__ void __launch_bounds__(NT, NT == 128 ? 3 : 2) forward_kernel(KArgs<R> g) {
    extern __shared__ double smem_raw[];
    R* smem = reinterpret_cast<R*>(smem_raw);
    const Th th{(int)threadIdx.x, (int)blockDim.x, sizeof(R) == 8 && !((g.m | g.p | g.r) & 7)};
    const int b = blockIdx.x;
    const int m = g.m, p = g.p, n = g.n, N1 = p + m;
    Arena<R> A;
    carve(A, smem, m, p, g.r, false);
    bool Hzero;
    int st = load_params(g, A, b, 0, true, &Hzero, th);
    bcopy(A.a, param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)m), m, th);
    bcopy(A.Pc, param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)m * m), m * m, th);
    const R* yb = g.y_shared ? g.y : g.y + (size_t)b * n * p;
    const size_t tstride = traj_record_elems(m, p);
    const size_t off_rf = (size_t)m + (size_t)m * m, off_rp = off_rf + (size_t)N1 * (N1 + 1) / 2;
    auto write_sq = [&](R* dst, const R* L, int k, int ls0, int ls1) {  // dst = L L^T with L(i,q) = L[i*ls0 + q*ls1]
        for (int idx = th.tid; idx < k * k; idx += th.nt) {
            int i = idx / k, j = idx - i * k;
            R acc = 0;
            for (int q = 0; q < k; ++q) acc = fma(L[i * ls0 + q * ls1], L[j * ls0 + q * ls1], acc);
            dst[idx] = acc;
        }
    };
    R logp = 0;
    for (int t = 0; t < n; ++t) {
        const size_t bt = (size_t)b * n + t;
        if (g.tv_mask && t > 0) st |= load_params(g, A, b, t, false, &Hzero, th);
        if (g.traj) {
            bcopy(g.traj + bt * tstride, A.a, m, th);
            bcopy(g.traj + bt * tstride + m, A.Pc, m * m, th);
        }
        if (g.pred_a) bcopy(g.pred_a + bt * m, A.a, m, th);
        if (g.pred_P) write_sq(g.pred_P + bt * m * m, A.Pc, m, m, 1);
        bool flag;
        step(g, A, yb + (size_t)t * p, Hzero, &flag, (const R*)nullptr, th);
        if (g.traj) store_upper(g.traj + bt * tstride + off_rf, A.Rf, N1, N1, th);
        if (g.obs_mu) bcopy(g.obs_mu + bt * p, A.sv, p, th);
        if (g.obs_cov) write_sq(g.obs_cov + bt * p * p, A.Rf, p, 1, N1);  // F_chol[i,q] = Rf[q,i]
        __syncthreads();
        const R ll = finish_update(g, A, flag, true, th);
        if (g.filt_a) bcopy(g.filt_a + bt * m, A.af, m, th);
        if (g.filt_P) write_sq(g.filt_P + bt * m * m, A.Pcf, m, m, 1);
        if (g.ll && th.tid == 0) g.ll[bt] = ll;
        logp += ll;
        if (g.traj) {  // Rp is complete once predict() has run; A.Pc (= Rp^T) is only read by the next step
            predict(g, A, A.a, A.Pc, (const R*)nullptr, th);
            store_upper(g.traj + bt * tstride + off_rp, A.Rp, m, m, th);
            __syncthreads();
        } else {
            predict(g, A, A.a, A.Pc, (const R*)nullptr, th);
        }
    }
    if (!isfinite(logp)) st |= BKF_STATUS_NONFINITE;
    if (th.tid == 0) {
        if (g.logp) g.logp[b] = logp;
        if (g.status) g.status[b] = st;
    }
}

// Cotangents of R and of the factor Qc from P = sum of Psi_p (or one step's Psi_p):
//   X1 = P R,  gR = X1 (Qc Qc^T) -> A.Pcfbar[:m*r],  Qcbar = tril(R^T X1 Qc) -> A.S1[:r*r].   Scratch: U1, S1, Abar.
template <typename R>
__device__ void rq_cotangents(Arena<R>& A, const R* P, const R* Rg, int m, int r, Th th) {
    mm<SET>(A.U1, r, P, m, 1, Rg, r, 1, m, r, m, th);            // X1
    mm<SET>(A.S1, r, A.Qc, r, 1, A.Qc, 1, r, r, r, r, th);        // Qc Qc^T
    mm<SET>(A.Pcfbar, r, A.U1, r, 1, A.S1, r, 1, m, r, r, th);    // gR
    mm<SET>(A.Abar, r, A.U1, r, 1, A.Qc, r, 1, m, r, r, th);      // X1 Qc (m x r)
    mm<SET>(A.S1, r, Rg, 1, r, A.Abar, r, 1, r, r, m, th);        // R^T X1 Qc (r x r)
    for (int idx = th.tid; idx < r * r; idx += th.nt)
        if ((idx % r) > (idx / r)) A.S1[idx] = 0;
    __syncthreads();
}

template <typename R>
__global__ void backward_kernel(KArgs<R> g) {
    extern __shared__ double smem_raw[];
    R* smem = reinterpret_cast<R*>(smem_raw);
    const Th th{(int)threadIdx.x, (int)blockDim.x, sizeof(R) == 8 && !((g.m | g.p | g.r) & 7)};
    const int b = blockIdx.x;
    const int m = g.m, p = g.p, r = g.r, n = g.n, N1 = p + m;
    Arena<R> A;
    carve(A, smem, m, p, r, true);
    bool Hzero;
    load_params(g, A, b, n - 1, true, &Hzero, th);
    bzero(A.abar, m, th); bzero(A.Pbar, m * m, th); bzero(A.gT, m * m, th); bzero(A.Psum, m * m, th);
    bzero(A.gZ, p * m, th); bzero(A.gc, m, th); bzero(A.gd, p, th); bzero(A.Hcbar, p * p, th);
    const R* yb = g.y_shared ? g.y : g.y + (size_t)b * n * p;
    const size_t tstride = traj_record_elems(m, p);
    const size_t off_rf = (size_t)m + (size_t)m * m, off_rp = off_rf + (size_t)N1 * (N1 + 1) / 2;
    // a time-varying parameter receives one gradient record per step: hand the step's contribution out and restart
    // the accumulator (the Cholesky pull-back of H_t runs per step as well)
    auto tv = [&](int which) { return (g.tv_mask >> which) & 1u; };
    const bool tv_rq = tv(BKF_PARAM_R) || tv(BKF_PARAM_Q);
    auto flush_tv = [&](int t) {
        if (!g.tv_mask) return;
        const size_t bt = (size_t)b * n + t;
        if (tv(BKF_PARAM_T)) {
            if (g.g_T) bcopy(g.g_T + bt * m * m, A.gT, m * m, th);
            bzero(A.gT, m * m, th);
        }
        if (tv(BKF_PARAM_Z)) {
            if (g.g_Z) bcopy(g.g_Z + bt * p * m, A.gZ, p * m, th);
            bzero(A.gZ, p * m, th);
        }
        if (tv(BKF_PARAM_C)) {
            if (g.g_c) bcopy(g.g_c + bt * m, A.gc, m, th);
            bzero(A.gc, m, th);
        }
        if (tv(BKF_PARAM_D)) {
            if (g.g_d) bcopy(g.g_d + bt * p, A.gd, p, th);
            bzero(A.gd, p, th);
        }
        if (tv(BKF_PARAM_H)) {
            if (g.g_H) {
                if (Hzero) {
                    bcopy(g.g_H + bt * p * p, A.Hcbar, p * p, th);
                } else {
                    chol_pullback(A.Psi, A.Hc, A.Hcbar, p, A.S1, th);
                    bcopy(g.g_H + bt * p * p, A.Psi, p * p, th);
                }
            }
            bzero(A.Hcbar, p * p, th);
        }
    };
    for (int t = n - 1; t >= 0; --t) {
        const size_t bt = (size_t)b * n + t;
        if (g.tv_mask && t < n - 1) load_params(g, A, b, t, false, &Hzero, th);
        const R gcot = g.ll_cot ? g.ll_cot[bt] : R(1);
        bcopy(A.a, g.traj + bt * tstride, m, th);
        bcopy(A.Pc, g.traj + bt * tstride + m, m * m, th);
        bool flag;
        // the two R factors come from the forward sweep's record: no Householder QR in the adjoint
        step(g, A, yb + (size_t)t * p, Hzero, &flag, g.traj + bt * tstride + off_rf, th);
        finish_update(g, A, flag, false, th);
        predict(g, A, (R*)nullptr, (R*)nullptr, g.traj + bt * tstride + off_rp, th);
        // ================= predict adjoint =================
        // Rp_bar = triu(Pbar^T): Rbar(i,j) = Pbar[j*m + i]  ->  strides (rs0, rs1) = (1, m)
        qr_pullback_psi(A.Psi, A.Rp, m, m, A.Pbar, 1, m, A.S1, A.uinv, th);
        if (!tv_rq) {
            for (int idx = th.tid; idx < m * m; idx += th.nt) A.Psum[idx] += A.Psi[idx];
        } else {
            // R_t or Q_t carries a time axis: this step's Psi is pulled back through this step's R and Qc at once.
            // The record goes out for the time-varying one; the other one accumulates in (the unused) Psum.
            rq_cotangents(A, A.Psi, param_ptr_t(g.Rm, g.shared_mask, g.tv_mask, BKF_PARAM_R, b, t, n, (size_t)m * r), m, r, th);
            if (tv(BKF_PARAM_R)) {
                if (g.g_R) bcopy(g.g_R + bt * m * r, A.Pcfbar, m * r, th);
            } else {
                for (int idx = th.tid; idx < m * r; idx += th.nt) A.Psum[idx] += A.Pcfbar[idx];
            }
            if (tv(BKF_PARAM_Q)) {
                if (g.g_Q) {
                    chol_pullback(A.U1, A.Qc, A.S1, r, A.Abar, th);
                    bcopy(g.g_Q + bt * r * r, A.U1, r * r, th);
                }
            } else {
                for (int idx = th.tid; idx < r * r; idx += th.nt) A.Psum[idx] += A.S1[idx];
            }
            __syncthreads();
        }
        // U1 = Psi (T Pcf), (T Pcf)[k,j] = M0[j*m + k]
        mm<SET>(A.U1, m, A.Psi, m, 1, A.M0, 1, m, m, m, m, th);
        // gT += U1 Pcf^T + abar af^T
        mm<ADD>(A.gT, m, A.U1, m, 1, A.Pcf, 1, m, m, m, m, th);
        for (int idx = th.tid; idx < m * m; idx += th.nt) A.gT[idx] = fma(A.abar[idx / m], A.af[idx % m], A.gT[idx]);
        // Pcfbar = T^T U1 ; afbar = T^T abar ; gc += abar
        mm<SET>(A.Pcfbar, m, A.T, 1, m, A.U1, m, 1, m, m, m, th);
        for (int i = th.tid; i < m; i += th.nt) {
            R acc = 0;
            for (int q = 0; q < m; ++q) acc = fma(A.T[q * m + i], A.abar[q], acc);
            A.afbar[i] = acc;
            A.gc[i] += A.abar[i];
        }
        __syncthreads();
        if (flag) {  // degenerate branch (:700-705): (a_f, Pcf0, ll) = (a, Pc, 0)
            bcopy(A.abar, A.afbar, m, th);
            bcopy(A.Pbar, A.Pcfbar, m * m, th);
            flush_tv(t);
            continue;
        }
        // ================= update adjoint =================
        // sbar = KF^T afbar, KF[i,o] = Rf[o, p+i];  loss/logdet cotangents;  the two lower solves with Fc = Rf[:p,:p]^T
        if (th.tid < 32) {  // warp 0 (every launch configuration has at least one full warp)
            const int lane = th.tid;
            const R lossbar = R(-0.5) * gcot, logdetbar = R(-0.5) * gcot * R(p);
            // column-oriented back substitution  x <- U^{-1} x  with U = Fc^T = Rf[:p,:p] (upper), x in shared memory
            auto solve_fcT = [&](R* x) {
                for (int i = p - 1; i >= 0; --i) {
                    __syncwarp();
                    const R xi = x[i] / A.Rf[i * N1 + i];
                    __syncwarp();
                    if (lane == 0) x[i] = xi;
                    for (int q = lane; q < i; q += 32) x[q] = fma(-A.Rf[q * N1 + i], xi, x[q]);
                }
                __syncwarp();
            };
            for (int o = lane; o < p; o += 32) {
                R acc = 0;
                for (int i = 0; i < m; ++i) acc = fma(A.Rf[o * N1 + (p + i)], A.afbar[i], acc);
                A.sbar[o] = acc;
                A.vbar[o] = lossbar * A.inner[o];
                A.t1[o] = lossbar * A.v[o];
            }
            __syncwarp();
            if (p <= 32) {
                R x = lane < p ? A.t1[lane] : R(0);
                x = warp_trisolve<false>(A.Rf, N1, p, x, lane);  // t1 = Fc^{-T} innerbar
                if (lane < p) A.t1[lane] = x;
            } else {
                solve_fcT(A.t1);
            }
            __syncwarp();
            // Bbar's Fc block goes straight into Abar scratch: Fcbar = tril(-t1 inner^T - t2 s^T) + diag(...)
            for (int idx = lane; idx < p * p; idx += 32) {
                int i = idx / p, j = idx - i * p;
                if (j <= i) A.Abar[i * N1 + j] = -A.t1[i] * A.inner[j];
            }
            for (int i = lane; i < p; i += 32) A.sbar[i] += A.t1[i];
            __syncwarp();
            if (p <= 32) {
                R x = lane < p ? A.sbar[lane] : R(0);
                x = warp_trisolve<false>(A.Rf, N1, p, x, lane);  // t2 = Fc^{-T} sbar  (reuse t1)
                __syncwarp();
                if (lane < p) A.t1[lane] = x;
            } else {
                for (int i = lane; i < p; i += 32) A.t1[i] = A.sbar[i];
                solve_fcT(A.t1);
            }
            __syncwarp();
            for (int idx = lane; idx < p * p; idx += 32) {
                int i = idx / p, j = idx - i * p;
                if (j <= i) {
                    R val = A.Abar[i * N1 + j] - A.t1[i] * A.sv[j];
                    if (i == j) val += logdetbar * R(2) / A.Rf[i * N1 + i];
                    A.Abar[i * N1 + j] = val;
                }
            }
            for (int i = lane; i < p; i += 32) A.vbar[i] += A.t1[i];
        }
        __syncthreads();
        // Bbar (lower, N1 x N1) assembled in Abar: [Fcbar 0; KFbar tril(Pcfbar)],  KFbar = afbar s^T
        for (int idx = th.tid; idx < N1 * N1; idx += th.nt) {
            int i = idx / N1, j = idx - i * N1;
            if (i < p) {
                if (j > i) A.Abar[idx] = 0;
            } else if (j < p) {
                A.Abar[idx] = A.afbar[i - p] * A.sv[j];
            } else {
                A.Abar[idx] = (j <= i) ? A.Pcfbar[(i - p) * m + (j - p)] : R(0);
            }
        }
        __syncthreads();
        // Rf_bar = Bbar^T: Rbar(i,j) = Bbar[j*N1 + i] -> strides (1, N1); Psi_f into Psi; then Abar = A0 Psi_f
        qr_pullback_psi(A.Psi, A.Rf, N1, N1, A.Abar, 1, N1, A.S1, A.uinv, th);
        mm<SET>(A.Abar, N1, A.A0, N1, 1, A.Psi, N1, 1, N1, N1, N1, th);
        // blocks: Hcbar += tril(Abar11^T); ZPbar = Abar21^T (p x m); Pbar = Abar22^T + Zm^T ZPbar
        for (int idx = th.tid; idx < p * p; idx += th.nt) {
            int i = idx / p, j = idx - i * p;
            if (Hzero) A.Hcbar[idx] += A.w[i] * A.Abar[j * N1 + i];
            else if (j <= i) A.Hcbar[idx] += A.Abar[j * N1 + i];
        }
        bool blocks_done = false;
        if constexpr (sizeof(R) == 8) {
            if (th.fast) {
                for (int idx = th.tid; idx < m * m; idx += th.nt) {
                    int i = idx / m, j = idx - i * m;
                    A.Pbar[idx] = A.Abar[(p + j) * N1 + (p + i)];
                }
                __syncthreads();
                // Pbar += Zm^T ZPbar,  ZPbar(o, j) = Abar[(p+j), o]
                mm_dmma<ADD, 0>(A.Pbar, m, 1, A.Zm, 1, m, A.Abar + p * N1, 1, N1, m, m, p, th);
                // S1[:p*m] = ZPbar Pc^T
                mm_dmma<SET, 0>(A.S1, m, 1, A.Abar + p * N1, 1, N1, A.Pc, 1, m, p, m, m, th);
                for (int idx = th.tid; idx < p * m; idx += th.nt) {
                    int o = idx / m, j = idx - o * m;
                    A.gZ[idx] = fma(A.w[o], fma(-A.vbar[o], A.a[j], A.S1[idx]), A.gZ[idx]);
                }
                blocks_done = true;
            }
        }
        if (!blocks_done) {
            for (int idx = th.tid; idx < m * m; idx += th.nt) {
                int i = idx / m, j = idx - i * m;
                R acc = A.Abar[(p + j) * N1 + (p + i)];
                for (int o = 0; o < p; ++o) acc = fma(A.Zm[o * m + i], A.Abar[(p + j) * N1 + o], acc);
                A.Pbar[idx] = acc;
            }
            // Zmbar = ZPbar Pc^T - vbar a^T ; gZ += W Zmbar ; abar = afbar - Zm^T vbar ; gd -= vbar
            for (int idx = th.tid; idx < p * m; idx += th.nt) {
                int o = idx / m, j = idx - o * m;
                R acc = 0;
                for (int q = 0; q < m; ++q) acc = fma(A.Abar[(p + q) * N1 + o], A.Pc[j * m + q], acc);
                acc = fma(-A.vbar[o], A.a[j], acc);
                A.gZ[idx] = fma(A.w[o], acc, A.gZ[idx]);
            }
        }
        for (int i = th.tid; i < m; i += th.nt) {
            R acc = A.afbar[i];
            for (int o = 0; o < p; ++o) acc = fma(-A.Zm[o * m + i], A.vbar[o], acc);
            A.abar[i] = acc;
        }
        for (int o = th.tid; o < p; o += th.nt) A.gd[o] -= A.vbar[o];
        __syncthreads();
        flush_tv(t);
    }
    // ---- gradients out ----
    auto put = [&](R* dst, const R* src, size_t sz) {
        if (dst) bcopy(dst + (size_t)b * sz, src, (int)sz, th);
    };
    put(g.g_a0, A.abar, m); put(g.g_P0, A.Pbar, (size_t)m * m);
    if (!tv(BKF_PARAM_C)) put(g.g_c, A.gc, m);
    if (!tv(BKF_PARAM_D)) put(g.g_d, A.gd, p);
    if (!tv(BKF_PARAM_T)) put(g.g_T, A.gT, (size_t)m * m);
    if (!tv(BKF_PARAM_Z)) put(g.g_Z, A.gZ, (size_t)p * m);
    if (g.g_H && !tv(BKF_PARAM_H)) {
        if (Hzero) {
            put(g.g_H, A.Hcbar, (size_t)p * p);
        } else {
            chol_pullback(A.Psi, A.Hc, A.Hcbar, p, A.S1, th);
            put(g.g_H, A.Psi, (size_t)p * p);
        }
    }
    if (!tv_rq) {
        rq_cotangents(A, A.Psum, param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)m * r), m, r, th);
        put(g.g_R, A.Pcfbar, (size_t)m * r);
        if (g.g_Q) {
            chol_pullback(A.Psi, A.Qc, A.S1, r, A.Abar, th);
            put(g.g_Q, A.Psi, (size_t)r * r);
        }
    } else if (!tv(BKF_PARAM_R)) {
        put(g.g_R, A.Psum, (size_t)m * r);  // accumulated over the steps, Q_t time-varying
    } else if (!tv(BKF_PARAM_Q)) {
        if (g.g_Q) {  // factor cotangent accumulated over the steps, pulled back once through the constant Qc
            chol_pullback(A.Psi, A.Qc, A.Psum, r, A.Abar, th);
            put(g.g_Q, A.Psi, (size_t)r * r);
        }
    }
}

template <typename R>
static int launch(const KArgs<R>& g, bool bwd, cudaStream_t stream, const char** err) {
    const size_t bytes = arena_elems(g.m, g.p, g.r, bwd) * sizeof(R);
    if (bytes > 227 * 1024 || (bwd && g.r > g.m) || (bwd && (size_t)g.r * g.r * 2 > (size_t)(g.p + g.m) * (g.p + g.m)) ||
        (bwd && (size_t)g.p * g.p * 2 > (size_t)(g.p + g.m) * (g.p + g.m))) {
        *err = "problem too large for the square-root filter kernels (shared memory)";
        return BKF_ERR_UNSUPPORTED;
    }
    const int N1 = g.p + g.m;
    int threads = N1 <= 12 ? 32 : (N1 <= 24 ? 64 : (N1 <= 40 ? 128 : 256));
    if (!bwd && threads == 256) {
        // forward sweep at large shapes: 3 CTAs x 4 warps per SM sustain ~1.2x the throughput of 2 x 8 warps, but a
        // batch that is only a little over a whole number of waves loses that to the tail
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long w2 = (g.B + 2L * sms - 1) / (2L * sms), w3 = (g.B + 3L * sms - 1) / (3L * sms);
        if (bytes * 3 <= 225 * 1024 && 5 * w3 < 4 * w2) threads = 128;
    }
    auto kern = bwd ? backward_kernel<R> : (threads == 256 ? forward_kernel<R, 256> : forward_kernel<R, 128>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<g.B, threads, bytes, stream>>>(g);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

size_t traj_record_elems_host(int m, int p) { return traj_record_elems(m, p); }
template <typename R> int launch_forward(const KArgs<R>& g, cudaStream_t s, const char** err) { return launch<R>(g, false, s, err); }
template <typename R> int launch_backward(const KArgs<R>& g, cudaStream_t s, const char** err) { return launch<R>(g, true, s, err); }
template int launch_forward<double>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_forward<float>(const KArgs<float>&, cudaStream_t, const char**);
template int launch_backward<double>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_backward<float>(const KArgs<float>&, cudaStream_t, const char**);

}  // namespace sq
}  // namespace bkf
