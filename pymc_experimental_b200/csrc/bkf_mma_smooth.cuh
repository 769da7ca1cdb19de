// bkf_mma_smooth.cuh -- RTS smoother on the FP64 tensor-core path (fragment layout of bkf_mma.cuh), 9 <= m <= 16.
//
// kalman_smoother.py:107-126 per step (t = n-2 .. 0), with (a, P) = filtered (a_t|t, P_t|t):
//   a_hat = T a (no intercept, quirk Q8);  P_hat = sym(T P T^T) + sym(R Q R^T) + eps I
//   G = (pinv(P_hat) T P)^T;  a_s = a + G (a_s+ - a_hat);  P_s = P + sym(G (P_s+ - P_hat) G^T) + eps I + eps0 I
//
// pinv policy: P_hat is symmetric positive definite after the jitter, so pinv(P_hat) is the
// inverse whenever no singular value falls under NumPy's 1e-15 cut-off.  Here the 16 x 16 inverse is built from
// 8 x 8 tiles: P_hat = [[A, B], [B^T, C]],  S = C - B^T A^-1 B,
//   P_hat^-1 = [[A^-1 + W S^-1 W^T, -W S^-1], [-S^-1 W^T, S^-1]],  W = A^-1 B,
// where the two 8 x 8 inverses are in-register Gauss-Jordan sweeps (quad shuffles) and every tile product is two
// DMMAs.  The pivots of the two sweeps are exactly the pivots of an unblocked Cholesky / LDL^T of P_hat, so the
// ill-conditioning test of the sub-warp smoother carries over unchanged: a pivot <= 0 or min/max < 1e-13 flags the
// series (BKF_STATUS_SMOOTHER_ILL) and the host re-runs the call with the generic kernel's eigen-decomposition pinv.
// Steps whose pivots span more than 1e3 get one step of iterative refinement on G (two more products).
#pragma once
#include "bkf_mma.cuh"

namespace bkf {
namespace tc {

// D (+)= X Y^T for 8 x 8 tiles held as accumulator fragments (2 DMMAs)
__device__ __forceinline__ void tile_nt(double (&D)[2], const double (&X)[2], const double (&Y)[2]) {
    dmma(D[0], D[1], X[0], Y[0]);
    dmma(D[0], D[1], X[1], Y[1]);
}

// In-place inverse of a symmetric positive definite 8 x 8 tile (Gauss-Jordan without pivoting).  The first `nreal`
// pivots enter the conditioning statistics (the rest belong to the identity padding).
__device__ __forceinline__ void inv8(double (&a)[2], int g, int t, int nreal, double& dmin, double& dmax, int& ill) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double src = (k & 1) ? a[1] : a[0];
        const double piv = __shfl_sync(FULL, src, 4 * k + (k >> 1));  // a[k][k]
        const double ck = __shfl_sync(FULL, src, 4 * g + (k >> 1));   // a[g][k]
        const double rk0 = __shfl_sync(FULL, a[0], 4 * k + t);        // a[k][2t]
        const double rk1 = __shfl_sync(FULL, a[1], 4 * k + t);        // a[k][2t+1]
        if (k < nreal) {
            if (!(piv > 0.0)) ill = 1;
            dmin = fmin(dmin, piv);
            dmax = fmax(dmax, piv);
        }
        const double ip = 1.0 / piv;
        const double s0 = rk0 * ip, s1 = rk1 * ip;
        const double nc = -ck * ip;
        const bool rowk = g == k;
        a[0] = rowk ? (2 * t == k ? ip : s0) : (2 * t == k ? nc : fma(-ck, s0, a[0]));
        a[1] = rowk ? (2 * t + 1 == k ? ip : s1) : (2 * t + 1 == k ? nc : fma(-ck, s1, a[1]));
    }
}

#ifndef BKF_TC_SM_MINB
#define BKF_TC_SM_MINB 1
#endif
__global__ void __launch_bounds__(64, BKF_TC_SM_MINB) smoother_kernel(KArgs<double> G, double jitter0, int* ill_list) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= G.B) return;
    double* sc = smem + warp * SCRATCH;
    const int m = G.m, n = G.n;

    Frag Tc, C0;
    load_frag<false>(Tc, param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m), m, g, t);
    rqr_frag(C0, param_ptr(G.Rm, G.shared_mask, BKF_PARAM_R, b, (size_t)m * G.r),
             param_ptr(G.Q, G.shared_mask, BKF_PARAM_Q, b, (size_t)G.r * G.r), m, G.r, g, t);
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int e = 0; e < 2; ++e)
            if (g == 2 * t + e) C0.v[I][I][e] = (8 * I + g < m) ? C0.v[I][I][e] + G.jitter : 1.0;  // identity on the padding

    const double* fa = G.sm_filt_a + (size_t)b * n * m;
    const double* fP = G.sm_filt_P + (size_t)b * n * m * m;
    double* oa = G.sm_a ? G.sm_a + (size_t)b * n * m : nullptr;
    double* oP = G.sm_P ? G.sm_P + (size_t)b * n * m * m : nullptr;

    // t = n-1: smoothed == filtered
    RVec as_r = load_rvec(fa + (size_t)(n - 1) * m, m, g);
    Frag Ps;
    load_frag<false>(Ps, fP + (size_t)(n - 1) * m * m, m, g, t);
    if (oa) store_rvec(oa + (size_t)(n - 1) * m, as_r, m, g, t);
    if (oP) store_frag(oP + (size_t)(n - 1) * m * m, Ps, m, g, t, 1.0);

    int ill = 0;
    Frag P, Pn;
    RVec a_r, an_r;
    CVec a_c, an_c;
    if (n >= 2) {
        load_frag<false>(Pn, fP + (size_t)(n - 2) * m * m, m, g, t);
        an_r = load_rvec(fa + (size_t)(n - 2) * m, m, g);
        an_c = load_cvec(fa + (size_t)(n - 2) * m, m, t);
    }
#pragma unroll 1
    for (int ts = n - 2; ts >= 0; --ts) {
        P = Pn;
        a_r = an_r;
        a_c = an_c;
        if (ts > 0) {  // prefetch one step ahead
            load_frag<false>(Pn, fP + (size_t)(ts - 1) * m * m, m, g, t);
            an_r = load_rvec(fa + (size_t)(ts - 1) * m, m, g);
            an_c = load_cvec(fa + (size_t)(ts - 1) * m, m, t);
        }
        // ---- X = T P ; P_hat = sym(X T^T) + sym(R Q R^T) + eps I ----
        Frag X, Y, Yt, Ph;
        mm_nt<false>(X, Tc, P);
        Y.v[0][0][0] = Y.v[0][0][1] = Y.v[1][1][0] = Y.v[1][1][1] = 0.0;
        Y.v[0][1][0] = C0.v[0][1][0];
        Y.v[0][1][1] = C0.v[0][1][1];
        mm_nt_upper(Y, X, Tc);
        transpose_upper(Yt, Y, sc, g, t);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            Ph.v[0][0][e] = fma(0.5, Y.v[0][0][e] + Yt.v[0][0][e], C0.v[0][0][e]);
            Ph.v[1][1][e] = fma(0.5, Y.v[1][1][e] + Yt.v[1][1][e], C0.v[1][1][e]);
            Ph.v[0][1][e] = Y.v[0][1][e];
            Ph.v[1][0][e] = Yt.v[1][0][e];
        }
        // w = a_s+ - T a
        const RVec Ta_r = mv(Tc, a_c);
        RVec w_r;
        w_r.v[0] = as_r.v[0] - Ta_r.v[0];
        w_r.v[1] = as_r.v[1] - Ta_r.v[1];
        const CVec w_c = row2col(w_r, t);
        // ---- P_hat^-1 by 8 x 8 tiles ----
        double dmin = 1e300, dmax = 0.0;
        Frag Pi;
        double Ai[2] = {Ph.v[0][0][0], Ph.v[0][0][1]};
        inv8(Ai, g, t, 8, dmin, dmax, ill);
        double W1[2] = {0.0, 0.0}, W1t[2] = {0.0, 0.0};
        tile_nt(W1, Ai, Ph.v[1][0]);   // A^-1 B      (Ph10 = B^T)
        tile_nt(W1t, Ph.v[1][0], Ai);  // B^T A^-1
        double nB[2] = {-Ph.v[1][0][0], -Ph.v[1][0][1]};
        double Si[2] = {Ph.v[1][1][0], Ph.v[1][1][1]};
        tile_nt(Si, W1t, nB);  // S = C - B^T A^-1 B
        {   // the Schur complement is symmetric up to rounding; the sweep reads both triangles
            inv8(Si, g, t, m - 8, dmin, dmax, ill);
        }
        if (dmin < 1e-13 * dmax) ill = 1;
        double nW1[2] = {-W1[0], -W1[1]};
        Pi.v[1][1][0] = Si[0];
        Pi.v[1][1][1] = Si[1];
        Pi.v[1][0][0] = Pi.v[1][0][1] = 0.0;
        tile_nt(Pi.v[1][0], Si, nW1);  // -S^-1 W^T
        Pi.v[0][1][0] = Pi.v[0][1][1] = 0.0;
        tile_nt(Pi.v[0][1], nW1, Si);  // -W S^-1
        Pi.v[0][0][0] = Ai[0];
        Pi.v[0][0][1] = Ai[1];
        tile_nt(Pi.v[0][0], Pi.v[0][1], nW1);  // A^-1 + W S^-1 W^T
        // ---- G = (P_hat^-1 X)^T = X^T P_hat^-1 ----
        Frag Xt, Gm;
        transpose(Xt, X, sc, g, t);
        mm_nt<false>(Gm, Xt, Pi);
        if (dmin < 1e-3 * dmax) {
            // One step of iterative refinement, G += (X^T - G P_hat) P_hat^-1, when the pivots span more than three
            // decades: the explicit tile inverse loses ~cond(P_hat) eps, the refined G is as accurate as a
            // Cholesky solve (warp-uniform branch; well-conditioned steps skip the two extra products).
            Frag Rr;
            mm_nt<false>(Rr, Gm, Ph);
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J < 2; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) Rr.v[I][J][e] = Xt.v[I][J][e] - Rr.v[I][J][e];
            mm_nt<true>(Gm, Rr, Pi);
        }
        // ---- a_s = a + G w ----
        const RVec gw = mv(Gm, w_c);
        as_r.v[0] = a_r.v[0] + gw.v[0];
        as_r.v[1] = a_r.v[1] + gw.v[1];
        // ---- P_s = P + sym(G D G^T) + eps I + eps0 I,  D = P_s+ - P_hat ----
        Frag D, V, M2, M2t;
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) D.v[I][J][e] = Ps.v[I][J][e] - Ph.v[I][J][e];
        mm_nt<false>(V, Gm, D);  // G D (D symmetric)
        M2.v[0][0][0] = M2.v[0][0][1] = M2.v[1][1][0] = M2.v[1][1][1] = 0.0;
        M2.v[0][1][0] = P.v[0][1][0];
        M2.v[0][1][1] = P.v[0][1][1];
        mm_nt_upper(M2, V, Gm);  // G D G^T (+ P on the off-diagonal tile)
        transpose_upper(M2t, M2, sc, g, t);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double d0 = fma(0.5, M2.v[0][0][e] + M2t.v[0][0][e], P.v[0][0][e]);
            double d1 = fma(0.5, M2.v[1][1][e] + M2t.v[1][1][e], P.v[1][1][e]);
            if (g == 2 * t + e) {
                d0 = (d0 + G.jitter) + jitter0;
                d1 = (8 + g < m) ? (d1 + G.jitter) + jitter0 : 0.0;
            }
            Ps.v[0][0][e] = d0;
            Ps.v[1][1][e] = d1;
            Ps.v[0][1][e] = M2.v[0][1][e];
            Ps.v[1][0][e] = M2t.v[1][0][e];
        }
        if (oa) store_rvec(oa + (size_t)ts * m, as_r, m, g, t);
        if (oP) store_frag(oP + (size_t)ts * m * m, Ps, m, g, t, 1.0);
    }
    if (ill && lane == 0) {
        if (G.status) G.status[b] |= BKF_STATUS_SMOOTHER_ILL;
        ill_list[1 + atomicAdd(ill_list, 1)] = b;  // ill_list[0]: how many, then which
    }
}

}  // namespace tc
}  // namespace bkf
