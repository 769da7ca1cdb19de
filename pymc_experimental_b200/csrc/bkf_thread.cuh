// bkf_thread.cuh -- "one thread per series" kernels for tiny state spaces (m <= 4, p = 1): ETS, local level/trend.
//
// Everything of one series (a, P, T, Z, RQR, c, adjoint accumulators) lives in the registers of one thread and
// every loop is fully unrolled, so a step is a few hundred dependent FMAs and the batch axis supplies all the
// parallelism (>= 1e5 series in flight on 148 SMs).  The forward -> backward trajectory is stored batch-innermost
// (SoA: [t][element][series]) so that every trajectory access of a warp is one fully coalesced 256-byte line;
// P_{t|t-1} is exactly symmetric for t >= 1 and is stored packed (m(m+1)/2).  Step 0 uses the caller's a0 / P0
// directly (P0 may be asymmetric; its gradient treats every entry as independent).
//
// Formulas: kalman_filter.py:594-617 (Standard, p = 1), :769-820 (Univariate, p = 1), :407-408 (predict);
// adjoint: SURVEY.md Appendix B.  General (row/column) forms are used throughout, so no symmetry is assumed.
#pragma once
#include "bkf_common.cuh"

#ifndef BKF_TPS_FWD_MINB  // minimum resident CTAs per SM the kernels are compiled for (register cap)
#define BKF_TPS_FWD_MINB 4  // forward: 128 registers (4 CTAs per SM) measured 5 % faster than 168 (3); adjoint: any cap spills
#define BKF_TPS_BWD_MINB 1
#endif

namespace bkf {
namespace tps {

enum { K_STD = 0, K_UNI = 1 };

template <int M>
struct TCfg {
    static constexpr int PACK = M * (M + 1) / 2;
    static constexpr int REC = M + PACK;  // trajectory elements per (series, step)
};

template <typename R, int M>
struct Consts {
    R T[M][M], RQR[M][M], Z[M], c[M], d, H;
};

template <typename R, int M>
struct Step {  // one measurement update
    R K[M], pz[M], zp[M], z[M], af[M], Pf[M][M];
    R F, f0, v, ll, yhat, Hm;
    bool keep, miss;
};

template <typename R, int M, int KIND>
__device__ __forceinline__ void update(Step<R, M>& s, const Consts<R, M>& k, const R (&a)[M], const R (&P)[M][M], R yv,
                                       R fill, R jitter) {
    const bool miss = (KIND == K_UNI) ? isnan(yv) : (isnan(yv) || yv == fill);
    const R w = miss ? R(0) : R(1);
    const R ym = miss ? R(0) : yv;
    s.miss = miss;
    s.Hm = w * k.H;
    R za = 0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        s.z[j] = w * k.Z[j];
        za = fma(s.z[j], a[j], za);
    }
    s.yhat = k.d + za;
    s.v = ym - s.yhat;
    R f0 = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        R r_ = 0, c_ = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            r_ = fma(P[i][j], s.z[j], r_);  // (P z^T)_i
            c_ = fma(s.z[j], P[j][i], c_);  // (z P)_i
        }
        s.pz[i] = r_;
        s.zp[i] = c_;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) f0 = fma(s.z[j], s.pz[j], f0);
    s.f0 = f0;
    if (KIND == K_UNI) {
        const R F0 = f0 + s.Hm;
        const bool zf = (F0 == R(0)) || miss;
        const R F = F0 + jitter;
        s.F = F;
        s.keep = !zf;
        const R kp = zf ? R(0) : R(1);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            s.K[i] = s.pz[i] / F * kp;
            s.af[i] = fma(s.K[i], s.v, a[i]);
        }
        const R lli = zf ? R(0) : log(F) + s.v * s.v / F;
        s.ll = R(-0.5) * ((lli != R(0) ? R(kLog2Pi) : R(0)) + lli);
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const R pu_ij = P[i][j] - (s.K[i] * s.K[j]) * F;
                const R pu_ji = P[j][i] - (s.K[j] * s.K[i]) * F;
                s.Pf[i][j] = R(0.5) * (pu_ij + pu_ji) + (i == j ? jitter : R(0));
            }
    } else {
        const R F = f0 + (s.Hm + jitter);
        s.F = F;
        s.keep = true;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            s.K[i] = s.pz[i] / F;
            s.af[i] = fma(s.K[i], s.v, a[i]);
        }
        s.ll = miss ? R(0) : R(-0.5) * (R(kLog2Pi) + log(F) + s.v * (s.v / F));
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                // Y = (P - K (zP)) - ((P z^T) - K f0) K^T  == (I - K z) P (I - K z)^T
                const R y_ij = fma(-fma(-s.K[i], f0, s.pz[i]), s.K[j], fma(-s.K[i], s.zp[j], P[i][j]));
                const R y_ji = fma(-fma(-s.K[j], f0, s.pz[j]), s.K[i], fma(-s.K[j], s.zp[i], P[j][i]));
                const R khk = R(0.5) * ((s.K[i] * s.Hm) * s.K[j] + (s.K[j] * s.Hm) * s.K[i]);
                s.Pf[i][j] = (R(0.5) * (y_ij + y_ji) + khk) + (i == j ? jitter : R(0));
            }
    }
}

template <typename R, int M>
__device__ __forceinline__ void load_consts(Consts<R, M>& k, const KArgs<R>& g, int b) {
    const R* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)M * M);
    const R* Zg = param_ptr(g.Z, g.shared_mask, BKF_PARAM_Z, b, (size_t)M);
    const R* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * g.r);
    const R* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)g.r * g.r);
    const R* cg = param_ptr(g.c, g.shared_mask, BKF_PARAM_C, b, (size_t)M);
    k.H = *param_ptr(g.H, g.shared_mask, BKF_PARAM_H, b, 1);
    k.d = *param_ptr(g.d, g.shared_mask, BKF_PARAM_D, b, 1);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        k.Z[i] = Zg[i];
        k.c[i] = cg[i];
#pragma unroll
        for (int j = 0; j < M; ++j) k.T[i][j] = Tg[i * M + j];
    }
    const int r = g.r;
    R raw[M][M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
            R acc = 0;
            for (int q = 0; q < r; ++q) {
                R rq = 0;
                for (int l = 0; l < r; ++l) rq = fma(Rg[i * r + l], Qg[l * r + q], rq);
                acc = fma(rq, Rg[j * r + q], acc);
            }
            raw[i][j] = acc;
        }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) k.RQR[i][j] = R(0.5) * (raw[i][j] + raw[j][i]);
}

// The observation-side / intercept parameters of step t when they carry a time axis (tv_mask on c, d, Z, H: regression
// components, time-varying intercepts and measurement noise; T, R, Q with a time axis take the generic kernels).
template <typename R, int M>
__device__ __forceinline__ void refresh_tv(Consts<R, M>& k, const KArgs<R>& g, int b, int t) {
    const int n = g.n;
    const R* Zt = param_ptr_t(g.Z, g.shared_mask, g.tv_mask, BKF_PARAM_Z, b, t, n, (size_t)M);
    const R* ct = param_ptr_t(g.c, g.shared_mask, g.tv_mask, BKF_PARAM_C, b, t, n, (size_t)M);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        k.Z[i] = Zt[i];
        k.c[i] = ct[i];
    }
    k.H = *param_ptr_t(g.H, g.shared_mask, g.tv_mask, BKF_PARAM_H, b, t, n, 1);
    k.d = *param_ptr_t(g.d, g.shared_mask, g.tv_mask, BKF_PARAM_D, b, t, n, 1);
}

template <typename R, int M, int KIND, bool TV = false>
__global__ void __launch_bounds__(128, BKF_TPS_FWD_MINB) forward_kernel(KArgs<R> g) {
    using C = TCfg<M>;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    const int n = g.n;
    const size_t B = g.B;
    Consts<R, M> k;
    load_consts<R, M>(k, g, b);
    R a[M], P[M][M];
    {
        const R* a0g = param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)M);
        const R* P0g = param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)M * M);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            a[i] = a0g[i];
#pragma unroll
            for (int j = 0; j < M; ++j) P[i][j] = P0g[i * M + j];
        }
    }
    const R* yb = g.y_shared ? g.y : g.y + (size_t)b * n;
    R logp = 0;
    Step<R, M> s;
    R ynext = yb[0];
    for (int t = 0; t < n; ++t) {
        const size_t bt = (size_t)b * n + t;
        const R yt = ynext;
        if (t + 1 < n) ynext = yb[t + 1];  // one step ahead: the load latency hides behind this step's arithmetic
        if (g.traj && t > 0) {  // batch-innermost SoA record: [t][e][b]
            R* tr = g.traj + (size_t)t * C::REC * B + b;
            int e = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) tr[(size_t)(e++) * B] = a[i];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) tr[(size_t)(e++) * B] = P[i][j];
        }
        if (g.pred_a)
#pragma unroll
            for (int i = 0; i < M; ++i) g.pred_a[bt * M + i] = a[i];
        if (g.pred_P)
#pragma unroll
            for (int i = 0; i < M * M; ++i) g.pred_P[bt * M * M + i] = P[i / M][i % M];
        if (TV) refresh_tv<R, M>(k, g, b, t);
        update<R, M, KIND>(s, k, a, P, yt, g.fill, g.jitter);
        logp += s.ll;
        if (g.ll) g.ll[bt] = s.ll;
        if (g.obs_mu) g.obs_mu[bt] = s.yhat;
        if (g.obs_cov) g.obs_cov[bt] = s.F;
        if (g.filt_a)
#pragma unroll
            for (int i = 0; i < M; ++i) g.filt_a[bt * M + i] = s.af[i];
        if (g.filt_P)
#pragma unroll
            for (int i = 0; i < M * M; ++i) g.filt_P[bt * M * M + i] = s.Pf[i / M][i % M];
        // predict
        R X[M][M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            R acc = k.c[i];
            R ta = 0;
#pragma unroll
            for (int q = 0; q < M; ++q) ta = fma(k.T[i][q], s.af[q], ta);
            a[i] = ta + acc;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                R x = 0;
#pragma unroll
                for (int q = 0; q < M; ++q) x = fma(k.T[i][q], s.Pf[q][j], x);
                X[i][j] = x;
            }
        }
        R Y[M][M];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                R y = 0;
#pragma unroll
                for (int q = 0; q < M; ++q) y = fma(X[i][q], k.T[j][q], y);
                Y[i][j] = y;
            }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) P[i][j] = R(0.5) * (Y[i][j] + Y[j][i]) + k.RQR[i][j];
    }
    if (g.logp) g.logp[b] = logp;
    if (g.status) g.status[b] = isfinite(logp) ? 0 : BKF_STATUS_NONFINITE;
}

template <typename R, int M, int KIND, bool TV = false>
__global__ void __launch_bounds__(128, BKF_TPS_BWD_MINB) backward_kernel(KArgs<R> g) {
    using C = TCfg<M>;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    const int n = g.n;
    const size_t B = g.B;
    Consts<R, M> k;
    load_consts<R, M>(k, g, b);
    const R* a0g = param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)M);
    const R* P0g = param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)M * M);
    const R* yb = g.y_shared ? g.y : g.y + (size_t)b * n;
    R abar[M], Pbar[M][M], gT[M][M], Gsum[M][M], gZ[M], gc[M];
    R gH = 0, gd = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        abar[i] = 0; gZ[i] = 0; gc[i] = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) { Pbar[i][j] = 0; gT[i][j] = 0; Gsum[i][j] = 0; }
    }
    Step<R, M> s;
    // Software pipeline: the record, observation and cotangent of step t-1 are requested at the top of step t, so
    // their HBM latency hides behind a full step of arithmetic (the kernel runs at ~8 warps/SM: latency is not hidden
    // by occupancy).
    R rec[C::REC], ynext, cnext;
    auto fetch = [&](int t) {
        if (t > 0) {
            const R* tr = g.traj + (size_t)t * C::REC * B + b;
#pragma unroll
            for (int e = 0; e < C::REC; ++e) rec[e] = __ldcs(tr + (size_t)e * B);
        }
        ynext = yb[t];
        cnext = g.ll_cot ? g.ll_cot[(size_t)b * n + t] : R(1);
    };
    fetch(n - 1);
    for (int t = n - 1; t >= 0; --t) {
        const R gcot = cnext, yt = ynext;
        R a[M], P[M][M];
        if (t > 0) {
            int e = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) a[i] = rec[e++];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    const R v = rec[e++];
                    P[i][j] = v;
                    P[j][i] = v;
                }
            fetch(t - 1);
        } else {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                a[i] = a0g[i];
#pragma unroll
                for (int j = 0; j < M; ++j) P[i][j] = P0g[i * M + j];
            }
        }
        if (TV) refresh_tv<R, M>(k, g, b, t);
        update<R, M, KIND>(s, k, a, P, yt, g.fill, g.jitter);
        // ---- predict adjoint ----
        R G[M][M], W[M][M], Pfb[M][M], afbar[M];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                G[i][j] = R(0.5) * (Pbar[i][j] + Pbar[j][i]);
                Gsum[i][j] += G[i][j];
            }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                R w = 0;
#pragma unroll
                for (int q = 0; q < M; ++q) w = fma(G[i][q], k.T[q][j], w);
                W[i][j] = w;
            }
#pragma unroll
        for (int i = 0; i < M; ++i) {
            R ab = 0;
#pragma unroll
            for (int q = 0; q < M; ++q) ab = fma(k.T[q][i], abar[q], ab);
            afbar[i] = ab;
            gc[i] += abar[i];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                R pf = 0, gt = 0;
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    pf = fma(k.T[q][i], W[q][j], pf);
                    gt = fma(W[i][q], s.Pf[q][j] + s.Pf[j][q], gt);
                }
                Pfb[i][j] = pf;
                gT[i][j] += fma(abar[i], s.af[j], gt);
            }
        }
        // ---- update adjoint ----
        R G2[M][M];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) G2[i][j] = R(0.5) * (Pfb[i][j] + Pfb[j][i]);
        if (KIND == K_UNI) {
            if (s.keep) {
                const R llbar = R(-0.5) * gcot, F = s.F, v = s.v;
                R Kbar[M], kpk = 0, ka = 0, kk = 0;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    R pk = 0;
#pragma unroll
                    for (int j = 0; j < M; ++j) pk = fma(G2[i][j], s.K[j], pk);
                    Kbar[i] = fma(R(-2) * F, pk, afbar[i] * v);
                    kpk = fma(s.K[i], pk, kpk);
                    ka = fma(s.K[i], afbar[i], ka);
                    kk = fma(Kbar[i], s.K[i], kk);
                }
                const R Fbar = -kpk + llbar * (R(1) / F - v * v / (F * F)) - kk / F;
                const R vbar = ka + llbar * R(2) * v / F;
                R pzbar[M];
#pragma unroll
                for (int i = 0; i < M; ++i) pzbar[i] = Kbar[i] / F + Fbar * s.z[i];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    R zb = Fbar * s.pz[j];
#pragma unroll
                    for (int i = 0; i < M; ++i) zb = fma(P[i][j], pzbar[i], zb);
                    zb = fma(-vbar, a[j], zb);
                    gZ[j] += zb;
                    abar[j] = fma(-vbar, s.z[j], afbar[j]);
                }
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int j = 0; j < M; ++j) Pbar[i][j] = fma(pzbar[i], s.z[j], G2[i][j]);
                gH += Fbar;
                gd -= vbar;
            } else {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    abar[i] = afbar[i];
#pragma unroll
                    for (int j = 0; j < M; ++j) Pbar[i][j] = G2[i][j];
                }
            }
        } else {
            const R F = s.F, v = s.v, f0 = s.f0, Hm = s.Hm, wv = v / F;
            R gk[M], u1[M], kgk = 0, vbar = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                R g1 = 0, g2 = 0;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    g1 = fma(G2[i][j], s.K[j], g1);
                    g2 = fma(G2[i][j], (s.pz[j] + s.zp[j]) - R(2) * f0 * s.K[j], g2);
                }
                gk[i] = g1;
                u1[i] = g2;
                kgk = fma(s.K[i], g1, kgk);
                vbar = fma(s.K[i], afbar[i], vbar);
            }
            R Fbar = 0;
            if (!s.miss) {
                Fbar = R(-0.5) * gcot * (R(1) / F - wv * wv);
                vbar -= gcot * wv;
            }
            R pztbar[M], kp = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const R Kbar = fma(R(2) * Hm, gk[i], afbar[i] * v) - u1[i];
                pztbar[i] = Kbar / F;
                kp = fma(s.K[i], pztbar[i], kp);
            }
            Fbar -= kp;
            const R Hmbar = kgk + Fbar;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                R zm = 0;
#pragma unroll
                for (int i = 0; i < M; ++i) zm = fma(gk[i], P[i][j] + P[j][i], zm);
                zm = -(zm - kgk * (s.zp[j] + s.pz[j]));
                zm = fma(Fbar, s.pz[j], zm);
                pztbar[j] = fma(Fbar, s.z[j], pztbar[j]);
                R acc = zm;
                // (P^T pztbar)_j needs the final pztbar: accumulated below
                gZ[j] += s.miss ? R(0) : acc;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                R zm = 0;
#pragma unroll
                for (int i = 0; i < M; ++i) zm = fma(pztbar[i], P[i][j], zm);
                zm = fma(-vbar, a[j], zm);
                gZ[j] += s.miss ? R(0) : zm;
                abar[j] = fma(-vbar, s.z[j], afbar[j]);
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    R pb = G2[i][j] - s.z[i] * gk[j] - gk[i] * s.z[j];
                    pb = fma(s.z[i] * s.z[j], kgk, pb);
                    Pbar[i][j] = fma(pztbar[i], s.z[j], pb);
                }
            gd -= vbar;
            gH += s.miss ? R(0) : Hmbar;
        }
        if (TV) {  // a gradient with a time axis is written out and cleared after every step
            const size_t bt = (size_t)b * n + t;
            if ((g.tv_mask >> BKF_PARAM_C) & 1u) {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    if (g.g_c) g.g_c[bt * M + i] = gc[i];
                    gc[i] = 0;
                }
            }
            if ((g.tv_mask >> BKF_PARAM_Z) & 1u) {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    if (g.g_Z) g.g_Z[bt * M + i] = gZ[i];
                    gZ[i] = 0;
                }
            }
            if ((g.tv_mask >> BKF_PARAM_D) & 1u) {
                if (g.g_d) g.g_d[bt] = gd;
                gd = 0;
            }
            if ((g.tv_mask >> BKF_PARAM_H) & 1u) {
                if (g.g_H) g.g_H[bt] = gH;
                gH = 0;
            }
        }
    }
    auto put = [&](R* dst, const R* src, int sz) {
        if (dst)
            for (int i = 0; i < sz; ++i) dst[(size_t)b * sz + i] = src[i];
    };
    const bool tvC = TV && ((g.tv_mask >> BKF_PARAM_C) & 1u), tvZ = TV && ((g.tv_mask >> BKF_PARAM_Z) & 1u),
               tvD = TV && ((g.tv_mask >> BKF_PARAM_D) & 1u), tvH = TV && ((g.tv_mask >> BKF_PARAM_H) & 1u);
    put(g.g_a0, abar, M); put(g.g_P0, &Pbar[0][0], M * M); put(g.g_T, &gT[0][0], M * M);
    if (!tvC) put(g.g_c, gc, M);
    if (!tvZ) put(g.g_Z, gZ, M);
    if (g.g_d && !tvD) g.g_d[b] = gd;
    if (g.g_H && !tvH) g.g_H[b] = gH;
    const int r = g.r;
    const R* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * r);
    const R* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    if (g.g_R)
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < r; ++j) {
                R acc = 0;
                for (int q = 0; q < r; ++q) {
                    R gr = 0;
                    for (int l = 0; l < M; ++l) gr = fma(Gsum[i][l], Rg[l * r + q], gr);
                    acc = fma(gr, Qg[q * r + j] + Qg[j * r + q], acc);
                }
                g.g_R[(size_t)b * M * r + i * r + j] = acc;
            }
    if (g.g_Q)
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < r; ++j) {
                R acc = 0;
                for (int q = 0; q < M; ++q) {
                    R gr = 0;
                    for (int l = 0; l < M; ++l) gr = fma(Gsum[q][l], Rg[l * r + j], gr);
                    acc = fma(Rg[q * r + i], gr, acc);
                }
                g.g_Q[(size_t)b * r * r + i * r + j] = acc;
            }
}

template <typename R, int M>
int launch(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err) {
    const int block = 128, grid = (g.B + block - 1) / block;
    const bool tv = g.tv_mask != 0;  // c, d, Z, H only (tv_supported)
#define BKF_TPS_LAUNCH(KIND)                                                                    \
    if (backward) {                                                                             \
        if (tv) backward_kernel<R, M, KIND, true><<<grid, block, 0, stream>>>(g);               \
        else backward_kernel<R, M, KIND, false><<<grid, block, 0, stream>>>(g);                 \
    } else {                                                                                    \
        if (tv) forward_kernel<R, M, KIND, true><<<grid, block, 0, stream>>>(g);                \
        else forward_kernel<R, M, KIND, false><<<grid, block, 0, stream>>>(g);                  \
    }
    if (g.kind == BKF_UNIVARIATE) { BKF_TPS_LAUNCH(K_UNI) } else { BKF_TPS_LAUNCH(K_STD) }
#undef BKF_TPS_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace tps
}  // namespace bkf
