/* kalman_c.c -- CPU oracle in plain C (OpenMP over series): StandardFilter / UnivariateFilter logp + gradient.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as oracle/kalman_np.py): used by tests/ as a checker and by bench.py as the
 * compiled multi-core CPU baseline ("cpu_baseline", "--impl reference").  Never linked into the product.
 *
 * Restates, with dense loops in reference order,
 *   pymc_extras/statespace/filters/kalman_filter.py:352-360 (masking), :594-617 (StandardFilter.update),
 *   :769-820 (UnivariateFilter), :407-408 (predict), :536 (stabilize), filters/utilities.py:50-59,
 * and the reverse sweep of SURVEY.md Appendix B (what pytensor.grad derives through Scan.L_op).  It is pinned
 * against the golden vectors produced from the reference's own source (tests/test_oracle.py::test_c_oracle_*).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libkalman_c.so oracle/kalman_c.c -lm   (oracle/build_c.py)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOG2PI 1.8378770664093454835606594728112
enum { KIND_STANDARD = 0, KIND_UNIVARIATE = 1 };

/* C[M,N] (op)= A(i,k) B(k,j), generic strides */
static void mm(int mode, double* C, const double* A, int as0, int as1, const double* B, int bs0, int bs1, int M, int N,
               int K) {
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += A[i * as0 + k * as1] * B[k * bs0 + j * bs1];
            if (mode == 0) C[i * N + j] = acc;
            else if (mode == 1) C[i * N + j] += acc;
            else C[i * N + j] -= acc;
        }
}
static void sym_into(double* D, const double* X, const double* Add, double jit, int m) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            double v = 0.5 * (X[i * m + j] + X[j * m + i]);
            if (Add) v += Add[i * m + j];
            if (i == j) v += jit;
            D[i * m + j] = v;
        }
}
static int chol(double* A, int k) {
    for (int j = 0; j < k; ++j) {
        double s = A[j * k + j];
        for (int q = 0; q < j; ++q) s -= A[j * k + q] * A[j * k + q];
        if (!(s > 0)) return 0;
        double l = sqrt(s);
        A[j * k + j] = l;
        for (int i = j + 1; i < k; ++i) {
            double t = A[i * k + j];
            for (int q = 0; q < j; ++q) t -= A[i * k + q] * A[j * k + q];
            A[i * k + j] = t / l;
        }
    }
    return 1;
}
static void chol_solve(const double* L, int k, double* x, int xs) {
    for (int i = 0; i < k; ++i) {
        double t = x[i * xs];
        for (int q = 0; q < i; ++q) t -= L[i * k + q] * x[q * xs];
        x[i * xs] = t / L[i * k + i];
    }
    for (int i = k - 1; i >= 0; --i) {
        double t = x[i * xs];
        for (int q = i + 1; q < k; ++q) t -= L[q * k + i] * x[q * xs];
        x[i * xs] = t / L[i * k + i];
    }
}

typedef struct {
    int m, p, r, kind;
    double jitter, fill;
    const double *T, *Z, *H, *c, *d;
    double *RQR, *a, *P, *w, *ym, *v, *Zm, *Hm, *PZT, *F, *Lf, *K, *KH, *Lm, *X, *Y, *Pf, *af, *wv, *Kst, *Fst, *vst,
        *keep, *Pu;
    int flag;
} Ws;

static double std_update(Ws* s, const double* yt) {
    const int m = s->m, p = s->p;
    double nobs = 0;
    for (int o = 0; o < p; ++o) {
        int miss = isnan(yt[o]) || yt[o] == s->fill;
        s->w[o] = miss ? 0.0 : 1.0;
        s->ym[o] = miss ? 0.0 : yt[o];
        nobs += s->w[o];
    }
    s->flag = (nobs == 0);
    for (int i = 0; i < p * m; ++i) s->Zm[i] = s->w[i / m] * s->Z[i];
    for (int i = 0; i < p * p; ++i) s->Hm[i] = s->w[i / p] * s->H[i];
    for (int o = 0; o < p; ++o) {
        double acc = 0;
        for (int j = 0; j < m; ++j) acc += s->Zm[o * m + j] * s->a[j];
        s->v[o] = s->ym[o] - (s->d[o] + acc);
    }
    mm(0, s->PZT, s->P, m, 1, s->Zm, 1, m, m, p, m);
    mm(0, s->F, s->Zm, m, 1, s->PZT, p, 1, p, p, m);
    for (int i = 0; i < p * p; ++i) {
        s->F[i] += s->Hm[i] + ((i / p) == (i % p) ? s->jitter : 0.0);
        s->Lf[i] = s->F[i];
    }
    chol(s->Lf, p);
    memcpy(s->K, s->PZT, sizeof(double) * m * p);
    memcpy(s->wv, s->v, sizeof(double) * p);
    for (int i = 0; i < m; ++i) chol_solve(s->Lf, p, s->K + i * p, 1);
    chol_solve(s->Lf, p, s->wv, 1);
    mm(0, s->Lm, s->K, p, 1, s->Zm, m, 1, m, m, p);
    for (int i = 0; i < m * m; ++i) s->Lm[i] = ((i / m) == (i % m) ? 1.0 : 0.0) - s->Lm[i];
    for (int i = 0; i < m; ++i) {
        double acc = 0;
        for (int o = 0; o < p; ++o) acc += s->K[i * p + o] * s->v[o];
        s->af[i] = s->a[i] + acc;
    }
    mm(0, s->X, s->Lm, m, 1, s->P, m, 1, m, m, m);
    mm(0, s->Y, s->X, m, 1, s->Lm, 1, m, m, m, m);
    mm(0, s->KH, s->K, p, 1, s->Hm, p, 1, m, p, p);
    mm(0, s->X, s->KH, p, 1, s->K, 1, p, m, m, p);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j)
            s->Pf[i * m + j] = (0.5 * (s->Y[i * m + j] + s->Y[j * m + i]) + 0.5 * (s->X[i * m + j] + s->X[j * m + i])) +
                               (i == j ? s->jitter : 0.0);
    double quad = 0, logdet = 0;
    for (int o = 0; o < p; ++o) {
        quad += s->v[o] * s->wv[o];
        logdet += log(s->Lf[o * p + o]);
    }
    logdet = (p == 1) ? log(s->F[0]) : 2.0 * logdet;
    return s->flag ? 0.0 : -0.5 * (LOG2PI + logdet + quad);
}

static double uni_update(Ws* s, const double* yt) {
    const int m = s->m, p = s->p;
    for (int o = 0; o < p; ++o) {
        int miss = isnan(yt[o]);
        s->w[o] = miss ? 0.0 : 1.0;
        s->ym[o] = miss ? 0.0 : yt[o];
    }
    memcpy(s->af, s->a, sizeof(double) * m);
    memcpy(s->Pu, s->P, sizeof(double) * m * m);
    double llsum = 0;
    int nnz = 0;
    for (int o = 0; o < p; ++o) {
        const double wo = s->w[o];
        const double* z = s->Z + o * m;
        double* K = s->Kst + o * m;
        double yh = 0;
        for (int j = 0; j < m; ++j) yh += wo * z[j] * s->af[j];
        yh += s->d[o];
        const double v = s->ym[o] - yh;
        for (int i = 0; i < m; ++i) {
            double acc = 0;
            for (int j = 0; j < m; ++j) acc += s->Pu[i * m + j] * (wo * z[j]);
            K[i] = acc;
        }
        double F0 = 0;
        for (int j = 0; j < m; ++j) F0 += wo * z[j] * K[j];
        F0 += wo * s->H[o * p + o];
        const int zf = (F0 == 0.0) || (wo == 0.0);
        const double F = F0 + s->jitter, keep = zf ? 0.0 : 1.0;
        for (int i = 0; i < m; ++i) K[i] = K[i] / F * keep;
        for (int i = 0; i < m; ++i) s->af[i] += K[i] * v;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) s->Pu[i * m + j] -= (K[i] * K[j]) * F;
        const double lli = zf ? 0.0 : log(F) + v * v / F;
        llsum += lli;
        nnz += (lli != 0.0);
        s->Fst[o] = F;
        s->vst[o] = v;
        s->keep[o] = keep;
    }
    sym_into(s->Pf, s->Pu, NULL, s->jitter, m);
    return -0.5 * (nnz * LOG2PI + llsum);
}

static void predict(Ws* s) {
    const int m = s->m;
    for (int i = 0; i < m; ++i) {
        double acc = 0;
        for (int j = 0; j < m; ++j) acc += s->T[i * m + j] * s->af[j];
        s->a[i] = acc + s->c[i];
    }
    mm(0, s->X, s->T, m, 1, s->Pf, m, 1, m, m, m);
    mm(0, s->Y, s->X, m, 1, s->T, 1, m, m, m, m);
    sym_into(s->P, s->Y, s->RQR, 0.0, m);
}

static const double* pp(const double* base, unsigned mask, int which, int b, size_t sz) {
    return ((mask >> which) & 1u) ? base : base + (size_t)b * sz;
}

/* One series: forward (storing the predictions) then the reverse sweep.  Gradient pointers may be NULL. */
static void one_series(int kind, int n, int m, int p, int r, const double* y, const double* a0, const double* P0,
                       const double* c, const double* d, const double* T, const double* Z, const double* R,
                       const double* H, const double* Q, const double* cot, double jitter, double fill, double* logp,
                       double* g_a0, double* g_P0, double* g_c, double* g_d, double* g_T, double* g_Z, double* g_R,
                       double* g_H, double* g_Q, int want_grad) {
    const size_t mm_ = (size_t)m * m, mp = (size_t)m * p, pp_ = (size_t)p * p;
    const size_t nfw = 9 * mm_ + 6 * mp + 4 * pp_ + 3 * m + 8 * p;
    const size_t nbw = 5 * mm_ + 5 * mp + 4 * pp_ + 3 * m + 2 * p;
    const size_t ntraj = want_grad ? (size_t)n * (m + mm_) : 0;
    double* buf = (double*)calloc(nfw + nbw + ntraj + 16, sizeof(double));
    double* q = buf;
#define TAKE(n_) (q += (n_), q - (n_))
    Ws s;
    s.m = m; s.p = p; s.r = r; s.kind = kind; s.jitter = jitter; s.fill = fill;
    s.T = T; s.Z = Z; s.H = H; s.c = c; s.d = d;
    s.RQR = TAKE(mm_); s.P = TAKE(mm_); s.Lm = TAKE(mm_); s.X = TAKE(mm_); s.Y = TAKE(mm_); s.Pf = TAKE(mm_);
    s.Pu = TAKE(mm_);
    double* tmp1 = TAKE(mm_);
    double* tmp2 = TAKE(mm_);
    s.Zm = TAKE(mp); s.PZT = TAKE(mp); s.K = TAKE(mp); s.KH = TAKE(mp); s.Kst = TAKE(mp);
    double* tmp_mp = TAKE(mp);
    s.Hm = TAKE(pp_); s.F = TAKE(pp_); s.Lf = TAKE(pp_);
    double* HH = TAKE(pp_);
    s.a = TAKE(m); s.af = TAKE(m);
    double* afbar = TAKE(m);
    s.w = TAKE(p); s.ym = TAKE(p); s.v = TAKE(p); s.wv = TAKE(p); s.Fst = TAKE(p); s.vst = TAKE(p); s.keep = TAKE(p);
    double* vbar = TAKE(p);
    double *Pbar = TAKE(mm_), *G = TAKE(mm_), *gT = TAKE(mm_), *Gsum = TAKE(mm_), *Lbar = TAKE(mm_);
    double *gZ = TAKE(mp), *Kbar = TAKE(mp), *PZTbar = TAKE(mp), *Zmbar = TAKE(mp), *tmp_mp2 = TAKE(mp);
    double *gH = TAKE(pp_), *Fbar = TAKE(pp_), *Finv = TAKE(pp_), *Hmbar = TAKE(pp_);
    double *abar = TAKE(m), *gc = TAKE(m), *pzbar = TAKE(m);
    double *gd = TAKE(p), *unused = TAKE(p);
    double* traj = TAKE(ntraj);
    (void)unused; (void)tmp_mp2;
    /* RQR = sym(R Q R^T) */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            double acc = 0;
            for (int k = 0; k < r; ++k) {
                double rq = 0;
                for (int l = 0; l < r; ++l) rq += R[i * r + l] * Q[l * r + k];
                acc += rq * R[j * r + k];
            }
            tmp1[i * m + j] = acc;
        }
    sym_into(s.RQR, tmp1, NULL, 0.0, m);
    memcpy(s.a, a0, sizeof(double) * m);
    memcpy(s.P, P0, sizeof(double) * mm_);
    double lp = 0;
    for (int t = 0; t < n; ++t) {
        if (want_grad) {
            memcpy(traj + (size_t)t * (m + mm_), s.a, sizeof(double) * m);
            memcpy(traj + (size_t)t * (m + mm_) + m, s.P, sizeof(double) * mm_);
        }
        lp += (kind == KIND_STANDARD) ? std_update(&s, y + (size_t)t * p) : uni_update(&s, y + (size_t)t * p);
        predict(&s);
    }
    *logp = lp;
    if (want_grad) {
        for (int t = n - 1; t >= 0; --t) {
            const double gcot = cot ? cot[t] : 1.0;
            memcpy(s.a, traj + (size_t)t * (m + mm_), sizeof(double) * m);
            memcpy(s.P, traj + (size_t)t * (m + mm_) + m, sizeof(double) * mm_);
            if (kind == KIND_STANDARD) std_update(&s, y + (size_t)t * p); else uni_update(&s, y + (size_t)t * p);
            /* predict adjoint */
            sym_into(G, Pbar, NULL, 0.0, m);
            for (size_t i = 0; i < mm_; ++i) Gsum[i] += G[i];
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < m; ++j) s.Y[i * m + j] = s.Pf[i * m + j] + s.Pf[j * m + i];
            mm(0, s.X, T, m, 1, s.Y, m, 1, m, m, m);
            mm(1, gT, G, m, 1, s.X, m, 1, m, m, m);
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < m; ++j) gT[i * m + j] += abar[i] * s.af[j];
            mm(0, s.X, G, m, 1, T, m, 1, m, m, m);
            mm(0, s.Y, T, 1, m, s.X, m, 1, m, m, m); /* Pfbar */
            for (int i = 0; i < m; ++i) {
                double acc = 0;
                for (int k = 0; k < m; ++k) acc += T[k * m + i] * abar[k];
                afbar[i] = acc;
                gc[i] += abar[i];
            }
            if (kind == KIND_STANDARD) {
                sym_into(G, s.Y, NULL, 0.0, m); /* G2 */
                for (int i = 0; i < m; ++i)
                    for (int j = 0; j < m; ++j) s.Y[i * m + j] = s.P[i * m + j] + s.P[j * m + i];
                mm(0, s.X, s.Lm, m, 1, s.Y, m, 1, m, m, m);
                mm(0, Lbar, G, m, 1, s.X, m, 1, m, m, m);
                mm(0, s.X, G, m, 1, s.Lm, m, 1, m, m, m);
                mm(0, Pbar, s.Lm, 1, m, s.X, m, 1, m, m, m);
                for (int i = 0; i < p; ++i)
                    for (int j = 0; j < p; ++j) HH[i * p + j] = s.Hm[i * p + j] + s.Hm[j * p + i];
                mm(0, tmp_mp, s.K, p, 1, HH, p, 1, m, p, p);
                mm(0, Kbar, G, m, 1, tmp_mp, p, 1, m, p, m);
                mm(0, tmp_mp, G, m, 1, s.K, p, 1, m, p, m);
                mm(0, Hmbar, s.K, 1, p, tmp_mp, p, 1, p, p, m);
                memcpy(abar, afbar, sizeof(double) * m);
                for (int i = 0; i < m; ++i)
                    for (int o = 0; o < p; ++o) Kbar[i * p + o] += afbar[i] * s.v[o];
                for (int o = 0; o < p; ++o) {
                    double acc = 0;
                    for (int i = 0; i < m; ++i) acc += s.K[i * p + o] * afbar[i];
                    vbar[o] = acc;
                }
                mm(2, Kbar, Lbar, m, 1, s.Zm, 1, m, m, p, m);
                mm(0, Zmbar, s.K, 1, p, Lbar, m, 1, p, m, m);
                for (size_t i = 0; i < mp; ++i) Zmbar[i] = -Zmbar[i];
                for (int i = 0; i < p * p; ++i) Finv[i] = ((i / p) == (i % p)) ? 1.0 : 0.0;
                for (int cidx = 0; cidx < p; ++cidx) chol_solve(s.Lf, p, Finv + cidx, p);
                for (int o = 0; o < p; ++o)
                    for (int k = 0; k < p; ++k)
                        Fbar[o * p + k] = s.flag ? 0.0 : -0.5 * gcot * (Finv[k * p + o] - s.wv[o] * s.wv[k]);
                if (!s.flag)
                    for (int o = 0; o < p; ++o) vbar[o] -= gcot * s.wv[o];
                mm(0, PZTbar, Kbar, p, 1, Finv, 1, p, m, p, p);
                mm(2, Fbar, s.K, 1, p, PZTbar, p, 1, p, p, m);
                mm(1, Zmbar, Fbar, p, 1, s.PZT, 1, p, p, m, p);
                mm(1, PZTbar, s.Zm, 1, m, Fbar, p, 1, m, p, p);
                for (size_t i = 0; i < pp_; ++i) Hmbar[i] += Fbar[i];
                mm(1, Pbar, PZTbar, p, 1, s.Zm, m, 1, m, m, p);
                mm(1, Zmbar, PZTbar, 1, p, s.P, m, 1, p, m, m);
                for (int o = 0; o < p; ++o) gd[o] -= vbar[o];
                for (int o = 0; o < p; ++o)
                    for (int j = 0; j < m; ++j) Zmbar[o * m + j] -= vbar[o] * s.a[j];
                for (int j = 0; j < m; ++j)
                    for (int o = 0; o < p; ++o) abar[j] -= s.Zm[o * m + j] * vbar[o];
                for (int o = 0; o < p; ++o)
                    for (int j = 0; j < m; ++j) gZ[o * m + j] += s.w[o] * Zmbar[o * m + j];
                for (int o = 0; o < p; ++o)
                    for (int k = 0; k < p; ++k) gH[o * p + k] += s.w[o] * Hmbar[o * p + k];
            } else {
                sym_into(Pbar, s.Y, NULL, 0.0, m);
                memcpy(abar, afbar, sizeof(double) * m);
                const double llbar = -0.5 * gcot;
                for (int o = p - 1; o >= 0; --o) {
                    if (s.keep[o] == 0.0) continue;
                    const double F = s.Fst[o], v = s.vst[o];
                    const double* K = s.Kst + o * m;
                    const double* z = Z + o * m;
                    for (int i = 0; i < m; ++i)
                        for (int j = 0; j < m; ++j) s.Pu[i * m + j] += (K[i] * K[j]) * F;
                    for (int i = 0; i < m; ++i) s.af[i] -= K[i] * v;
                    double kpk = 0, ka = 0, kk = 0;
                    double* Kb = Kbar; /* m entries */
                    for (int i = 0; i < m; ++i) {
                        double acc = 0, pk = 0;
                        for (int j = 0; j < m; ++j) {
                            acc += (Pbar[i * m + j] + Pbar[j * m + i]) * K[j];
                            pk += Pbar[i * m + j] * K[j];
                        }
                        Kb[i] = -acc * F + abar[i] * v;
                        kpk += K[i] * pk;
                        ka += K[i] * abar[i];
                        kk += Kb[i] * K[i];
                    }
                    double Fb = -kpk + llbar * (1.0 / F - v * v / (F * F));
                    const double vb = ka + llbar * 2.0 * v / F;
                    Fb -= kk / F;
                    for (int i = 0; i < m; ++i) pzbar[i] = Kb[i] / F + Fb * z[i];
                    for (int j = 0; j < m; ++j) {
                        double acc = 0;
                        for (int i = 0; i < m; ++i) acc += s.Pu[i * m + j] * pzbar[i];
                        double zb = Fb * (K[j] * F) + acc - vb * s.af[j];
                        gZ[o * m + j] += zb;
                        abar[j] -= vb * z[j];
                    }
                    for (int i = 0; i < m; ++i)
                        for (int j = 0; j < m; ++j) Pbar[i * m + j] += pzbar[i] * z[j];
                    gH[o * p + o] += Fb;
                    gd[o] -= vb;
                }
            }
        }
        if (g_a0) memcpy(g_a0, abar, sizeof(double) * m);
        if (g_P0) memcpy(g_P0, Pbar, sizeof(double) * mm_);
        if (g_c) memcpy(g_c, gc, sizeof(double) * m);
        if (g_d) memcpy(g_d, gd, sizeof(double) * p);
        if (g_T) memcpy(g_T, gT, sizeof(double) * mm_);
        if (g_Z) memcpy(g_Z, gZ, sizeof(double) * mp);
        if (g_H) memcpy(g_H, gH, sizeof(double) * pp_);
        /* GR = Gsum R */
        double* GR = tmp1;
        mm(0, GR, Gsum, m, 1, R, r, 1, m, r, m);
        if (g_R)
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < r; ++j) {
                    double acc = 0;
                    for (int k = 0; k < r; ++k) acc += GR[i * r + k] * (Q[k * r + j] + Q[j * r + k]);
                    g_R[i * r + j] = acc;
                }
        if (g_Q) mm(0, g_Q, R, 1, r, GR, r, 1, r, r, m);
    }
    (void)tmp2;
    free(buf);
}

/* Batched entry point.  shared_mask bit i (a0,P0,c,d,T,Z,R,H,Q order): parameter i has no batch axis. */
int kalman_c_logp_grad(int kind, int B, int n, int m, int p, int r, int y_shared, unsigned shared_mask, const double* y,
                       const double* a0, const double* P0, const double* c, const double* d, const double* T,
                       const double* Z, const double* R, const double* H, const double* Q, const double* cot,
                       double jitter, double fill, double* logp, double* g_a0, double* g_P0, double* g_c, double* g_d,
                       double* g_T, double* g_Z, double* g_R, double* g_H, double* g_Q, int want_grad) {
    if (r > m) return -1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
#define OUT(ptr, sz) ((ptr) ? (ptr) + (size_t)b * (sz) : NULL)
        one_series(kind, n, m, p, r, y_shared ? y : y + (size_t)b * n * p, pp(a0, shared_mask, 0, b, m),
                   pp(P0, shared_mask, 1, b, (size_t)m * m), pp(c, shared_mask, 2, b, m), pp(d, shared_mask, 3, b, p),
                   pp(T, shared_mask, 4, b, (size_t)m * m), pp(Z, shared_mask, 5, b, (size_t)p * m),
                   pp(R, shared_mask, 6, b, (size_t)m * r), pp(H, shared_mask, 7, b, (size_t)p * p),
                   pp(Q, shared_mask, 8, b, (size_t)r * r), cot ? cot + (size_t)b * n : NULL, jitter, fill, logp + b,
                   OUT(g_a0, m), OUT(g_P0, (size_t)m * m), OUT(g_c, m), OUT(g_d, p), OUT(g_T, (size_t)m * m),
                   OUT(g_Z, (size_t)p * m), OUT(g_R, (size_t)m * r), OUT(g_H, (size_t)p * p), OUT(g_Q, (size_t)r * r),
                   want_grad);
    }
    return 0;
}
