// bkf_generic.cu -- generic (any m, p, r that fits shared memory) kernels: one warp, or one CTA of 2, 4 or 8 warps, per series.
//
// One group of threads (Lane<W>: a warp, W = 1, or a CTA of W warps) owns one time-sequential filter; every matrix of
// that series lives in the group's shared-memory arena and the group's lanes split each small dense operation.  This
// family is the always-available correct path (and the fallback for shapes without a specialised kernel: several observed
// series under the Standard / Univariate filters, more than 16 states, float32 above the sub-warp sizes, T / R / Q with a
// time axis); the hot shapes of BASELINE.json are served by the register-tiled kernels in bkf_mma / bkf_subwarp /
// bkf_thread / bkf_sqrt_warp.  The launcher picks W from the arena size (pick_w): large arenas leave an SM only a few
// series, which then get the warps the SM would otherwise leave idle.
//
// Formulas: see bkf_common.cuh for the reference file:line map; the adjoint recurrences are those of
// SURVEY.md Appendix B (Standard) and their rank-1 analogue (Univariate).
//
// The file is compiled twice: as namespace gen (this unit: scalar products, float and double) and, through
// bkf_generic_tc.cu with BKF_GEN_DMMA defined, as namespace gent (double only), where every dense product of at least
// 8 x 8 x 8 runs on the FP64 tensor pipe.  The dispatcher sends m >= 16 to gent: the tile products are 1.6 x (m = 16) to
// 2.1 x (m = 32) faster there, but they cost the kernels ~120 more registers, which halves the occupancy the small shapes
// live on (m = 8 and m = 4 measured 1.9 x slower in one combined build).
#include "bkf_common.cuh"
#include "bkf_launch.h"

#ifndef BKF_GEN_NS
#define BKF_GEN_NS gen
#endif
#ifndef BKF_GEN_W
#define BKF_GEN_W 1
#endif
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace bkf {
namespace BKF_GEN_NS {

enum { SET = 0, ADD = 1, SUB = 2 };

// Who runs one series.  W = 1: one warp (several series per CTA, the original mapping; the barrier is __syncwarp).
// W > 1: one CTA of W warps (one series per CTA, the barrier is __syncthreads): chosen by the launcher when the arena of
// a series is so large that only a few of them fit one SM, so that the SM still has W warps per series to issue from.
// L.v is the thread's index within the group, 32 * W the stride of every element loop.
template <int W>
struct Lane { int v; };

template <int W>
__device__ __forceinline__ void gsync() {
    if constexpr (W == 1) __syncwarp();
    else __syncthreads();
}

// One element global -> shared by the asynchronous copy unit (LDGSTS): the adjoint fetches the record of the NEXT step while it works on this one.
template <typename R>
__device__ __forceinline__ void cp_async_elem(R* dst_smem, const R* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(src), "n"(sizeof(R)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Three sums over the group at once (every thread gets the totals).  red: 3 * W values of scratch.
template <int W, typename R>
__device__ __forceinline__ void gsum3(R& a, R& b, R& c, R* red, int lane) {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if constexpr (W > 1) {
        if ((lane & 31) == 0) {
            red[3 * (lane >> 5)] = a; red[3 * (lane >> 5) + 1] = b; red[3 * (lane >> 5) + 2] = c;
        }
        __syncthreads();
        a = red[0]; b = red[1]; c = red[2];
#pragma unroll
        for (int w = 1; w < W; ++w) { a += red[3 * w]; b += red[3 * w + 1]; c += red[3 * w + 2]; }
        __syncthreads();
    }
}

// The same product on the FP64 tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): the warp walks the 8 x 8 output tiles,
// every DMMA takes one A and one B element per lane straight from shared memory (arbitrary strides, zero-filled borders),
// 256 FMAs for two loads per lane where the scalar loop below spends two loads per FMA.  Two tiles are in flight per
// round so that their dependent accumulator chains overlap.  Same sums in a different order: equal to rounding.
template <int MODE, int W>
__device__ __forceinline__ void mm_dmma(double* __restrict__ C, const double* __restrict__ A, int as0, int as1,
                                        const double* __restrict__ Bm, int bs0, int bs1, int M, int N, int K, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    const int g = (lane & 31) >> 2, t = lane & 3;
    const int tn = (N + 7) >> 3, ntile = ((M + 7) >> 3) * tn;
    // two tiles in flight per warp, or one when the group has a warp for every tile
    const int per = (W > 1 && ntile <= W) ? 1 : 2;
    for (int tile = per * (lane >> 5); tile < ntile; tile += per * W) {
        const int i0 = (tile / tn) * 8, j0 = (tile % tn) * 8;
        const bool two = per == 2 && tile + 1 < ntile;
        const int i1 = two ? ((tile + 1) / tn) * 8 : i0, j1 = two ? ((tile + 1) % tn) * 8 : j0;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        const bool ra0 = i0 + g < M, rb0 = j0 + g < N, ra1 = i1 + g < M, rb1 = j1 + g < N;
        const double* a0 = A + (i0 + g) * as0;
        const double* b0 = Bm + (j0 + g) * bs1;
        const double* a1 = A + (i1 + g) * as0;
        const double* b1 = Bm + (j1 + g) * bs1;
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int k = k0 + t;
            const bool kk = k < K;
            const double x0 = (ra0 && kk) ? a0[k * as1] : 0.0, y0 = (rb0 && kk) ? b0[k * bs0] : 0.0;
            const double x1 = (ra1 && kk) ? a1[k * as1] : 0.0, y1 = (rb1 && kk) ? b1[k * bs0] : 0.0;
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c00), "+d"(c01) : "d"(x0), "d"(y0));
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c10), "+d"(c11) : "d"(x1), "d"(y1));
        }
        auto put = [&](int i, int j, double v) {
            if (i < M && j < N) {
                double* c = C + i * N + j;
                if (MODE == SET) *c = v;
                else if (MODE == ADD) *c += v;
                else *c -= v;
            }
        };
        put(i0 + g, j0 + 2 * t, c00);
        put(i0 + g, j0 + 2 * t + 1, c01);
        if (two) {
            put(i1 + g, j1 + 2 * t, c10);
            put(i1 + g, j1 + 2 * t + 1, c11);
        }
    }
    gsync<W>();
}

// C[M,N] (row-major) (op)= sum_k A(i,k) * B(k,j) with A(i,k) = A[i*as0 + k*as1], B(k,j) = B[k*bs0 + j*bs1]
// (One copy per MODE in the DMMA build: inlined at its fifty call sites the adjoint kernel was 460 KB of SASS and its top stall
// in the tile products was instruction fetch.)
#ifdef BKF_GEN_DMMA
#define BKF_MM_INLINE __noinline__
#else
#define BKF_MM_INLINE __forceinline__
#endif
template <int MODE, int W, typename R>
__device__ BKF_MM_INLINE void mm(R* __restrict__ C, const R* __restrict__ A, int as0, int as1,
                                   const R* __restrict__ Bm, int bs0, int bs1, int M, int N, int K, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
#ifdef BKF_GEN_DMMA
    if constexpr (sizeof(R) == 8) {
        if (M >= 8 && N >= 8 && K >= 8) {  // (smaller products: a DMMA tile would be mostly padding)
            mm_dmma<MODE, W>(reinterpret_cast<double*>(C), reinterpret_cast<const double*>(A), as0, as1,
                          reinterpret_cast<const double*>(Bm), bs0, bs1, M, N, K, L);
            return;
        }
    }
#endif
    for (int idx = lane; idx < M * N; idx += NT) {
        int i = idx / N, j = idx - i * N;
        const R* a = A + i * as0;
        const R* b = Bm + j * bs1;
        R acc = 0;
        for (int k = 0; k < K; ++k) acc = fma(a[k * as1], b[k * bs0], acc);
        if (MODE == SET) C[idx] = acc;
        else if (MODE == ADD) C[idx] += acc;
        else C[idx] -= acc;
    }
    gsync<W>();
}

// y[M] (op)= sum_k A(i,k) x[k]
template <int MODE, int W, typename R>
__device__ __forceinline__ void mv(R* __restrict__ y, const R* __restrict__ A, int as0, int as1,
                                   const R* __restrict__ x, int M, int K, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    for (int i = lane; i < M; i += NT) {
        R acc = 0;
        for (int k = 0; k < K; ++k) acc = fma(A[i * as0 + k * as1], x[k], acc);
        if (MODE == SET) y[i] = acc;
        else if (MODE == ADD) y[i] += acc;
        else y[i] -= acc;
    }
    gsync<W>();
}

template <typename R, int W>
__device__ __forceinline__ void wcopy(R* __restrict__ dst, const R* __restrict__ src, int n, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    for (int i = lane; i < n; i += NT) dst[i] = src[i];
    gsync<W>();
}
template <typename R, int W>
__device__ __forceinline__ void wzero(R* dst, int n, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    for (int i = lane; i < n; i += NT) dst[i] = 0;
    gsync<W>();
}

// D = 0.5 (X + X^T) [+ Add] [+ jit * I]      (utilities.py:57-59 followed by stabilize :50-54)
template <typename R, int W>
__device__ __forceinline__ void sym_into(R* __restrict__ D, const R* __restrict__ X, const R* __restrict__ Add,
                                         R jit, int m, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    for (int idx = lane; idx < m * m; idx += NT) {
        int i = idx / m, j = idx - i * m;
        R v = R(0.5) * (X[idx] + X[j * m + i]);
        if (Add) v += Add[idx];
        if (i == j) v += jit;
        D[idx] = v;
    }
    gsync<W>();
}

// In-place lower Cholesky of the k x k matrix A (reads the lower triangle), column by column on the first warp of the
// group: every lane forms the pivot, the lanes split the rows below it (the sums run in the order of the one-lane loop
// this replaces, so the factor has the same bits).  Returns false when a pivot is not positive (its entry is NaN then).
template <typename R, int W>
__device__ bool chol_lower(R* A, int k, Lane<W> L) {
    const int lane = L.v;
    if (lane < 32) {
        for (int j = 0; j < k; ++j) {
            R s = A[j * k + j];
            for (int q = 0; q < j; ++q) s -= A[j * k + q] * A[j * k + q];
            if (!(s > 0)) s = R(NAN);
            const R ljj = sqrt(s);
            __syncwarp();  // every lane has read the pivot before it is overwritten
            for (int i = j + 1 + lane; i < k; i += 32) {
                R t = A[i * k + j];
                for (int q = 0; q < j; ++q) t -= A[i * k + q] * A[j * k + q];
                A[i * k + j] = t / ljj;
            }
            if (lane == 0) A[j * k + j] = ljj;
            __syncwarp();
        }
    }
    gsync<W>();
    bool ok = true;
    for (int j = 0; j < k; ++j) ok = ok && !isnan(A[j * k + j]);
    return ok;
}

// x <- (L L^T)^{-1} x for one right-hand side held by ONE lane (x has stride xs).
template <typename R>
__device__ __forceinline__ void chol_solve_vec(const R* L, int k, R* x, int xs) {
    for (int i = 0; i < k; ++i) {
        R t = x[i * xs];
        for (int q = 0; q < i; ++q) t -= L[i * k + q] * x[q * xs];
        x[i * xs] = t / L[i * k + i];
    }
    for (int i = k - 1; i >= 0; --i) {
        R t = x[i * xs];
        for (int q = i + 1; q < k; ++q) t -= L[q * k + i] * x[q * xs];
        x[i * xs] = t / L[i * k + i];
    }
}

// ----------------------------------------------------------------------------------------------------
// shared-memory arena of one warp
// ----------------------------------------------------------------------------------------------------
template <typename R>
struct Arena {
    // parameters
    R *T, *Z, *RQR, *H, *c, *d;
    // carried prediction and per-step scratch
    R *a, *P, *w, *ym, *v, *Zm, *Hm, *PZT, *F, *Lf, *K, *KH, *Lm, *X, *Y, *Pf, *af, *wv;
    // univariate
    R *Kst, *Fst, *vst, *zfst, *Pu;
    // backward
    R *abar, *Pbar, *G, *gT, *gZ, *gH, *gc, *gd, *Gsum, *Kbar, *PZTbar, *Fbar, *Finv, *Lbar, *Zmbar, *Hmbar,
        *vbar, *afbar, *tmp_mp;
};

__host__ __device__ inline size_t arena_elems(int m, int p, int r, bool backward) {
    size_t mm_ = (size_t)m * m, mp = (size_t)m * p, pp = (size_t)p * p;
    size_t fwd = /*T,RQR,P,Lm,X,Y,Pf,Pu*/ 8 * mm_ + /*Z,Zm,PZT,K,KH,Kst*/ 6 * mp + /*H,Hm,F,Lf*/ 4 * pp +
                 /*c,a,af*/ 3 * (size_t)m + /*d,w,ym,v,wv,Fst,vst,zfst*/ 8 * (size_t)p;
    size_t bwd = /*Pbar,G,gT,Gsum,Lbar*/ 5 * mm_ + /*gZ,Kbar,PZTbar,Zmbar,tmp_mp*/ 5 * mp +
                 /*gH,Fbar,Finv,Hmbar*/ 4 * pp + /*abar,gc,afbar*/ 3 * (size_t)m + /*gd,vbar*/ 2 * (size_t)p;
    size_t tot = fwd + (backward ? bwd : 0);
    return (tot + 1) & ~(size_t)1;
}

// The adjoint's arena with a staging buffer for the next step's record [a | P] behind it, when that still fits.
__host__ __device__ inline size_t stage_elems(int m) { return ((size_t)m + (size_t)m * m + 1) & ~(size_t)1; }
template <typename R>
__host__ __device__ inline size_t bwd_arena_elems(int m, int p, int r) {
    const size_t base = arena_elems(m, p, r, true), with = base + stage_elems(m);
    return with * sizeof(R) <= 225 * 1024 ? with : base;
}

template <typename R>
__device__ void carve(Arena<R>& A, R* s, int m, int p, int r, bool backward) {
    size_t mm_ = (size_t)m * m, mp = (size_t)m * p, pp = (size_t)p * p;
    auto take = [&](size_t n) { R* q = s; s += n; return q; };
    A.T = take(mm_); A.RQR = take(mm_); A.P = take(mm_); A.Lm = take(mm_); A.X = take(mm_); A.Y = take(mm_);
    A.Pf = take(mm_); A.Pu = take(mm_);
    A.Z = take(mp); A.Zm = take(mp); A.PZT = take(mp); A.K = take(mp); A.KH = take(mp); A.Kst = take(mp);
    A.H = take(pp); A.Hm = take(pp); A.F = take(pp); A.Lf = take(pp);
    A.c = take(m); A.a = take(m); A.af = take(m);
    A.d = take(p); A.w = take(p); A.ym = take(p); A.v = take(p); A.wv = take(p); A.Fst = take(p);
    A.vst = take(p); A.zfst = take(p);
    if (backward) {
        A.Pbar = take(mm_); A.G = take(mm_); A.gT = take(mm_); A.Gsum = take(mm_); A.Lbar = take(mm_);
        A.gZ = take(mp); A.Kbar = take(mp); A.PZTbar = take(mp); A.Zmbar = take(mp); A.tmp_mp = take(mp);
        A.gH = take(pp); A.Fbar = take(pp); A.Finv = take(pp); A.Hmbar = take(pp);
        A.abar = take(m); A.gc = take(m); A.afbar = take(m);
        A.gd = take(p); A.vbar = take(p);
    }
}

// dst = sym(R Q R^T) for global-memory R [m,r], Q [r,r] (once per series; no r <= m assumption). tmp: m*m.
template <typename R, int W>
__device__ void rqr_sym(R* dst, R* tmp, const R* __restrict__ Rg, const R* __restrict__ Qg, int m, int r, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    for (int idx = lane; idx < m * m; idx += NT) {
        int i = idx / m, j = idx - i * m;
        R acc = 0;
        for (int k = 0; k < r; ++k) {
            R rq = 0;  // (R Q)[i,k]
            for (int l = 0; l < r; ++l) rq = fma(Rg[i * r + l], Qg[l * r + k], rq);
            acc = fma(rq, Rg[j * r + k], acc);
        }
        tmp[idx] = acc;
    }
    gsync<W>();
    sym_into<R>(dst, tmp, nullptr, R(0), m, L);
}

// Load T, Z, H, c, d and RQR = sym(R Q R^T) (kalman_filter.py:408, second term) for series b at step t.
// all = true loads everything (first step of a sweep); otherwise only the time-varying parameters are refreshed
// (kalman_filter.py:110-142: the scan hands matrix[t] of every sequence parameter to kalman_step).
template <typename R, int W>
__device__ void load_params(const KArgs<R>& g, Arena<R>& A, int b, int t, bool all, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    const int m = g.m, p = g.p, r = g.r, n = g.n;
    const unsigned tv = g.tv_mask;
    auto at = [&](const R* base, int which, size_t sz) {
        return param_ptr_t(base, g.shared_mask, tv, which, b, t, n, sz);
    };
    auto need = [&](int which) { return all || ((tv >> which) & 1u); };
    if (need(BKF_PARAM_T)) wcopy(A.T, at(g.T, BKF_PARAM_T, (size_t)m * m), m * m, L);
    if (need(BKF_PARAM_Z)) wcopy(A.Z, at(g.Z, BKF_PARAM_Z, (size_t)p * m), p * m, L);
    if (need(BKF_PARAM_H)) wcopy(A.H, at(g.H, BKF_PARAM_H, (size_t)p * p), p * p, L);
    if (need(BKF_PARAM_C)) wcopy(A.c, at(g.c, BKF_PARAM_C, (size_t)m), m, L);
    if (need(BKF_PARAM_D)) wcopy(A.d, at(g.d, BKF_PARAM_D, (size_t)p), p, L);
    if (need(BKF_PARAM_R) || need(BKF_PARAM_Q))
        rqr_sym(A.RQR, A.Y, at(g.Rm, BKF_PARAM_R, (size_t)m * r), at(g.Q, BKF_PARAM_Q, (size_t)r * r), m, r, L);
}

// a+ = T af + c ; P+ = sym(T Pf T^T) + RQR         (kalman_filter.py:407-408).  Leaves X = T Pf.
template <typename R, int W>
__device__ __forceinline__ void predict(Arena<R>& A, int m, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    mv<SET>(A.a, A.T, m, 1, A.af, m, m, L);
    for (int i = lane; i < m; i += NT) A.a[i] += A.c[i];
    mm<SET>(A.X, A.T, m, 1, A.Pf, m, 1, m, m, m, L);
    mm<SET>(A.Y, A.X, m, 1, A.T, 1, m, m, m, m, L);
    sym_into<R>(A.P, A.Y, A.RQR, R(0), m, L);
}

// ----------------------------------------------------------------------------------------------------
// StandardFilter.update + stabilize (kalman_filter.py:352-360, 594-617, 536).
// In: A.a, A.P, y_t.  Out: A.af, A.Pf (jittered), A.v, A.Zm, A.Hm, A.PZT, A.F, A.Lf, A.K, A.Lm, A.wv.
// Returns ll; *flag_out = all-missing; *ok = F positive definite.
// ----------------------------------------------------------------------------------------------------
// Right-hand sides of the update's solve with F (rows of P Z^T, v, and the identity when the adjoint wants F^{-1}) and
// whether their transposed copy fits the X..Y scratch (2 m^2 values; it does unless p is about m).
__host__ __device__ inline int solve_rhs(int m, int p, bool finv) { return m + 1 + (finv ? p : 0); }
__host__ __device__ inline bool solve_scratch_fits(int m, int p, bool finv) {
    return (size_t)p * (solve_rhs(m, p, finv) | 1) <= 2 * (size_t)m * m;
}

template <typename R, int W>
__device__ R std_update(const KArgs<R>& g, Arena<R>& A, const R* yt, R* obs_mu, R* obs_cov, bool* flag_out,
                        bool* ok, R* Finv_out, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    const int m = g.m, p = g.p;
    for (int o = lane; o < p; o += NT) {
        R yv = yt[o];
        bool miss = is_missing(yv, g.fill, true);
        A.w[o] = miss ? R(0) : R(1);
        A.ym[o] = miss ? R(0) : yv;
    }
    gsync<W>();
    R nobs = 0;
    for (int o = 0; o < p; ++o) nobs += A.w[o];
    const bool flag = (nobs == R(0));
    for (int idx = lane; idx < p * m; idx += NT) A.Zm[idx] = A.w[idx / m] * A.Z[idx];
    for (int idx = lane; idx < p * p; idx += NT) A.Hm[idx] = A.w[idx / p] * A.H[idx];
    gsync<W>();
    // y_hat = d + Zm a ; v = ym - y_hat
    for (int o = lane; o < p; o += NT) {
        R acc = 0;
        for (int j = 0; j < m; ++j) acc = fma(A.Zm[o * m + j], A.a[j], acc);
        R yh = A.d[o] + acc;
        A.v[o] = A.ym[o] - yh;
        if (obs_mu) obs_mu[o] = yh;
    }
    // PZT = P Zm^T ; F = Zm PZT + (Hm + eps I)
    mm<SET>(A.PZT, A.P, m, 1, A.Zm, 1, m, m, p, m, L);
    mm<SET>(A.F, A.Zm, m, 1, A.PZT, p, 1, p, p, m, L);
    for (int idx = lane; idx < p * p; idx += NT) {
        int o = idx / p, q = idx - o * p;
        R f = A.F[idx] + (A.Hm[idx] + (o == q ? g.jitter : R(0)));
        A.F[idx] = f;
        A.Lf[idx] = f;
        if (obs_cov) obs_cov[idx] = f;
    }
    gsync<W>();
    *ok = chol_lower(A.Lf, p, L);
    // K = PZT F^{-1} (row i of K solves F k = pzt_i, F symmetric) ; wv = F^{-1} v ; for the adjoint also Finv = F^{-1}.
    // One right-hand side per thread, all solved at once in a transposed scratch (X..Y, free here) whose odd leading
    // dimension puts the threads' columns into distinct banks (in place, row i of K sits p doubles from row i + 1: every
    // access of the substitution was a 16-way bank conflict at p = 16).
    const int nrhs = solve_rhs(m, p, Finv_out != nullptr), ld = nrhs | 1;
    if (solve_scratch_fits(m, p, Finv_out != nullptr)) {
        R* S = A.X;
        for (int idx = lane; idx < p * nrhs; idx += NT) {
            const int q = idx / nrhs, i = idx - q * nrhs;
            S[q * ld + i] = i < m ? A.PZT[i * p + q] : (i == m ? A.v[q] : (i - m - 1 == q ? R(1) : R(0)));
        }
        gsync<W>();
        for (int i = lane; i < nrhs; i += NT) chol_solve_vec(A.Lf, p, S + i, ld);
        gsync<W>();
        for (int idx = lane; idx < p * nrhs; idx += NT) {
            const int q = idx / nrhs, i = idx - q * nrhs;
            const R x = S[q * ld + i];
            if (i < m) A.K[i * p + q] = x;
            else if (i == m) A.wv[q] = x;
            else Finv_out[q * p + (i - m - 1)] = x;
        }
        gsync<W>();
    } else {
        for (int idx = lane; idx < m * p; idx += NT) A.K[idx] = A.PZT[idx];
        for (int o = lane; o < p; o += NT) A.wv[o] = A.v[o];
        gsync<W>();
        for (int i = lane; i < m + 1; i += NT) {
            if (i < m) chol_solve_vec(A.Lf, p, A.K + i * p, 1);
            else chol_solve_vec(A.Lf, p, A.wv, 1);
        }
        gsync<W>();
    }
    // Lm = I - K Zm ; af = a + K v
    mm<SET>(A.Lm, A.K, p, 1, A.Zm, m, 1, m, m, p, L);
    for (int idx = lane; idx < m * m; idx += NT) A.Lm[idx] = ((idx / m) == (idx % m) ? R(1) : R(0)) - A.Lm[idx];
    for (int i = lane; i < m; i += NT) {
        R acc = 0;
        for (int o = 0; o < p; ++o) acc = fma(A.K[i * p + o], A.v[o], acc);
        A.af[i] = A.a[i] + acc;
    }
    gsync<W>();
    // Pf = sym(Lm P Lm^T) + sym(K Hm K^T) + eps I
    mm<SET>(A.X, A.Lm, m, 1, A.P, m, 1, m, m, m, L);
    mm<SET>(A.Y, A.X, m, 1, A.Lm, 1, m, m, m, m, L);
    mm<SET>(A.KH, A.K, p, 1, A.Hm, p, 1, m, p, p, L);
    mm<SET>(A.X, A.KH, p, 1, A.K, 1, p, m, m, p, L);
    for (int idx = lane; idx < m * m; idx += NT) {
        int i = idx / m, j = idx - i * m;
        R v1 = R(0.5) * (A.Y[idx] + A.Y[j * m + i]);
        R v2 = R(0.5) * (A.X[idx] + A.X[j * m + i]);
        A.Pf[idx] = (v1 + v2) + (i == j ? g.jitter : R(0));
    }
    gsync<W>();
    // ll = -0.5 (log 2pi + log det F + v' F^{-1} v)   (:609-615; log2pi counted once)
    R quad = 0, logdet = 0;
    for (int o = 0; o < p; ++o) {
        quad = fma(A.v[o], A.wv[o], quad);
        logdet += log(A.Lf[o * p + o]);
    }
    logdet = (p == 1) ? log(A.F[0]) : R(2) * logdet;
    *flag_out = flag;
    return flag ? R(0) : R(-0.5) * (R(kLog2Pi) + logdet + quad);
}

// ----------------------------------------------------------------------------------------------------
// UnivariateFilter inner loop + symmetrise/stabilise (kalman_filter.py:769-815).
// In: A.a, A.P.  Out: A.af, A.Pf (sym + jitter), A.Pu (P before symmetrisation), A.Kst/Fst/vst/zfst, A.w.
// ----------------------------------------------------------------------------------------------------
template <typename R, int W>
__device__ R uni_update(const KArgs<R>& g, Arena<R>& A, const R* yt, R* obs_mu, R* obs_cov, const bool psym, Lane<W> L) {
    const int lane = L.v;
    constexpr int NT = 32 * W;
    (void)lane; (void)NT;
    const int m = g.m, p = g.p;
    for (int o = lane; o < p; o += NT) {
        R yv = yt[o];
        bool miss = isnan(yv);  // :792 -- the fill value is ignored here
        A.w[o] = miss ? R(0) : R(1);
        A.ym[o] = miss ? R(0) : yv;
    }
    wcopy(A.af, A.a, m, L);
    wcopy(A.Pu, A.P, m * m, L);
    R llsum = 0;
    int nnz = 0;
    for (int o = 0; o < p; ++o) {
        const R wo = A.w[o];
        const R* z = A.Z + o * m;
        R yh = 0;
        for (int j = 0; j < m; ++j) yh = fma(wo * z[j], A.af[j], yh);
        yh += A.d[o];
        const R v = A.ym[o] - yh;
        // pz = P z^T  (kept in K row o), F0 = z pz + H_oo
        // psym: P is bitwise symmetric (every prediction is, sym_into; the rank-one updates below keep it so), and lane i
        // reads column i instead of row i -- the same values in the same order, without the bank conflicts of rows that
        // lie m doubles apart (16-way at m = 16, 24, 32, 40).  Only P0 (step 0) may be asymmetric and is read by rows.
        const int s_row = psym ? 1 : m, s_col = psym ? m : 1;
        for (int i = lane; i < m; i += NT) {
            R acc = 0;
            for (int j = 0; j < m; ++j) acc = fma(A.Pu[i * s_row + j * s_col], wo * z[j], acc);
            A.Kst[o * m + i] = acc;
        }
        gsync<W>();
        R F0 = 0;
        for (int j = 0; j < m; ++j) F0 = fma(wo * z[j], A.Kst[o * m + j], F0);
        F0 += wo * A.H[o * p + o];
        const bool zf = (F0 == R(0)) || (wo == R(0));
        const R F = F0 + g.jitter;
        const R keep = zf ? R(0) : R(1);
        gsync<W>();
        for (int i = lane; i < m; i += NT) A.Kst[o * m + i] = A.Kst[o * m + i] / F * keep;
        gsync<W>();
        for (int i = lane; i < m; i += NT) A.af[i] = fma(A.Kst[o * m + i], v, A.af[i]);
        for (int idx = lane; idx < m * m; idx += NT) {
            int i = idx / m, j = idx - i * m;
            A.Pu[idx] -= (A.Kst[o * m + i] * A.Kst[o * m + j]) * F;
        }
        const R lli = zf ? R(0) : log(F) + v * v / F;
        llsum += lli;
        nnz += (lli != R(0));
        if (lane == 0) {
            A.Fst[o] = F;
            A.vst[o] = v;
            A.zfst[o] = keep;
            if (o == p - 1) {
                if (obs_mu) obs_mu[0] = yh;
                if (obs_cov) obs_cov[0] = F;
            }
        }
        gsync<W>();
    }
    sym_into<R>(A.Pf, A.Pu, nullptr, g.jitter, m, L);
    return R(-0.5) * (R(nnz) * R(kLog2Pi) + llsum);
}

// ----------------------------------------------------------------------------------------------------
// forward kernel: standard / univariate
// ----------------------------------------------------------------------------------------------------
template <typename R, int W>
__global__ void __launch_bounds__(W == 1 ? 128 : 32 * W) forward_kernel(KArgs<R> g, int wpc, size_t arena) {
    extern __shared__ double smem_raw[];
    R* smem = reinterpret_cast<R*>(smem_raw);
    constexpr int NT = 32 * W;
    const int warp = W == 1 ? threadIdx.x >> 5 : 0, lane = W == 1 ? threadIdx.x & 31 : threadIdx.x;
    const Lane<W> L{lane};
    const int b = W == 1 ? blockIdx.x * wpc + warp : blockIdx.x;
    if (b >= g.B) return;
    const int m = g.m, p = g.p, n = g.n;
    const int po = (g.kind == BKF_UNIVARIATE) ? 1 : p;
#ifdef BKF_GEN_DMMA
    Arena<R> A;
    carve(A, smem + (size_t)warp * arena, m, p, g.r, false);
#else
    // Small shapes (this build serves m < 16): the arena's pointer table lives in shared memory.  Its some 45 pointers cost
    // 90 registers otherwise; measured +10 to 15 % at (8, 2, 2) and (4, 2, 2).  The DMMA build keeps them in registers: with
    // large arenas the shared memory bounds the occupancy, not the registers, and the re-loads after every barrier cost
    // the univariate loops 20 to 45 % (tools/bench_generic_groups.py).
    __shared__ Arena<R> arenas[4];
    Arena<R>& A = arenas[warp];
    if (lane == 0) carve(A, smem + (size_t)warp * arena, m, p, g.r, false);
    gsync<W>();
#endif
    wcopy(A.a, param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)m), m, L);
    wcopy(A.P, param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)m * m), m * m, L);
    const R* yb = g.y_shared ? g.y : g.y + (size_t)b * n * p;
    const size_t tstride = (size_t)m + (size_t)m * m;
    R logp = 0;
    int st = 0;
    for (int t = 0; t < n; ++t) {
        const size_t bt = (size_t)b * n + t;
        if (t == 0 || g.tv_mask) load_params(g, A, b, t, t == 0, L);
        if (g.traj) {
            wcopy(g.traj + bt * tstride, A.a, m, L);
            wcopy(g.traj + bt * tstride + m, A.P, m * m, L);
        }
        if (g.pred_a) wcopy(g.pred_a + bt * m, A.a, m, L);
        if (g.pred_P) wcopy(g.pred_P + bt * m * m, A.P, m * m, L);
        R* om = g.obs_mu ? g.obs_mu + bt * po : nullptr;
        R* oc = g.obs_cov ? g.obs_cov + bt * po * po : nullptr;
        R ll;
        if (g.kind == BKF_STANDARD) {
            bool flag, ok;
            ll = std_update(g, A, yb + (size_t)t * p, om, oc, &flag, &ok, (R*)nullptr, L);
            if (!ok) st |= BKF_STATUS_NOT_PD;
        } else {
            ll = uni_update(g, A, yb + (size_t)t * p, om, oc, t > 0, L);
        }
        if (g.filt_a) wcopy(g.filt_a + bt * m, A.af, m, L);
        if (g.filt_P) wcopy(g.filt_P + bt * m * m, A.Pf, m * m, L);
        if (g.ll && lane == 0) g.ll[bt] = ll;
        logp += ll;
        predict(A, m, L);
    }
    if (!isfinite(logp)) st |= BKF_STATUS_NONFINITE;
    if (lane == 0) {
        if (g.logp) g.logp[b] = logp;
        if (g.status) g.status[b] = st;
    }
}

// ----------------------------------------------------------------------------------------------------
// backward kernel (standard / univariate): reverse sweep over the stored predictions
// ----------------------------------------------------------------------------------------------------
template <typename R, int W>
// (minimum CTAs per SM: keeps the two- and four-warp builds at 168 registers = 12 warps per SM and the one-warp build for
//  small shapes at 128, whatever ptxas would pick; measured, tools/bench_generic_groups.py / bench_multivariate.py)
#ifdef BKF_GEN_DMMA
#define BKF_BWD_MINB(W) ((W) == 1 ? 2 : ((W) == 2 ? 6 : ((W) == 4 ? 3 : 1)))
#else
#define BKF_BWD_MINB(W) 4
#endif
__global__ void __launch_bounds__(W == 1 ? 128 : 32 * W, BKF_BWD_MINB(W))
backward_kernel(KArgs<R> g, int wpc, size_t arena) {
    extern __shared__ double smem_raw[];
    R* smem = reinterpret_cast<R*>(smem_raw);
    constexpr int NT = 32 * W;
    const int warp = W == 1 ? threadIdx.x >> 5 : 0, lane = W == 1 ? threadIdx.x & 31 : threadIdx.x;
    const Lane<W> L{lane};
    const int b = W == 1 ? blockIdx.x * wpc + warp : blockIdx.x;
    if (b >= g.B) return;
    const int m = g.m, p = g.p, r = g.r, n = g.n;
#ifdef BKF_GEN_DMMA
    Arena<R> A;
    carve(A, smem + (size_t)warp * arena, m, p, r, true);
#else
    // Small shapes (this build serves m < 16): the arena's pointer table lives in shared memory.  Its some 45 pointers cost
    // 90 registers otherwise; measured +10 to 15 % at (8, 2, 2) and (4, 2, 2).  The DMMA build keeps them in registers: with
    // large arenas the shared memory bounds the occupancy, not the registers, and the re-loads after every barrier cost
    // the univariate loops 20 to 45 % (tools/bench_generic_groups.py).
    __shared__ Arena<R> arenas[4];
    Arena<R>& A = arenas[warp];
    if (lane == 0) carve(A, smem + (size_t)warp * arena, m, p, r, true);
    gsync<W>();
#endif
    const unsigned tv = g.tv_mask;
    const bool tvRQ = ((tv >> BKF_PARAM_R) | (tv >> BKF_PARAM_Q)) & 1u;
    auto is_tv = [&](int which) { return ((tv >> which) & 1u) != 0; };
    // gradient of a time-varying parameter has a time axis: element offset of (b, t) in units of one parameter
    auto grec = [&](int which, int t) { return is_tv(which) ? (size_t)b * n + t : (size_t)b; };
    if (tvRQ) {  // accumulated in place (one lane owns each element; the sweep over t is sequential)
        if (g.g_R && !is_tv(BKF_PARAM_R)) wzero(g.g_R + (size_t)b * m * r, m * r, L);
        if (g.g_Q && !is_tv(BKF_PARAM_Q)) wzero(g.g_Q + (size_t)b * r * r, r * r, L);
    }
    wzero(A.abar, m, L); wzero(A.Pbar, m * m, L); wzero(A.gT, m * m, L); wzero(A.Gsum, m * m, L);
    wzero(A.gZ, p * m, L); wzero(A.gH, p * p, L); wzero(A.gc, m, L); wzero(A.gd, p, L);
    const R* yb = g.y_shared ? g.y : g.y + (size_t)b * n * p;
    const size_t tstride = (size_t)m + (size_t)m * m;
    // With room behind the arena the record [a | P] of step t - 1 travels by cp.async while step t is worked on (the
    // synchronous copy was 18 % of the univariate adjoint's samples at m = 27: global-memory latency with nothing to hide it).
    R* const stage = arena > arena_elems(m, p, r, true) ? smem + (size_t)warp * arena + arena_elems(m, p, r, true) : nullptr;
    auto fetch = [&](int t) {
        const R* src = g.traj + ((size_t)b * n + t) * tstride;
        for (int i = lane; i < (int)tstride; i += NT) cp_async_elem(stage + i, src + i);
    };
    if (stage) fetch(n - 1);
    for (int t = n - 1; t >= 0; --t) {
        const size_t bt = (size_t)b * n + t;
        const R gcot = g.ll_cot ? g.ll_cot[bt] : R(1);
        if (t == n - 1 || tv) load_params(g, A, b, t, t == n - 1, L);
        if (stage) {  // the record was fetched during the previous step; start on the next one
            cp_async_wait_all();
            gsync<W>();
            for (int i = lane; i < m; i += NT) A.a[i] = stage[i];
            for (int i = lane; i < m * m; i += NT) A.P[i] = stage[m + i];
            gsync<W>();
            if (t > 0) fetch(t - 1);
        } else {
            wcopy(A.a, g.traj + bt * tstride, m, L);
            wcopy(A.P, g.traj + bt * tstride + m, m * m, L);
        }
        bool flag = false, ok = true;
        if (g.kind == BKF_STANDARD) std_update(g, A, yb + (size_t)t * p, (R*)nullptr, (R*)nullptr, &flag, &ok, A.Finv, L);
        else uni_update(g, A, yb + (size_t)t * p, (R*)nullptr, (R*)nullptr, t > 0, L);
        // ---------------- predict adjoint ----------------
        // G = S(Pbar+); Gsum += G; gT += G T (Pf + Pf^T) + abar+ af^T; Pfbar = T^T G T; afbar = T^T abar+; gc += abar+
        sym_into<R>(A.G, A.Pbar, nullptr, R(0), m, L);
        for (int idx = lane; idx < m * m; idx += NT) {
            A.Gsum[idx] += A.G[idx];
            A.Y[idx] = A.Pf[idx] + A.Pf[(idx % m) * m + idx / m];
        }
        gsync<W>();
        if (tvRQ) {
            // R or Q varies in time: Rbar_t = G_t R_t (Q_t + Q_t^T), Qbar_t = R_t^T G_t R_t, per step
            const R* Rg = param_ptr_t(g.Rm, g.shared_mask, tv, BKF_PARAM_R, b, t, n, (size_t)m * r);
            const R* Qg = param_ptr_t(g.Q, g.shared_mask, tv, BKF_PARAM_Q, b, t, n, (size_t)r * r);
            if (g.g_R) {
                R* dst = g.g_R + grec(BKF_PARAM_R, t) * m * r;
                for (int idx = lane; idx < m * r; idx += NT) {
                    int i = idx / r, j = idx - i * r;
                    R acc = 0;
                    for (int k = 0; k < r; ++k) {
                        R gr = 0;
                        for (int l = 0; l < m; ++l) gr = fma(A.G[i * m + l], Rg[l * r + k], gr);
                        acc = fma(gr, Qg[k * r + j] + Qg[j * r + k], acc);
                    }
                    dst[idx] = is_tv(BKF_PARAM_R) ? acc : dst[idx] + acc;
                }
            }
            if (g.g_Q) {
                R* dst = g.g_Q + grec(BKF_PARAM_Q, t) * r * r;
                for (int idx = lane; idx < r * r; idx += NT) {
                    int i = idx / r, j = idx - i * r;
                    R acc = 0;
                    for (int k = 0; k < m; ++k) {
                        R gr = 0;
                        for (int l = 0; l < m; ++l) gr = fma(A.G[k * m + l], Rg[l * r + j], gr);
                        acc = fma(Rg[k * r + i], gr, acc);
                    }
                    dst[idx] = is_tv(BKF_PARAM_Q) ? acc : dst[idx] + acc;
                }
            }
            gsync<W>();
        }
        mm<SET>(A.X, A.T, m, 1, A.Y, m, 1, m, m, m, L);
        mm<ADD>(A.gT, A.G, m, 1, A.X, m, 1, m, m, m, L);
        for (int idx = lane; idx < m * m; idx += NT) A.gT[idx] = fma(A.abar[idx / m], A.af[idx % m], A.gT[idx]);
        mm<SET>(A.X, A.G, m, 1, A.T, m, 1, m, m, m, L);
        mm<SET>(A.Y, A.T, 1, m, A.X, m, 1, m, m, m, L);  // Y = Pfbar
        mv<SET>(A.afbar, A.T, 1, m, A.abar, m, m, L);
        for (int i = lane; i < m; i += NT) A.gc[i] += A.abar[i];
        gsync<W>();
        if (g.kind == BKF_STANDARD) {
            // ---------------- Joseph-form adjoint ----------------
            sym_into<R>(A.G, A.Y, nullptr, R(0), m, L);  // G2
            for (int idx = lane; idx < m * m; idx += NT) A.Y[idx] = A.P[idx] + A.P[(idx % m) * m + idx / m];
            gsync<W>();
            mm<SET>(A.X, A.Lm, m, 1, A.Y, m, 1, m, m, m, L);
            mm<SET>(A.Lbar, A.G, m, 1, A.X, m, 1, m, m, m, L);       // Lbar = G2 Lm (P + P^T)
            mm<SET>(A.X, A.G, m, 1, A.Lm, m, 1, m, m, m, L);
            mm<SET>(A.Pbar, A.Lm, 1, m, A.X, m, 1, m, m, m, L);       // Pbar = Lm^T G2 Lm
            for (int idx = lane; idx < p * p; idx += NT) A.Fbar[idx] = A.Hm[idx] + A.Hm[(idx % p) * p + idx / p];
            gsync<W>();
            mm<SET>(A.tmp_mp, A.K, p, 1, A.Fbar, p, 1, m, p, p, L);   // K (Hm + Hm^T)
            mm<SET>(A.Kbar, A.G, m, 1, A.tmp_mp, p, 1, m, p, m, L);   // Kbar = G2 K (Hm + Hm^T)
            mm<SET>(A.tmp_mp, A.G, m, 1, A.K, p, 1, m, p, m, L);      // G2 K
            mm<SET>(A.Hmbar, A.K, 1, p, A.tmp_mp, p, 1, p, p, m, L);  // Hmbar = K^T G2 K
            // a_f = a + K v
            wcopy(A.abar, A.afbar, m, L);
            for (int idx = lane; idx < m * p; idx += NT) A.Kbar[idx] = fma(A.afbar[idx / p], A.v[idx % p], A.Kbar[idx]);
            mv<SET>(A.vbar, A.K, 1, p, A.afbar, p, m, L);
            // Lm = I - K Zm
            mm<SUB>(A.Kbar, A.Lbar, m, 1, A.Zm, 1, m, m, p, m, L);    // Kbar -= Lbar Zm^T
            mm<SET>(A.Zmbar, A.K, 1, p, A.Lbar, m, 1, p, m, m, L);    // K^T Lbar
            for (int idx = lane; idx < p * m; idx += NT) A.Zmbar[idx] = -A.Zmbar[idx];
            // Finv = F^{-1} (already there when std_update solved for it beside K)
            if (!solve_scratch_fits(m, p, true)) {
                for (int idx = lane; idx < p * p; idx += NT) A.Finv[idx] = ((idx / p) == (idx % p)) ? R(1) : R(0);
                gsync<W>();
                for (int q = lane; q < p; q += NT) chol_solve_vec(A.Lf, p, A.Finv + q, p);
            }
            gsync<W>();
            // ll adjoint
            for (int idx = lane; idx < p * p; idx += NT) {
                int o = idx / p, q = idx - o * p;
                A.Fbar[idx] = flag ? R(0) : R(-0.5) * gcot * (A.Finv[q * p + o] - A.wv[o] * A.wv[q]);
            }
            if (!flag)
                for (int o = lane; o < p; o += NT) A.vbar[o] -= gcot * A.wv[o];
            gsync<W>();
            // K = PZT F^{-1}
            mm<SET>(A.PZTbar, A.Kbar, p, 1, A.Finv, 1, p, m, p, p, L);  // Kbar F^{-T}
            mm<SUB>(A.Fbar, A.K, 1, p, A.PZTbar, p, 1, p, p, m, L);     // Fbar -= K^T PZTbar
            // F = Zm PZT + Hm + eps I
            mm<ADD>(A.Zmbar, A.Fbar, p, 1, A.PZT, 1, p, p, m, p, L);    // Zmbar += Fbar PZT^T
            mm<ADD>(A.PZTbar, A.Zm, 1, m, A.Fbar, p, 1, m, p, p, L);    // PZTbar += Zm^T Fbar
            for (int idx = lane; idx < p * p; idx += NT) A.Hmbar[idx] += A.Fbar[idx];
            // PZT = P Zm^T
            mm<ADD>(A.Pbar, A.PZTbar, p, 1, A.Zm, m, 1, m, m, p, L);    // Pbar += PZTbar Zm
            mm<ADD>(A.Zmbar, A.PZTbar, 1, p, A.P, m, 1, p, m, m, L);    // Zmbar += PZTbar^T P
            // v = ym - d - Zm a
            for (int o = lane; o < p; o += NT) A.gd[o] -= A.vbar[o];
            for (int idx = lane; idx < p * m; idx += NT) A.Zmbar[idx] = fma(-A.vbar[idx / m], A.a[idx % m], A.Zmbar[idx]);
            mv<SUB>(A.abar, A.Zm, 1, m, A.vbar, m, p, L);               // abar -= Zm^T vbar
            // mask
            for (int idx = lane; idx < p * m; idx += NT) A.gZ[idx] = fma(A.w[idx / m], A.Zmbar[idx], A.gZ[idx]);
            for (int idx = lane; idx < p * p; idx += NT) A.gH[idx] = fma(A.w[idx / p], A.Hmbar[idx], A.gH[idx]);
            gsync<W>();
        } else {
            // ---------------- univariate adjoint ----------------
            // Pf = 0.5 (Pu + Pu^T) + eps I
            sym_into<R>(A.Pbar, A.Y, nullptr, R(0), m, L);
            wcopy(A.abar, A.afbar, m, L);
            const R llbar = R(-0.5) * gcot;
            for (int o = p - 1; o >= 0; --o) {
                const R keep = A.zfst[o];
                if (keep == R(0)) continue;  // K = 0: a, P pass through; no parameter gradient
                const R F = A.Fst[o], v = A.vst[o], wo = A.w[o];
                const R* K = A.Kst + o * m;
                const R* z = A.Z + o * m;  // wo == 1 here
                // state before this inner update
                for (int idx = lane; idx < m * m; idx += NT) A.Pu[idx] += (K[idx / m] * K[idx % m]) * F;
                for (int i = lane; i < m; i += NT) A.af[i] = fma(-K[i], v, A.af[i]);
                gsync<W>();
                // Kbar_i = -sum_j (Pbar_ij + Pbar_ji) K_j F + abar_i v   (tmp in X[0:m])
                R kpk = 0, ka = 0, kk = 0;
                if constexpr (W > 1) {
                    // (Pbar K)_i and (Pbar^T K)_i: warp w sums its share of the j, lane i of every warp holds row / column i;
                    // the partial sums meet in the scratch (X is free in this branch; gsum3 uses its first 3 W values)
                    R* part = A.X + 32;
                    const int wq = lane >> 5;
                    for (int i = lane & 31; i < m; i += 32) {
                        R rs = 0, cs = 0;
                        for (int j = wq; j < m; j += W) {
                            rs = fma(A.Pbar[i * m + j], K[j], rs);
                            cs = fma(A.Pbar[j * m + i], K[j], cs);
                        }
                        part[(2 * wq) * m + i] = rs;
                        part[(2 * wq + 1) * m + i] = cs;
                    }
                    gsync<W>();
                }
                for (int i = lane; i < m; i += NT) {
                    R acc = 0, pk = 0;
                    if constexpr (W > 1) {
                        R cs = 0;
                        for (int w = 0; w < W; ++w) { pk += A.X[32 + (2 * w) * m + i]; cs += A.X[32 + (2 * w + 1) * m + i]; }
                        acc = pk + cs;
                    } else {
                        for (int j = 0; j < m; ++j) acc = fma(A.Pbar[i * m + j] + A.Pbar[j * m + i], K[j], acc);
                        for (int j = 0; j < m; ++j) pk = fma(A.Pbar[i * m + j], K[j], pk);
                    }
                    R kb = fma(-acc, F, A.abar[i] * v);
                    A.Kbar[i] = kb;
                    kpk = fma(K[i], pk, kpk);
                    ka = fma(K[i], A.abar[i], ka);
                    kk = fma(kb, K[i], kk);
                }
                gsum3<W>(kpk, ka, kk, A.X, lane);  // (X is free in this branch: scratch of the CTA-wide sum)
                R Fbar = -kpk + llbar * (R(1) / F - v * v / (F * F));
                const R vbar = ka + llbar * R(2) * v / F;
                Fbar -= kk / F;
                gsync<W>();
                // pzbar = Kbar / F + Fbar z ; zbar = Fbar pz + P^T pzbar - vbar a ; (pz = K F)
                for (int i = lane; i < m; i += NT) A.PZTbar[i] = A.Kbar[i] / F + Fbar * z[i];
                gsync<W>();
                const R* pzbar = A.PZTbar;
                for (int j = lane; j < m; j += NT) {
                    R acc = 0;
                    for (int i = 0; i < m; ++i) acc = fma(A.Pu[i * m + j], pzbar[i], acc);
                    R zb = fma(Fbar, K[j] * F, acc);
                    zb = fma(-vbar, A.af[j], zb);
                    A.gZ[o * m + j] = fma(wo, zb, A.gZ[o * m + j]);
                    A.abar[j] = fma(-vbar, z[j], A.abar[j]);
                }
                for (int idx = lane; idx < m * m; idx += NT) A.Pbar[idx] = fma(pzbar[idx / m], z[idx % m], A.Pbar[idx]);
                if (lane == 0) {
                    A.gH[o * p + o] = fma(wo, Fbar, A.gH[o * p + o]);
                    A.gd[o] -= vbar;
                }
                gsync<W>();
            }
        }
        if (tv) {  // time-varying parameters: this step's contribution IS the gradient of x[t]
            auto flush = [&](int which, R* dst, R* acc, size_t sz) {
                if (!is_tv(which)) return;
                if (dst) wcopy(dst + ((size_t)b * n + t) * sz, acc, (int)sz, L);
                wzero(acc, (int)sz, L);
            };
            flush(BKF_PARAM_T, g.g_T, A.gT, (size_t)m * m);
            flush(BKF_PARAM_C, g.g_c, A.gc, m);
            flush(BKF_PARAM_Z, g.g_Z, A.gZ, (size_t)p * m);
            flush(BKF_PARAM_H, g.g_H, A.gH, (size_t)p * p);
            flush(BKF_PARAM_D, g.g_d, A.gd, p);
        }
    }
    // ---------------- write gradients ----------------
    auto put = [&](int which, R* dst, const R* src, size_t sz) {
        if (dst && !is_tv(which)) wcopy(dst + (size_t)b * sz, src, (int)sz, L);
    };
    put(BKF_PARAM_A0, g.g_a0, A.abar, m); put(BKF_PARAM_P0, g.g_P0, A.Pbar, (size_t)m * m);
    put(BKF_PARAM_C, g.g_c, A.gc, m); put(BKF_PARAM_D, g.g_d, A.gd, p);
    put(BKF_PARAM_T, g.g_T, A.gT, (size_t)m * m); put(BKF_PARAM_Z, g.g_Z, A.gZ, (size_t)p * m);
    put(BKF_PARAM_H, g.g_H, A.gH, (size_t)p * p);
    if (tvRQ) return;  // Rbar, Qbar were produced step by step
    // Rbar = Gsum R (Q + Q^T) ; Qbar = R^T Gsum R      (sym(R Q R^T) is time-invariant: sum the cotangents first)
    const R* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)m * r);
    const R* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    if (g.g_R) {
        for (int idx = lane; idx < m * r; idx += NT) {
            int i = idx / r, j = idx - i * r;
            R acc = 0;
            for (int k = 0; k < r; ++k) {
                R gr = 0;  // (Gsum R)[i,k]
                for (int l = 0; l < m; ++l) gr = fma(A.Gsum[i * m + l], Rg[l * r + k], gr);
                acc = fma(gr, Qg[k * r + j] + Qg[j * r + k], acc);
            }
            g.g_R[(size_t)b * m * r + idx] = acc;
        }
    }
    if (g.g_Q) {
        for (int idx = lane; idx < r * r; idx += NT) {
            int i = idx / r, j = idx - i * r;
            R acc = 0;
            for (int k = 0; k < m; ++k) {
                R gr = 0;  // (Gsum R)[k,j]
                for (int l = 0; l < m; ++l) gr = fma(A.Gsum[k * m + l], Rg[l * r + j], gr);
                acc = fma(Rg[k * r + i], gr, acc);
            }
            g.g_Q[(size_t)b * r * r + idx] = acc;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// smoother kernel (kalman_smoother.py:66-126)
// ----------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t smooth_arena_elems(int m, int r) {
    size_t tot = /*T,RQR,P,Ps,Phat,X,Y,V,G*/ 9 * (size_t)m * m + /*a,as,ahat,lam*/ 4 * (size_t)m;
    return (tot + 1) & ~(size_t)1;
}

// Symmetric eigen-decomposition by cyclic Jacobi: A (destroyed; diag -> eigenvalues), V (eigenvectors in columns).
template <typename R>
__device__ void jacobi_eig(R* A, R* V, int m, int lane) {
    constexpr int W = 1, NT = 32;  // the eigen-decomposition runs on one warp
    for (int idx = lane; idx < m * m; idx += NT) V[idx] = ((idx / m) == (idx % m)) ? R(1) : R(0);
    gsync<W>();
    const R tol2 = sizeof(R) == 8 ? R(1e-34) : R(1e-15);
    for (int sweep = 0; sweep < 30; ++sweep) {
        R off = 0, dg = 0;
        for (int idx = lane; idx < m * m; idx += NT) {
            R x = A[idx];
            if ((idx / m) == (idx % m)) dg = fma(x, x, dg); else off = fma(x, x, off);
        }
        off = warp_sum(off); dg = warp_sum(dg);
        if (off <= tol2 * dg) break;
        for (int pi = 0; pi < m - 1; ++pi) {
            for (int qi = pi + 1; qi < m; ++qi) {
                const R apq = A[pi * m + qi];
                const R app = A[pi * m + pi], aqq = A[qi * m + qi];
                gsync<W>();
                if (apq == R(0)) continue;
                const R theta = (aqq - app) / (R(2) * apq);
                const R tt = (theta >= R(0) ? R(1) : R(-1)) / (fabs(theta) + sqrt(fma(theta, theta, R(1))));
                const R cs = R(1) / sqrt(fma(tt, tt, R(1))), sn = tt * cs;
                for (int k = lane; k < m; k += NT) {
                    if (k != pi && k != qi) {
                        R akp = A[k * m + pi], akq = A[k * m + qi];
                        R np_ = cs * akp - sn * akq, nq = sn * akp + cs * akq;
                        A[k * m + pi] = np_; A[pi * m + k] = np_;
                        A[k * m + qi] = nq; A[qi * m + k] = nq;
                    }
                    R vkp = V[k * m + pi], vkq = V[k * m + qi];
                    V[k * m + pi] = cs * vkp - sn * vkq;
                    V[k * m + qi] = sn * vkp + cs * vkq;
                }
                if (lane == 0) {
                    A[pi * m + pi] = app - tt * apq;
                    A[qi * m + qi] = aqq + tt * apq;
                    A[pi * m + qi] = 0; A[qi * m + pi] = 0;
                }
                gsync<W>();
            }
        }
    }
}

// One series of the smoother sweep by one warp or, W > 1, by a CTA of W warps (s: the series' arena).
template <typename R, int W>
__device__ void smooth_series(const KArgs<R>& g, const int b, R* s, const R jitter0, const bool force_pinv, const int lane) {
    constexpr int NT = 32 * W;
    const Lane<W> L{lane};
    const int m = g.m, r = g.r, n = g.n;
    const size_t mm_ = (size_t)m * m;
    auto take = [&](size_t k) { R* q = s; s += k; return q; };
    R *T = take(mm_), *RQR = take(mm_), *P = take(mm_), *Ps = take(mm_), *Ph = take(mm_), *X = take(mm_),
      *Y = take(mm_), *V = take(mm_), *G = take(mm_), *a = take(m), *as = take(m), *ah = take(m), *lam = take(m);
    // Time-varying T, R, Q: the reference's scan runs go_backwards over sequences of unequal length
    // (filtered[:-1] has n-1 rows, T has n): PyTensor reverses each sequence and then truncates to the shortest, so
    // the smoother step of filtered index t sees T[t+1], R[t+1], Q[t+1] (kalman_smoother.py:84-92; DESIGN quirk Q9).
    const unsigned tv = g.tv_mask;
    auto load_trq = [&](int t, bool all) {
        if (all || ((tv >> BKF_PARAM_T) & 1u))
            wcopy(T, param_ptr_t(g.T, g.shared_mask, tv, BKF_PARAM_T, b, t, n, mm_), m * m, L);
        if (all || (((tv >> BKF_PARAM_R) | (tv >> BKF_PARAM_Q)) & 1u))
            rqr_sym(RQR, Y, param_ptr_t(g.Rm, g.shared_mask, tv, BKF_PARAM_R, b, t, n, (size_t)m * r),
                    param_ptr_t(g.Q, g.shared_mask, tv, BKF_PARAM_Q, b, t, n, (size_t)r * r), m, r, L);
    };
    load_trq(n - 1, true);
    const R* fa = g.sm_filt_a + (size_t)b * n * m;
    const R* fP = g.sm_filt_P + (size_t)b * n * mm_;
    R* oa = g.sm_a ? g.sm_a + (size_t)b * n * m : nullptr;
    R* oP = g.sm_P ? g.sm_P + (size_t)b * n * mm_ : nullptr;
    wcopy(as, fa + (size_t)(n - 1) * m, m, L);
    wcopy(Ps, fP + (size_t)(n - 1) * mm_, m * m, L);
    if (oa) wcopy(oa + (size_t)(n - 1) * m, as, m, L);
    if (oP) wcopy(oP + (size_t)(n - 1) * mm_, Ps, m * m, L);
    int st = 0;
    for (int t = n - 2; t >= 0; --t) {
        if (tv && t != n - 2) load_trq(t + 1, false);
        wcopy(a, fa + (size_t)t * m, m, L);
        wcopy(P, fP + (size_t)t * mm_, m * m, L);
        mv<SET>(ah, T, m, 1, a, m, m, L);  // no state intercept (:122, quirk Q8)
        mm<SET>(X, T, m, 1, P, m, 1, m, m, m, L);
        mm<SET>(Y, X, m, 1, T, 1, m, m, m, m, L);
        sym_into<R>(Ph, Y, RQR, g.jitter, m, L);
        // pinv(P_hat) T P (:112).  P_hat is symmetric positive definite but for degenerate models, and then its
        // pseudo-inverse IS its inverse: float64 steps whose Cholesky pivots are positive and span less than 1e13 (the
        // test of the specialised smoothers, bkf_mma_smooth.cuh) solve P_hat W = T P by substitution, one column per lane
        // -- some 10 k cycles at m = 27 where the eigen-decomposition below takes several 100 k.  Everything else (a
        // failed test, float32, the series a specialised smoother handed over: force_pinv) takes the eigen-decomposition.
        bool direct = false;
        if (sizeof(R) == 8 && !force_pinv) {
            wcopy(V, Ph, m * m, L);
            const bool ok = chol_lower(V, m, L);
            R dmin = V[0], dmax = V[0];
            for (int i = 1; i < m; ++i) { dmin = fmin(dmin, V[i * m + i]); dmax = fmax(dmax, V[i * m + i]); }
            direct = ok && !(dmin * dmin < R(1e-13) * (dmax * dmax));
            if (direct) {
                for (int c = lane; c < m; c += NT) chol_solve_vec(V, m, X + c, m);  // X = T P  ->  P_hat^{-1} T P
                gsync<W>();
            }
        }
        if (!direct) {
        // symmetric eigen-decomposition, singular values below 1e-15 * max dropped
        wcopy(Y, Ph, m * m, L);
        if (lane < 32) jacobi_eig(Y, V, m, lane);  // (one warp; the rest of a CTA-per-series group waits)
        gsync<W>();
        R lmax = 0;
        for (int i = 0; i < m; ++i) lmax = fmax(lmax, fabs(Y[i * m + i]));
        const R cutoff = R(1e-15) * lmax;
        for (int i = lane; i < m; i += NT) {
            R l = Y[i * m + i];
            lam[i] = fabs(l) > cutoff ? R(1) / l : R(0);
        }
        gsync<W>();
        for (int idx = lane; idx < m * m; idx += NT) {  // Y = V diag(lam) V^T
            int i = idx / m, j = idx - i * m;
            R acc = 0;
            for (int k = 0; k < m; ++k) acc = fma(V[i * m + k] * lam[k], V[j * m + k], acc);
            X[idx] = acc;
        }
        gsync<W>();
        // X = pinv T P
        mm<SET>(Y, X, m, 1, T, m, 1, m, m, m, L);
        mm<SET>(X, Y, m, 1, P, m, 1, m, m, m, L);
        }
        // G = (pinv T P)^T
        for (int idx = lane; idx < m * m; idx += NT) G[idx] = X[(idx % m) * m + idx / m];
        for (int i = lane; i < m; i += NT) ah[i] = as[i] - ah[i];
        gsync<W>();
        for (int i = lane; i < m; i += NT) {
            R acc = 0;
            for (int k = 0; k < m; ++k) acc = fma(G[i * m + k], ah[k], acc);
            as[i] = a[i] + acc;
        }
        for (int idx = lane; idx < m * m; idx += NT) Y[idx] = Ps[idx] - Ph[idx];
        gsync<W>();
        mm<SET>(X, G, m, 1, Y, m, 1, m, m, m, L);
        mm<SET>(Y, X, m, 1, G, 1, m, m, m, m, L);
        for (int idx = lane; idx < m * m; idx += NT) {
            int i = idx / m, j = idx - i * m;
            R val = P[idx] + R(0.5) * (Y[idx] + Y[j * m + i]);
            if (i == j) val = (val + g.jitter) + jitter0;  // double jitter (:116-117, quirk Q8)
            Ps[idx] = val;
        }
        gsync<W>();
        if (oa) wcopy(oa + (size_t)t * m, as, m, L);
        if (oP) wcopy(oP + (size_t)t * mm_, Ps, m * m, L);
    }
    if (lane == 0 && g.status) g.status[b] |= st;
}

// list == nullptr: warp w of the grid takes series w.  Otherwise the series are list[0 .. *count) (the ones a fast
// smoother flagged as ill-conditioned, bkf_api.cu::run_smoother): the grid strides over them and a launch with
// *count == 0 returns at once -- no host round trip decides whether this kernel has work.
template <typename R, int W>
__global__ void __launch_bounds__(W == 1 ? 128 : 32 * W)
smoother_kernel(KArgs<R> g, int wpc, size_t arena, R jitter0, const int* list, const int* count, bool force_pinv) {
    extern __shared__ double smem_raw[];
    R* s = reinterpret_cast<R*>(smem_raw);
    if constexpr (W > 1) {  // one series per CTA (never with a list: the launcher keeps W = 1 there)
        if ((int)blockIdx.x < g.B) smooth_series<R, W>(g, blockIdx.x, s, jitter0, force_pinv, threadIdx.x);
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        s += (size_t)warp * arena;
        if (!list) {
            const int b = blockIdx.x * wpc + warp;
            if (b < g.B) smooth_series<R, 1>(g, b, s, jitter0, force_pinv, lane);
            return;
        }
        const int cnt = *count;
        for (int i = blockIdx.x * wpc + warp; i < cnt; i += gridDim.x * wpc) {
            smooth_series<R, 1>(g, list[i], s, jitter0, true, lane);
            __syncwarp();
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------------
static int pick_wpc(size_t arena_bytes, int B) {
    const size_t budget = 200 * 1024;
    int wpc = 4;
    while (wpc > 1 && arena_bytes * wpc > budget) wpc >>= 1;
    (void)B;
    return wpc;
}

// W == 1: wpc series per CTA, one warp each.  W > 1: one series per CTA of W warps.
template <typename R, int W, typename KernelT>
static int launch_group_kernel(KernelT kern, const KArgs<R>& g, size_t arena, cudaStream_t stream, const char** err) {
    size_t arena_bytes = arena * sizeof(R);
    if (arena_bytes > 225 * 1024) {  // (227 KB less the static pointer tables)
        *err = "problem too large for the generic shared-memory kernel (m, p, r)";
        return BKF_ERR_UNSUPPORTED;
    }
    int wpc = W == 1 ? pick_wpc(arena_bytes, g.B) : 1;
    size_t smem = arena_bytes * wpc;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    int grid = (g.B + wpc - 1) / wpc;
    kern<<<grid, W == 1 ? 32 * wpc : 32 * W, smem, stream>>>(g, wpc, arena);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

template <typename R, int W>
int launch_forward_w(const KArgs<R>& g, cudaStream_t stream, const char** err) {
    return launch_group_kernel<R, W>(forward_kernel<R, W>, g, arena_elems(g.m, g.p, g.r, false), stream, err);
}
template <typename R, int W>
int launch_backward_w(const KArgs<R>& g, cudaStream_t stream, const char** err) {
    return launch_group_kernel<R, W>(backward_kernel<R, W>, g, bwd_arena_elems<R>(g.m, g.p, g.r), stream, err);
}

// Series of this build that one SM holds at a time (registers, shared memory and the CTA limit, from the runtime's
// occupancy calculator); W == 1 packs wpc series into a CTA.
enum { K_FWD = 0, K_BWD = 1, K_SMOOTH = 2 };
template <typename R, int W>
int resident_series_w(int which, size_t arena_bytes) {
    const int wpc = W == 1 ? pick_wpc(arena_bytes, 0) : 1;
    const size_t smem = arena_bytes * wpc;
    const int threads = W == 1 ? 32 * wpc : 32 * W;
    int ctas = 0;
    auto query = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, threads, smem);
    };
    const cudaError_t e = which == K_FWD ? query(forward_kernel<R, W>)
                        : which == K_BWD ? query(backward_kernel<R, W>) : query(smoother_kernel<R, W>);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return ctas * wpc;
}

// The smoother sweep of one build (list / count: the W = 1 build only).
template <typename R, int W>
int launch_smoother_w(const KArgs<R>& g, const int* list, const int* count, cudaStream_t stream, const char** err) {
    size_t arena = smooth_arena_elems(g.m, g.r);
    size_t arena_bytes = arena * sizeof(R);
    if (arena_bytes > 227 * 1024) { *err = "problem too large for the generic smoother kernel"; return BKF_ERR_UNSUPPORTED; }
    int wpc = W == 1 ? pick_wpc(arena_bytes, g.B) : 1;
    size_t smem = arena_bytes * wpc;
    auto kern = smoother_kernel<R, W>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    int grid = (g.B + wpc - 1) / wpc;
    if (list && grid > 148) grid = 148;  // flagged series only: a small strided grid (usually there are none)
    R jitter0 = sizeof(R) == 8 ? R(1e-8) : R(1e-6);  // JITTER_DEFAULT, utils/constants.py:20
    // BKF_GEN_SMOOTH_PINV=1: the eigen-decomposition pinv at every step (what this kernel did before it tried a Cholesky solve)
    const char* env = getenv("BKF_GEN_SMOOTH_PINV");
    const bool force_pinv = env && env[0] == '1';
    kern<<<grid, W == 1 ? 32 * wpc : 32 * W, smem, stream>>>(g, wpc, arena, jitter0, list, count, force_pinv);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

#if BKF_GEN_W == 1
#ifdef BKF_GEN_DMMA
// The CTA-per-series builds live in their own translation units (bkf_generic_tc_w{2,4,8}.cu).
extern template int launch_forward_w<double, 2>(const KArgs<double>&, cudaStream_t, const char**);
extern template int launch_forward_w<double, 4>(const KArgs<double>&, cudaStream_t, const char**);
extern template int launch_forward_w<double, 8>(const KArgs<double>&, cudaStream_t, const char**);
extern template int launch_backward_w<double, 2>(const KArgs<double>&, cudaStream_t, const char**);
extern template int launch_backward_w<double, 4>(const KArgs<double>&, cudaStream_t, const char**);
extern template int launch_backward_w<double, 8>(const KArgs<double>&, cudaStream_t, const char**);
extern template int resident_series_w<double, 2>(int, size_t);
extern template int resident_series_w<double, 4>(int, size_t);
extern template int resident_series_w<double, 8>(int, size_t);
extern template int launch_smoother_w<double, 2>(const KArgs<double>&, const int*, const int*, cudaStream_t, const char**);
extern template int launch_smoother_w<double, 4>(const KArgs<double>&, const int*, const int*, cudaStream_t, const char**);
extern template int launch_smoother_w<double, 8>(const KArgs<double>&, const int*, const int*, cudaStream_t, const char**);

// Warps per series.  Measured (tools/bench_generic_groups.py): a series runs about sqrt(W) times faster on W warps (the
// dense products scale, the Cholesky, the substitutions and the barriers do not), so the launcher takes the width with the
// largest (series resident per SM) * sqrt(W); the resident series come from the occupancy calculator, i.e. from the
// registers the build at hand really has.  With small arenas that is the one-warp mapping, at the VARMAX shape (one series
// per SM) it is eight warps.  BKF_GEN_W=<1|2|4|8> in the environment overrides the choice.
static int pick_w(size_t arena_bytes, int which) {
    if (const char* e = getenv("BKF_GEN_W")) {
        int w = atoi(e);
        if (w == 1 || w == 2 || w == 4 || w == 8) return w;
    }
    if (arena_bytes > 225 * 1024) return 1;  // (the launch reports the error)
    static std::mutex mu;
    static std::map<std::pair<size_t, int>, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({arena_bytes, which});
    if (it != cache.end()) return it->second;
    const int series[4] = {resident_series_w<double, 1>(which, arena_bytes), resident_series_w<double, 2>(which, arena_bytes),
                           resident_series_w<double, 4>(which, arena_bytes), resident_series_w<double, 8>(which, arena_bytes)};
    const double gain[4] = {1.0, 1.41, 2.0, 2.83};
    double best = series[0] * gain[0];
    int w_best = 1;
    for (int k = 1; k < 4; ++k)
        if (series[k] * gain[k] > best * 1.05) { best = series[k] * gain[k]; w_best = 1 << k; }
    cache[{arena_bytes, which}] = w_best;
    return w_best;
}
int width(const KArgs<double>& g, bool backward) {
    return pick_w((backward ? bwd_arena_elems<double>(g.m, g.p, g.r) : arena_elems(g.m, g.p, g.r, false)) * sizeof(double),
                  backward ? K_BWD : K_FWD);
}
#endif

template <typename R>
int launch_forward(const KArgs<R>& g, cudaStream_t stream, const char** err) {
#ifdef BKF_GEN_DMMA
    switch (width(g, false)) {
        case 2: return launch_forward_w<R, 2>(g, stream, err);
        case 4: return launch_forward_w<R, 4>(g, stream, err);
        case 8: return launch_forward_w<R, 8>(g, stream, err);
        default: break;
    }
#endif
    return launch_forward_w<R, 1>(g, stream, err);
}

template <typename R>
int launch_backward(const KArgs<R>& g, cudaStream_t stream, const char** err) {
#ifdef BKF_GEN_DMMA
    switch (width(g, true)) {
        case 2: return launch_backward_w<R, 2>(g, stream, err);
        case 4: return launch_backward_w<R, 4>(g, stream, err);
        case 8: return launch_backward_w<R, 8>(g, stream, err);
        default: break;
    }
#endif
    return launch_backward_w<R, 1>(g, stream, err);
}

template <typename R>
int launch_smoother(const KArgs<R>& g, const int* list, const int* count, cudaStream_t stream, const char** err) {
#ifdef BKF_GEN_DMMA
    if (!list) {  // (the flagged series of a specialised smoother are few: one warp each)
        switch (pick_w(smooth_arena_elems(g.m, g.r) * sizeof(R), K_SMOOTH)) {
            case 2: return launch_smoother_w<R, 2>(g, list, count, stream, err);
            case 4: return launch_smoother_w<R, 4>(g, list, count, stream, err);
            case 8: return launch_smoother_w<R, 8>(g, list, count, stream, err);
            default: break;
        }
    }
#endif
    return launch_smoother_w<R, 1>(g, list, count, stream, err);
}

template int launch_forward<double>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_backward<double>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_smoother<double>(const KArgs<double>&, const int*, const int*, cudaStream_t, const char**);
#ifndef BKF_GEN_DMMA
template int launch_forward<float>(const KArgs<float>&, cudaStream_t, const char**);
template int launch_backward<float>(const KArgs<float>&, cudaStream_t, const char**);
template int launch_smoother<float>(const KArgs<float>&, const int*, const int*, cudaStream_t, const char**);
#endif
#else   // BKF_GEN_W > 1: this unit holds one CTA-per-series build of the two sweeps
template int launch_forward_w<double, BKF_GEN_W>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_backward_w<double, BKF_GEN_W>(const KArgs<double>&, cudaStream_t, const char**);
template int resident_series_w<double, BKF_GEN_W>(int, size_t);
template int launch_smoother_w<double, BKF_GEN_W>(const KArgs<double>&, const int*, const int*, cudaStream_t, const char**);
#endif

}  // namespace BKF_GEN_NS
}  // namespace bkf
