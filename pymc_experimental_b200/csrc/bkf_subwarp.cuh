// bkf_subwarp.cuh -- register-tiled "sub-warp per series" kernels, p = 1, float64; instantiated for 5 <= m <= 8
// (9 <= m <= 16 moved to the FP64 tensor-core family, bkf_mma.cuh).
//
// Mapping (DESIGN.md "Kernels"): LPS = 8 or 16 lanes own one time-sequential filter; lane j keeps COLUMN j of
// every m x m state matrix (P, P_f, Pbar, ...) in registers.  A dense product C = A B is evaluated as
//     C[:, j] = sum_k A[:, k] * B[k, j]
// with B[k, j] a register of lane j and A[:, k] a 128-bit *broadcast* shared-memory read, so one LDS.128 feeds
// two DFMAs of every lane and the FP64 pipe -- not shared memory -- is the limiter.  Results that are needed as
// the A operand of the next product are written back to shared memory column by column.  For t >= 1 every
// covariance is exactly symmetric (0.5 (X + X^T) is), so "row j" == "column j" and no transposes are needed in
// the O(m^2) update; the (possibly asymmetric) user-supplied P0 at t = 0 goes through the ASYM instantiation.
//
// Reference formulas: kalman_filter.py:529-539 (step), :594-617 (StandardFilter.update, p = 1),
// :769-820 (UnivariateFilter, p = 1), :407-408 (predict); adjoint: SURVEY.md Appendix B specialised to p = 1.
#pragma once
#include "bkf_common.cuh"

namespace bkf {
namespace sw {

enum { K_STD = 0, K_UNI = 1 };

template <int M>
struct Cfg {
    static constexpr int LPS = (M <= 8) ? 8 : 16;  // lanes per series
    static constexpr int SPW = 32 / LPS;           // series per warp
    static constexpr int LD = LPS + 2;             // leading dim of column-major smem matrices (10 / 18):
                                                   // even (16-byte aligned columns) and LD % 16 in {2, 10} makes
                                                   // per-lane 128-bit column accesses bank-conflict free
    static constexpr int MAT = LD * LPS;           // doubles per smem matrix
    static constexpr int DENSE = M + M * M;        // doubles of one trajectory record: a (M) then P (M x M)
    static constexpr int DENSE_PAD = (DENSE + 1) & ~1;
    static constexpr int VEC = LPS;                // doubles per smem vector
};

__device__ __forceinline__ double sub_sum(double v, int lps) {
    if (lps > 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// x[i] = S[k*LD + i], i < M: broadcast read of column k (all lanes of the series use the same address)
template <int M>
__device__ __forceinline__ void load_vec(const double* __restrict__ S, double (&x)[M]) {
    const double2* c = reinterpret_cast<const double2*>(S);
#pragma unroll
    for (int q = 0; q < M / 2; ++q) {
        double2 d = c[q];
        x[2 * q] = d.x;
        x[2 * q + 1] = d.y;
    }
    if (M & 1) x[M - 1] = S[M - 1];
}
template <int M>
__device__ __forceinline__ void store_vec(double* __restrict__ S, const double (&x)[M]) {
    double2* c = reinterpret_cast<double2*>(S);
#pragma unroll
    for (int q = 0; q < M / 2; ++q) c[q] = make_double2(x[2 * q], x[2 * q + 1]);
    if (M & 1) S[M - 1] = x[M - 1];
}

// acc[:] = sum_k As[:, k] * b[k]   (As column-major in smem with leading dim LD, b in registers)
template <int M, int LD>
__device__ __forceinline__ void prod(const double* __restrict__ As, const double (&b)[M], double (&acc)[M]) {
#pragma unroll
    for (int k = 0; k < M; ++k) {
        double col[M];
        load_vec<M>(As + k * LD, col);
#pragma unroll
        for (int i = 0; i < M; ++i) acc[i] = (k == 0) ? col[i] * b[0] : fma(col[i], b[k], acc[i]);
    }
}

template <int M>
__device__ __forceinline__ double dot(const double (&x)[M], const double (&y)[M]) {
    double s = x[0] * y[0];
#pragma unroll
    for (int i = 1; i < M; ++i) s = fma(x[i], y[i], s);
    return s;
}

// Per-series shared-memory layouts ------------------------------------------------------------------
template <int M>
struct FwdSmem {
    using C = Cfg<M>;
    static constexpr int T_cm = 0;                    // T column-major: S[k*LD + i] = T[i,k]
    static constexpr int X = T_cm + C::MAT;           // T Pf (columns)
    static constexpr int Y = X + C::MAT;              // X T^T (columns)
    static constexpr int RQR = Y + C::MAT;            // sym(R Q R^T) (columns)
    static constexpr int OUT = RQR + C::MAT;          // dense trajectory record being assembled
    static constexpr int VK = OUT + C::DENSE_PAD;     // K
    static constexpr int VA = VK + C::VEC;            // a_f
    static constexpr int VZ = VA + C::VEC;            // Z row (unmasked)
    static constexpr int VP = VZ + C::VEC;            // pzt (standard)
    static constexpr int VQ = VP + C::VEC;            // zp (standard, t = 0)
    // series blocks are offset by 128/SPW bytes (mod 128) so that the broadcast reads of the SPW series of one
    // warp fall into different shared-memory banks
    static constexpr int TOTAL = ((VQ + C::VEC + 15) & ~15) + 16 / C::SPW;
};

template <int M>
struct BwdSmem {
    using C = Cfg<M>;
    static constexpr int T_rm = 0;                    // T row-major: S[k*LD + i] = T[k,i]  (== T^T column-major)
    static constexpr int GW = T_rm + C::MAT;          // Pbar -> G -> W = G T (columns)
    static constexpr int GSUM = GW + C::MAT;          // sum_t G_t (columns)
    static constexpr int GT = GSUM + C::MAT;          // dL/dT (columns)
    static constexpr int IN0 = GT + C::MAT;           // dense trajectory record, double buffered
    static constexpr int IN1 = IN0 + C::DENSE_PAD;
    static constexpr int VK = IN1 + C::DENSE_PAD;     // K
    static constexpr int VA = VK + C::VEC;            // abar+
    static constexpr int VZ = VA + C::VEC;            // Z row (unmasked)
    static constexpr int VP = VZ + C::VEC;            // scratch vectors
    static constexpr int VQ = VP + C::VEC;
    static constexpr int VR = VQ + C::VEC;
    static constexpr int TOTAL = ((VR + C::VEC + 15) & ~15) + 16 / C::SPW;
};

// Quantities of one measurement update that the predict step / the adjoint need.
template <int M>
struct Upd {
    double Pf[M];   // column j of P_f (after + eps I)
    double kv[M];   // K (all entries)
    double K, pz, zp, F, f0, v, af, keep, ll, zj, yhat;
    bool flag;
};

// ---------------------------------------------------------------------------------------------------
// Measurement update, p = 1.  Pc = column j of P, Pr = row j of P (aliases Pc when !ASYM).
// Vectors K (and pzt, zp for the standard filter) are exchanged through smem (vK, vP, vQ); zv = unmasked Z.
// ---------------------------------------------------------------------------------------------------
template <int M, int KIND, bool ASYM>
__device__ __forceinline__ void update(Upd<M>& u, const double (&Pc)[M], const double (&Pr)[M], double a_j, int j,
                                       double yv, double fill, double jitter, double Zj, const double (&zv)[M],
                                       double d0, double H0, double* vK, double* vP, double* vQ) {
    constexpr int LPS = Cfg<M>::LPS;
    const bool miss = (KIND == K_UNI) ? isnan(yv) : (isnan(yv) || yv == fill);  // :792 vs :352 (quirk Q6)
    const double wobs = miss ? 0.0 : 1.0;
    const double ym = miss ? 0.0 : yv;
    const double zj = wobs * Zj;
    const double Hm = wobs * H0;
    u.zj = zj;
    u.flag = miss;
    const double yhat = d0 + sub_sum(zj * a_j, LPS);  // d + Z a (:594) == Z a + d (:770)
    u.yhat = yhat;
    u.v = ym - yhat;
    // pz_j = sum_i P[j,i] z_i (row j);  zp_j = sum_i z_i P[i,j] (column j)
    double pz = wobs * dot<M>(Pr, zv);
    u.pz = pz;
    u.zp = ASYM ? wobs * dot<M>(Pc, zv) : pz;
    const double f0 = sub_sum(zj * pz, LPS);
    u.f0 = f0;
    if (KIND == K_UNI) {
        const double F0 = f0 + Hm;
        const bool zf = (F0 == 0.0) || miss;  // :777
        const double F = F0 + jitter;
        u.keep = zf ? 0.0 : 1.0;
        u.F = F;
        u.K = pz / F * u.keep;  // :782
        u.af = fma(u.K, u.v, a_j);
        const double lli = zf ? 0.0 : log(F) + u.v * u.v / F;  // :787
        u.ll = -0.5 * ((lli != 0.0 ? kLog2Pi : 0.0) + lli);    // :818
        if (j < M) vK[j] = u.K;
        __syncwarp();
        load_vec<M>(vK, u.kv);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            double pu = Pc[i] - (u.kv[i] * u.K) * F;  // :785
            if (ASYM) {
                double put = Pr[i] - (u.K * u.kv[i]) * F;
                pu = 0.5 * (pu + put);  // :815 (exact no-op when P is symmetric)
            }
            u.Pf[i] = (i == j) ? pu + jitter : pu;
        }
    } else {
        const double F = f0 + (Hm + jitter);  // :598
        u.F = F;
        u.keep = 1.0;
        u.K = pz / F;  // solve(F^T, PZT^T)^T, p = 1  (:600)
        u.af = fma(u.K, u.v, a_j);
        u.ll = miss ? 0.0 : -0.5 * (kLog2Pi + log(F) + u.v * (u.v / F));  // :611-615
        if (j < M) {
            vK[j] = u.K;
            vP[j] = pz;
            if (ASYM) vQ[j] = u.zp;
        }
        __syncwarp();
        double pv[M];
        load_vec<M>(vK, u.kv);
        load_vec<M>(vP, pv);
        double qv[M];
        if (ASYM) load_vec<M>(vQ, qv);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            // Joseph form with L = I - K z implicit:  Y = (P - K (zP)) - ((P z^T) - K f0) K^T     (:601-604)
            const double zp_i = ASYM ? qv[i] : pv[i];
            const double y_ij = fma(-(fma(-u.kv[i], f0, pv[i])), u.K, fma(-u.kv[i], u.zp, Pc[i]));
            const double y_ji = fma(-(fma(-u.K, f0, pz)), u.kv[i], fma(-u.K, zp_i, Pr[i]));
            const double khk = 0.5 * ((u.kv[i] * Hm) * u.K + (u.K * Hm) * u.kv[i]);
            const double pf = 0.5 * (y_ij + y_ji) + khk;
            u.Pf[i] = (i == j) ? pf + jitter : pf;  // stabilize, :536
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------
template <int M, int KIND>
__global__ void __launch_bounds__(64) forward_kernel(KArgs<double> g) {
    using C = Cfg<M>;
    using S = FwdSmem<M>;
    constexpr int LPS = C::LPS, LD = C::LD;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPS, j = lane % LPS;
    const int series = (blockIdx.x * (blockDim.x >> 5) + warp) * C::SPW + sub;
    const bool valid = series < g.B;
    const int b = valid ? series : g.B - 1;
    double* sm = smem + (size_t)((warp * C::SPW + sub)) * S::TOTAL;
    const int n = g.n;
    const bool act = j < M;

    // ---- per-series constants ----
    const double* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)M * M);
    const double* Zg = param_ptr(g.Z, g.shared_mask, BKF_PARAM_Z, b, (size_t)M);
    const double* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * g.r);
    const double* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)g.r * g.r);
    const double H0 = *param_ptr(g.H, g.shared_mask, BKF_PARAM_H, b, 1);
    const double d0 = *param_ptr(g.d, g.shared_mask, BKF_PARAM_D, b, 1);
    const double* cg = param_ptr(g.c, g.shared_mask, BKF_PARAM_C, b, (size_t)M);
    const double* a0g = param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)M);
    const double* P0g = param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)M * M);
    double Trow[M];  // row j of T
#pragma unroll
    for (int k = 0; k < M; ++k) Trow[k] = act ? Tg[j * M + k] : 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) sm[S::T_cm + k * LD + j] = Trow[k];  // T[j,k] -> column k, row j
    const double Zj = act ? Zg[j] : 0.0;
    sm[S::VZ + j] = Zj;
    const double cj = act ? cg[j] : 0.0;
    {  // RQR = sym(R Q R^T): lane j computes column j (and row j) once per series
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
    __syncwarp();
    double zv[M];
    load_vec<M>(sm + S::VZ, zv);

    // ---- state ----
    double Pc[M], Pr[M];
    double a_j = act ? a0g[j] : 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        Pc[i] = act ? P0g[i * M + j] : 0.0;
        Pr[i] = act ? P0g[j * M + i] : 0.0;
    }
    const double* yb = g.y_shared ? g.y : g.y + (size_t)b * n;
    const size_t rec = (size_t)C::DENSE;
    double logp = 0;
    Upd<M> u;
    for (int t = 0; t < n; ++t) {
        const size_t bt = (size_t)b * n + t;
        const double yv = yb[t];
        // ---- emit the prediction (a, P) of step t: trajectory for the adjoint and/or user outputs ----
        if (g.traj || g.pred_a || g.pred_P) {
            double* o = sm + S::OUT;
            if (act) {
                o[j] = a_j;
#pragma unroll
                for (int i = 0; i < M; ++i) o[M + i * M + j] = Pc[i];  // row-major P
            }
            __syncwarp();
            if (valid) {
                if (g.traj) {
                    double2* dst = reinterpret_cast<double2*>(g.traj + bt * rec);
                    const double2* src = reinterpret_cast<const double2*>(o);
                    for (int q = j; q < C::DENSE / 2; q += LPS) dst[q] = src[q];
                }
                if (g.pred_a && act) g.pred_a[bt * M + j] = a_j;
                if (g.pred_P)
                    for (int q = j; q < M * M; q += LPS) g.pred_P[bt * M * M + q] = o[M + q];
            }
            __syncwarp();
        }
        if (t == 0)
            update<M, KIND, true>(u, Pc, Pr, a_j, j, yv, g.fill, g.jitter, Zj, zv, d0, H0, sm + S::VK, sm + S::VP, sm + S::VQ);
        else
            update<M, KIND, false>(u, Pc, Pc, a_j, j, yv, g.fill, g.jitter, Zj, zv, d0, H0, sm + S::VK, sm + S::VP, sm + S::VQ);
        logp += u.ll;
        if (valid) {
            if (g.ll && j == 0) g.ll[bt] = u.ll;
            if (g.obs_mu && j == 0) g.obs_mu[bt] = u.yhat;
            if (g.obs_cov && j == 0) g.obs_cov[bt] = u.F;
            if (g.filt_a && act) g.filt_a[bt * M + j] = u.af;
            if (g.filt_P && act) {
#pragma unroll
                for (int i = 0; i < M; ++i) g.filt_P[bt * M * M + (size_t)j * M + i] = u.Pf[i];  // symmetric
            }
        }
        // ---- predict: a+ = T af + c ; P+ = sym(T Pf T^T) + RQR ----
        if (act) sm[S::VA + j] = u.af;
        double X[M];
        prod<M, LD>(sm + S::T_cm, u.Pf, X);
        store_vec<M>(sm + S::X + j * LD, X);
        __syncwarp();
        {
            double av[M];
            load_vec<M>(sm + S::VA, av);
            a_j = dot<M>(Trow, av) + cj;
        }
        double Y[M];
        prod<M, LD>(sm + S::X, Trow, Y);
        store_vec<M>(sm + S::Y + j * LD, Y);
        __syncwarp();
        {
            double rq[M];
            load_vec<M>(sm + S::RQR + j * LD, rq);
#pragma unroll
            for (int i = 0; i < M; ++i) Pc[i] = 0.5 * (Y[i] + sm[S::Y + i * LD + j]) + rq[i];
        }
        __syncwarp();
    }
    if (valid && j == 0) {
        if (g.logp) g.logp[b] = logp;
        if (g.status) g.status[b] = isfinite(logp) ? 0 : BKF_STATUS_NONFINITE;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward kernel
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int M, int KIND>
__global__ void __launch_bounds__(64) backward_kernel(KArgs<double> g) {
    using C = Cfg<M>;
    using S = BwdSmem<M>;
    constexpr int LPS = C::LPS, LD = C::LD;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPS, j = lane % LPS;
    const int series = (blockIdx.x * (blockDim.x >> 5) + warp) * C::SPW + sub;
    const bool valid = series < g.B;
    const int b = valid ? series : g.B - 1;
    double* sm = smem + (size_t)((warp * C::SPW + sub)) * S::TOTAL;
    const int n = g.n;
    const bool act = j < M;

    const double* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)M * M);
    const double* Zg = param_ptr(g.Z, g.shared_mask, BKF_PARAM_Z, b, (size_t)M);
    const double H0 = *param_ptr(g.H, g.shared_mask, BKF_PARAM_H, b, 1);
    const double d0 = *param_ptr(g.d, g.shared_mask, BKF_PARAM_D, b, 1);
    double Tcol[M];  // column j of T
#pragma unroll
    for (int k = 0; k < M; ++k) Tcol[k] = act ? Tg[k * M + j] : 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) sm[S::T_rm + k * LD + j] = Tcol[k];  // T[k,j] -> "column" k (row k of T), entry j
    const double Zj = act ? Zg[j] : 0.0;
    sm[S::VZ + j] = Zj;
#pragma unroll
    for (int i = 0; i < LPS; ++i) {
        sm[S::GSUM + j * LD + i] = 0.0;
        sm[S::GT + j * LD + i] = 0.0;
    }
    __syncwarp();
    double zv[M];
    load_vec<M>(sm + S::VZ, zv);

    const double* yb = g.y_shared ? g.y : g.y + (size_t)b * n;
    const size_t rec = (size_t)C::DENSE;
    auto prefetch = [&](int t, int buf) {
        const char* src = reinterpret_cast<const char*>(g.traj + ((size_t)b * n + t) * rec);
        char* dst = reinterpret_cast<char*>(sm + (buf ? S::IN1 : S::IN0));
        for (int q = j; q < C::DENSE / 2; q += LPS) cp_async16(dst + q * 16, src + q * 16);
        cp_async_commit();
    };
    double abar = 0.0, gz = 0.0, gc = 0.0, gH = 0.0, gd = 0.0;
    double Pbar[M];
#pragma unroll
    for (int i = 0; i < M; ++i) Pbar[i] = 0.0;
    prefetch(n - 1, (n - 1) & 1);
    Upd<M> u;
    for (int t = n - 1; t >= 0; --t) {
        const size_t bt = (size_t)b * n + t;
        const double gcot = g.ll_cot ? g.ll_cot[bt] : 1.0;
        const double yv = yb[t];
        cp_async_wait<0>();
        __syncwarp();
        if (t > 0) prefetch(t - 1, (t - 1) & 1);
        const double* in = sm + ((t & 1) ? S::IN1 : S::IN0);
        const double a_j = act ? in[j] : 0.0;
        double Pc[M];
        double S2c[M];  // column j of P + P^T (standard adjoint only)
        if (t == 0) {
            double Pr[M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                Pc[i] = act ? in[M + i * M + j] : 0.0;
                Pr[i] = act ? in[M + j * M + i] : 0.0;
            }
            update<M, KIND, true>(u, Pc, Pr, a_j, j, yv, g.fill, g.jitter, Zj, zv, d0, H0, sm + S::VK, sm + S::VP, sm + S::VQ);
#pragma unroll
            for (int i = 0; i < M; ++i) S2c[i] = Pc[i] + Pr[i];
        } else {
#pragma unroll
            for (int i = 0; i < M; ++i) Pc[i] = act ? in[M + j * M + i] : 0.0;
            update<M, KIND, false>(u, Pc, Pc, a_j, j, yv, g.fill, g.jitter, Zj, zv, d0, H0, sm + S::VK, sm + S::VP, sm + S::VQ);
#pragma unroll
            for (int i = 0; i < M; ++i) S2c[i] = 2.0 * Pc[i];
        }
        // ================= predict adjoint =================
        // G = S(Pbar+)
        store_vec<M>(sm + S::GW + j * LD, Pbar);
        if (act) sm[S::VA + j] = abar;
        __syncwarp();
        double Gc[M];
#pragma unroll
        for (int i = 0; i < M; ++i) Gc[i] = 0.5 * (Pbar[i] + sm[S::GW + i * LD + j]);
        __syncwarp();
        store_vec<M>(sm + S::GW + j * LD, Gc);
        {
            double gs[M];
            load_vec<M>(sm + S::GSUM + j * LD, gs);
#pragma unroll
            for (int i = 0; i < M; ++i) gs[i] += Gc[i];
            store_vec<M>(sm + S::GSUM + j * LD, gs);
        }
        __syncwarp();
        // W = G T
        double W[M];
        prod<M, LD>(sm + S::GW, Tcol, W);
        __syncwarp();
        store_vec<M>(sm + S::GW + j * LD, W);
        __syncwarp();
        // Pfbar = T^T W
        double Pfb[M];
        prod<M, LD>(sm + S::T_rm, W, Pfb);
        // gT += W (Pf + Pf^T) + abar+ af^T      (Pf is symmetric: 2 W Pf)
        double abv[M];
        load_vec<M>(sm + S::VA, abv);
        {
            double tmp[M], gt[M];
            prod<M, LD>(sm + S::GW, u.Pf, tmp);
            load_vec<M>(sm + S::GT + j * LD, gt);
#pragma unroll
            for (int i = 0; i < M; ++i) gt[i] += fma(abv[i], u.af, 2.0 * tmp[i]);
            store_vec<M>(sm + S::GT + j * LD, gt);
        }
        const double afbar = dot<M>(Tcol, abv);
        gc += abar;
        // ================= update adjoint =================
        if (KIND == K_UNI) {
            // Branch-free on purpose: the series sharing a warp may disagree on `keep` (missing data), and the
            // shuffles / __syncwarp below must be executed by all 32 lanes.  When keep == 0 (K = 0: a and P pass
            // through, no parameter gradient) the computed values are discarded by the selects.
            const bool keep = u.keep != 0.0;
            const double llbar = -0.5 * gcot;
            const double F = u.F, v = u.v;
            const double pk = dot<M>(Pfb, u.kv);  // (Pu_bar K)_j, Pu_bar = S(Pfbar) ~ Pfbar
            const double Kbar = fma(-2.0 * F, pk, afbar * v);
            const double kpk = sub_sum(u.K * pk, LPS);
            const double ka = sub_sum(u.K * afbar, LPS);
            const double kk = sub_sum(Kbar * u.K, LPS);
            const double Fbar = -kpk + llbar * (1.0 / F - v * v / (F * F)) - kk / F;
            const double vbar = ka + llbar * 2.0 * v / F;
            const double pzbar = Kbar / F + Fbar * u.zj;
            if (act) sm[S::VP + j] = pzbar;
            __syncwarp();
            double pzv[M];
            load_vec<M>(sm + S::VP, pzv);
            const double zbar = fma(Fbar, u.pz, dot<M>(Pc, pzv)) - vbar * a_j;
            abar = keep ? fma(-vbar, u.zj, afbar) : afbar;
#pragma unroll
            for (int i = 0; i < M; ++i) Pbar[i] = keep ? fma(pzv[i], u.zj, Pfb[i]) : Pfb[i];
            gz += keep ? zbar : 0.0;
            gH += keep ? Fbar : 0.0;
            gd -= keep ? vbar : 0.0;
            __syncwarp();
        } else {
            // G2 = S(Pfbar) ~ Pfbar (symmetric up to rounding): column j doubles as row j
            const double F = u.F, v = u.v, f0 = u.f0;
            const double Hm = u.flag ? 0.0 : H0;
            const double wv = v / F;
            // xz2 = (P + P^T) z - 2 f0 K ;  u1 = G2 xz2 ;  gk = G2 K
            const double xz2_j = (u.pz + u.zp) - 2.0 * f0 * u.K;
            if (act) {
                sm[S::VP + j] = xz2_j;
            }
            __syncwarp();
            double xv[M];
            load_vec<M>(sm + S::VP, xv);
            const double u1 = dot<M>(Pfb, xv);
            const double gk = dot<M>(Pfb, u.kv);
            const double kgk = sub_sum(u.K * gk, LPS);
            if (act) sm[S::VQ + j] = gk;
            __syncwarp();
            double gkv[M];
            load_vec<M>(sm + S::VQ, gkv);
            // Zmbar = -K^T Lbar = -(gk^T (P + P^T) - (gk.K) z^T (P + P^T))
            double zmbar = -(dot<M>(gkv, S2c) - kgk * (u.zp + u.pz));
            // Kbar = G2 K (Hm + Hm^T) + afbar v - Lbar z
            double Kbar = fma(2.0 * Hm, gk, afbar * v) - u1;
            double Hmbar = kgk;
            double vbar = sub_sum(u.K * afbar, LPS);
            double Fbar = 0.0;
            if (!u.flag) {
                Fbar = -0.5 * gcot * (1.0 / F - wv * wv);
                vbar -= gcot * wv;
            }
            // K = PZT / F
            double pztbar = Kbar / F;
            Fbar -= sub_sum(u.K * pztbar, LPS);
            // F = Zm PZT + Hm + eps
            zmbar = fma(Fbar, u.pz, zmbar);
            pztbar = fma(Fbar, u.zj, pztbar);
            Hmbar += Fbar;
            if (act) sm[S::VR + j] = pztbar;
            __syncwarp();
            double pbv[M];
            load_vec<M>(sm + S::VR, pbv);
            // PZT = P Zm^T
            zmbar += dot<M>(Pc, pbv);
            // v = ym - d - Zm a
            zmbar = fma(-vbar, a_j, zmbar);
            abar = fma(-vbar, u.zj, afbar);
            gd -= vbar;
            // Pbar = L^T G2 L + PZTbar Zm
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double pb = Pfb[i] - zv[i] * (u.flag ? 0.0 : 1.0) * gk - gkv[i] * u.zj;
                pb = fma(zv[i] * (u.flag ? 0.0 : 1.0) * u.zj, kgk, pb);
                Pbar[i] = fma(pbv[i], u.zj, pb);
            }
            if (!u.flag) {
                gz += zmbar;
                gH += Hmbar;
            }
            __syncwarp();
        }
    }
    // ---- gradients out ----
    if (valid) {
        if (g.g_a0 && act) g.g_a0[(size_t)b * M + j] = abar;
        if (g.g_c && act) g.g_c[(size_t)b * M + j] = gc;
        if (g.g_Z && act) g.g_Z[(size_t)b * M + j] = gz;
        if (g.g_d && j == 0) g.g_d[b] = gd;
        if (g.g_H && j == 0) g.g_H[b] = gH;
        if (g.g_P0 && act) {
#pragma unroll
            for (int i = 0; i < M; ++i) g.g_P0[(size_t)b * M * M + i * M + j] = Pbar[i];
        }
    }
    __syncwarp();
    if (valid && g.g_T) {
        for (int q = j; q < M * M; q += LPS) {
            int i = q / M, k = q - i * M;
            g.g_T[(size_t)b * M * M + q] = sm[S::GT + k * LD + i];
        }
    }
    // Rbar = Gsum R (Q + Q^T) ; Qbar = R^T Gsum R   (sym(R Q R^T) is time-invariant: cotangents summed first)
    const int r = g.r;
    const double* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * r);
    const double* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    if (valid && g.g_R) {
        for (int idx = j; idx < M * r; idx += LPS) {
            int i = idx / r, jj = idx - i * r;
            double acc = 0;
            for (int k = 0; k < r; ++k) {
                double gr = 0;
                for (int l = 0; l < M; ++l) gr = fma(sm[S::GSUM + l * LD + i], Rg[l * r + k], gr);
                acc = fma(gr, Qg[k * r + jj] + Qg[jj * r + k], acc);
            }
            g.g_R[(size_t)b * M * r + idx] = acc;
        }
    }
    if (valid && g.g_Q) {
        for (int idx = j; idx < r * r; idx += LPS) {
            int i = idx / r, jj = idx - i * r;
            double acc = 0;
            for (int k = 0; k < M; ++k) {
                double gr = 0;
                for (int l = 0; l < M; ++l) gr = fma(sm[S::GSUM + l * LD + k], Rg[l * r + jj], gr);
                acc = fma(Rg[k * r + i], gr, acc);
            }
            g.g_Q[(size_t)b * r * r + idx] = acc;
        }
    }
}

template <int M>
int launch(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    using C = Cfg<M>;
    const int warps = 2;
    const int series_per_cta = warps * C::SPW;
    const int grid = (g.B + series_per_cta - 1) / series_per_cta;
    const size_t smem = (size_t)series_per_cta * (backward ? BwdSmem<M>::TOTAL : FwdSmem<M>::TOTAL) * sizeof(double);
    cudaError_t e;
#define BKF_SW_LAUNCH(KERN)                                                                            \
    e = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }                       \
    KERN<<<grid, 32 * warps, smem, stream>>>(g);
    if (g.kind == BKF_UNIVARIATE) {
        if (backward) { BKF_SW_LAUNCH((backward_kernel<M, K_UNI>)) } else { BKF_SW_LAUNCH((forward_kernel<M, K_UNI>)) }
    } else {
        if (backward) { BKF_SW_LAUNCH((backward_kernel<M, K_STD>)) } else { BKF_SW_LAUNCH((forward_kernel<M, K_STD>)) }
    }
#undef BKF_SW_LAUNCH
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace sw
}  // namespace bkf
