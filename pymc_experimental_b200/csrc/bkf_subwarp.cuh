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
#include "bkf_common.cuh"

// The file is written once for a scalar type: bkf_subwarp.cu / bkf_subwarp_inst_*.cu include it as namespace sw with
// real = double (5 <= m <= 8); the bkf_subwarp_f32*.cu units define BKF_SW_F32 first and get namespace swf with
// real = float (5 <= m <= 16: the float32 kernels of the family that the FP64 tensor-core kernels serve in double
// precision; FFMA at twice the DFMA rate, half the shared-memory and record traffic).
#ifdef BKF_SW_F32
#ifndef BKF_SUBWARP_CUH_F32
#define BKF_SUBWARP_CUH_F32
#define BKF_SW_BODY
#define BKF_SW_NS swf
#define BKF_SW_REAL float
#define BKF_SW_REAL2 float2
#endif
#else
#ifndef BKF_SUBWARP_CUH_F64
#define BKF_SUBWARP_CUH_F64
#define BKF_SW_BODY
#define BKF_SW_NS sw
#define BKF_SW_REAL double
#define BKF_SW_REAL2 double2
#endif
#endif
#ifdef BKF_SW_BODY
#undef BKF_SW_BODY

namespace bkf {
namespace BKF_SW_NS {

using real = BKF_SW_REAL;
using real2 = BKF_SW_REAL2;
__device__ __forceinline__ real2 mk2(real x, real y) {
    real2 v;
    v.x = x;
    v.y = y;
    return v;
}

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

__device__ __forceinline__ real sub_sum(real v, int lps) {
    if (lps > 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// x[i] = S[k*LD + i], i < M: broadcast read of column k (all lanes of the series use the same address)
template <int M>
__device__ __forceinline__ void load_vec(const real* __restrict__ S, real (&x)[M]) {
    const real2* c = reinterpret_cast<const real2*>(S);
#pragma unroll
    for (int q = 0; q < M / 2; ++q) {
        real2 d = c[q];
        x[2 * q] = d.x;
        x[2 * q + 1] = d.y;
    }
    if (M & 1) x[M - 1] = S[M - 1];
}
template <int M>
__device__ __forceinline__ void store_vec(real* __restrict__ S, const real (&x)[M]) {
    real2* c = reinterpret_cast<real2*>(S);
#pragma unroll
    for (int q = 0; q < M / 2; ++q) c[q] = mk2(x[2 * q], x[2 * q + 1]);
    if (M & 1) S[M - 1] = x[M - 1];
}

// acc[:] = sum_k As[:, k] * b[k]   (As column-major in smem with leading dim LD, b in registers)
template <int M, int LD>
__device__ __forceinline__ void prod(const real* __restrict__ As, const real (&b)[M], real (&acc)[M]) {
#pragma unroll
    for (int k = 0; k < M; ++k) {
        real col[M];
        load_vec<M>(As + k * LD, col);
#pragma unroll
        for (int i = 0; i < M; ++i) acc[i] = (k == 0) ? col[i] * b[0] : fma(col[i], b[k], acc[i]);
    }
}

template <int M>
__device__ __forceinline__ real dot(const real (&x)[M], const real (&y)[M]) {
    real s = x[0] * y[0];
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
    static constexpr int IN0 = GT + C::MAT;           // dense trajectory record, real buffered
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
    real Pf[M];   // column j of P_f (after + eps I)
    real kv[M];   // K (all entries)
    real K, pz, zp, F, f0, v, af, keep, ll, zj, yhat;
    bool flag;
};

// ---------------------------------------------------------------------------------------------------
// Measurement update, p = 1.  Pc = column j of P, Pr = row j of P (aliases Pc when !ASYM).
// Vectors K (and pzt, zp for the standard filter) are exchanged through smem (vK, vP, vQ); zv = unmasked Z.
// ---------------------------------------------------------------------------------------------------
template <int M, int KIND, bool ASYM>
__device__ __forceinline__ void update(Upd<M>& u, const real (&Pc)[M], const real (&Pr)[M], real a_j, int j,
                                       real yv, real fill, real jitter, real Zj, const real (&zv)[M],
                                       real d0, real H0, real* vK, real* vP, real* vQ) {
    constexpr int LPS = Cfg<M>::LPS;
    const bool miss = (KIND == K_UNI) ? isnan(yv) : (isnan(yv) || yv == fill);  // :792 vs :352 (quirk Q6)
    const real wobs = miss ? real(0.0) : real(1.0);
    const real ym = miss ? real(0.0) : yv;
    const real zj = wobs * Zj;
    const real Hm = wobs * H0;
    u.zj = zj;
    u.flag = miss;
    const real yhat = d0 + sub_sum(zj * a_j, LPS);  // d + Z a (:594) == Z a + d (:770)
    u.yhat = yhat;
    u.v = ym - yhat;
    // pz_j = sum_i P[j,i] z_i (row j);  zp_j = sum_i z_i P[i,j] (column j)
    real pz = wobs * dot<M>(Pr, zv);
    u.pz = pz;
    u.zp = ASYM ? wobs * dot<M>(Pc, zv) : pz;
    const real f0 = sub_sum(zj * pz, LPS);
    u.f0 = f0;
    if (KIND == K_UNI) {
        const real F0 = f0 + Hm;
        const bool zf = (F0 == real(0.0)) || miss;  // :777
        const real F = F0 + jitter;
        u.keep = zf ? real(0.0) : real(1.0);
        u.F = F;
        u.K = pz / F * u.keep;  // :782
        u.af = fma(u.K, u.v, a_j);
        const real lli = zf ? real(0.0) : log(F) + u.v * u.v / F;  // :787
        u.ll = -real(0.5) * ((lli != real(0.0) ? real(kLog2Pi) : real(0.0)) + lli);    // :818
        if (j < M) vK[j] = u.K;
        __syncwarp();
        load_vec<M>(vK, u.kv);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            real pu = Pc[i] - (u.kv[i] * u.K) * F;  // :785
            if (ASYM) {
                real put = Pr[i] - (u.K * u.kv[i]) * F;
                pu = real(0.5) * (pu + put);  // :815 (exact no-op when P is symmetric)
            }
            u.Pf[i] = (i == j) ? pu + jitter : pu;
        }
    } else {
        const real F = f0 + (Hm + jitter);  // :598
        u.F = F;
        u.keep = real(1.0);
        u.K = pz / F;  // solve(F^T, PZT^T)^T, p = 1  (:600)
        u.af = fma(u.K, u.v, a_j);
        u.ll = miss ? real(0.0) : -real(0.5) * (real(kLog2Pi) + log(F) + u.v * (u.v / F));  // :611-615
        if (j < M) {
            vK[j] = u.K;
            vP[j] = pz;
            if (ASYM) vQ[j] = u.zp;
        }
        __syncwarp();
        real pv[M];
        load_vec<M>(vK, u.kv);
        load_vec<M>(vP, pv);
        real qv[M];
        if (ASYM) load_vec<M>(vQ, qv);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            // Joseph form with L = I - K z implicit:  Y = (P - K (zP)) - ((P z^T) - K f0) K^T     (:601-604)
            const real zp_i = ASYM ? qv[i] : pv[i];
            const real y_ij = fma(-(fma(-u.kv[i], f0, pv[i])), u.K, fma(-u.kv[i], u.zp, Pc[i]));
            const real y_ji = fma(-(fma(-u.K, f0, pz)), u.kv[i], fma(-u.K, zp_i, Pr[i]));
            const real khk = real(0.5) * ((u.kv[i] * Hm) * u.K + (u.K * Hm) * u.kv[i]);
            const real pf = real(0.5) * (y_ij + y_ji) + khk;
            u.Pf[i] = (i == j) ? pf + jitter : pf;  // stabilize, :536
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------
// TV: c, d, Z and / or H carry a time axis (g.tv_mask): re-read from x[t] at the top of every step (regression
// components, time-varying intercepts / measurement noise); gradients with a time axis are flushed per step.
template <int M, int KIND, bool TV = false>
__global__ void __launch_bounds__(64) forward_kernel(KArgs<real> g) {
    using C = Cfg<M>;
    using S = FwdSmem<M>;
    constexpr int LPS = C::LPS, LD = C::LD;
    extern __shared__ real smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPS, j = lane % LPS;
    const int series = (blockIdx.x * (blockDim.x >> 5) + warp) * C::SPW + sub;
    const bool valid = series < g.B;
    const int b = valid ? series : g.B - 1;
    real* sm = smem + (size_t)((warp * C::SPW + sub)) * S::TOTAL;
    // Rows j >= M of the shared matrices are never written but are read by the idle lanes of a series (M < LPS), whose
    // values then meet a zero weight: 0 * (whatever the previous kernel left there) must not be 0 * NaN.
    for (int i = j; i < S::TOTAL; i += LPS) sm[i] = real(0.0);
    __syncwarp();
    const int n = g.n;
    const bool act = j < M;

    // ---- per-series constants ----
    const real* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)M * M);
    const real* Zg = param_ptr(g.Z, g.shared_mask, BKF_PARAM_Z, b, (size_t)M);
    const real* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * g.r);
    const real* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)g.r * g.r);
    real H0 = *param_ptr(g.H, g.shared_mask, BKF_PARAM_H, b, 1);
    real d0 = *param_ptr(g.d, g.shared_mask, BKF_PARAM_D, b, 1);
    const real* cg = param_ptr(g.c, g.shared_mask, BKF_PARAM_C, b, (size_t)M);
    const real* a0g = param_ptr(g.a0, g.shared_mask, BKF_PARAM_A0, b, (size_t)M);
    const real* P0g = param_ptr(g.P0, g.shared_mask, BKF_PARAM_P0, b, (size_t)M * M);
    real Trow[M];  // row j of T
#pragma unroll
    for (int k = 0; k < M; ++k) Trow[k] = act ? Tg[j * M + k] : real(0.0);
#pragma unroll
    for (int k = 0; k < M; ++k) sm[S::T_cm + k * LD + j] = Trow[k];  // T[j,k] -> column k, row j
    real Zj = act ? Zg[j] : real(0.0);
    sm[S::VZ + j] = Zj;
    real cj = act ? cg[j] : real(0.0);
    {  // RQR = sym(R Q R^T): lane j computes column j (and row j) once per series
        const int r = g.r;
#pragma unroll 1
        for (int i = 0; i < M; ++i) {
            real s_ij = 0, s_ji = 0;
            if (act) {
                for (int k = 0; k < r; ++k) {
                    real rq_i = 0, rq_j = 0;
                    for (int l = 0; l < r; ++l) {
                        rq_i = fma(Rg[i * r + l], Qg[l * r + k], rq_i);
                        rq_j = fma(Rg[j * r + l], Qg[l * r + k], rq_j);
                    }
                    s_ij = fma(rq_i, Rg[j * r + k], s_ij);
                    s_ji = fma(rq_j, Rg[i * r + k], s_ji);
                }
            }
            sm[S::RQR + j * LD + i] = real(0.5) * (s_ij + s_ji);
        }
    }
    __syncwarp();
    real zv[M];
    load_vec<M>(sm + S::VZ, zv);

    // ---- state ----
    real Pc[M], Pr[M];
    real a_j = act ? a0g[j] : real(0.0);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        Pc[i] = act ? P0g[i * M + j] : real(0.0);
        Pr[i] = act ? P0g[j * M + i] : real(0.0);
    }
    const real* yb = g.y_shared ? g.y : g.y + (size_t)b * n;
    const size_t rec = (size_t)C::DENSE;
    real logp = 0;
    Upd<M> u;
    for (int t = 0; t < n; ++t) {
        const size_t bt = (size_t)b * n + t;
        const real yv = yb[t];
        // ---- emit the prediction (a, P) of step t: trajectory for the adjoint and/or user outputs ----
        if (g.traj || g.pred_a || g.pred_P) {
            real* o = sm + S::OUT;
            if (act) {
                o[j] = a_j;
#pragma unroll
                for (int i = 0; i < M; ++i) o[M + i * M + j] = Pc[i];  // row-major P
            }
            __syncwarp();
            if (valid) {
                if (g.traj) {
                    real2* dst = reinterpret_cast<real2*>(g.traj + bt * rec);
                    const real2* src = reinterpret_cast<const real2*>(o);
                    for (int q = j; q < C::DENSE / 2; q += LPS) dst[q] = src[q];
                }
                if (g.pred_a && act) g.pred_a[bt * M + j] = a_j;
                if (g.pred_P)
                    for (int q = j; q < M * M; q += LPS) g.pred_P[bt * M * M + q] = o[M + q];
            }
            __syncwarp();
        }
        if (TV) {
            const real* Zt = param_ptr_t(g.Z, g.shared_mask, g.tv_mask, BKF_PARAM_Z, b, t, n, (size_t)M);
            const real* ct = param_ptr_t(g.c, g.shared_mask, g.tv_mask, BKF_PARAM_C, b, t, n, (size_t)M);
#pragma unroll
            for (int i = 0; i < M; ++i) zv[i] = Zt[i];
            Zj = act ? Zt[j] : real(0.0);
            cj = act ? ct[j] : real(0.0);
            H0 = *param_ptr_t(g.H, g.shared_mask, g.tv_mask, BKF_PARAM_H, b, t, n, 1);
            d0 = *param_ptr_t(g.d, g.shared_mask, g.tv_mask, BKF_PARAM_D, b, t, n, 1);
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
        real X[M];
        prod<M, LD>(sm + S::T_cm, u.Pf, X);
        store_vec<M>(sm + S::X + j * LD, X);
        __syncwarp();
        {
            real av[M];
            load_vec<M>(sm + S::VA, av);
            a_j = dot<M>(Trow, av) + cj;
        }
        real Y[M];
        prod<M, LD>(sm + S::X, Trow, Y);
        store_vec<M>(sm + S::Y + j * LD, Y);
        __syncwarp();
        {
            real rq[M];
            load_vec<M>(sm + S::RQR + j * LD, rq);
#pragma unroll
            for (int i = 0; i < M; ++i) Pc[i] = real(0.5) * (Y[i] + sm[S::Y + i * LD + j]) + rq[i];
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
template <int BYTES>
__device__ __forceinline__ void cp_async_pair(void* smem_dst, const void* gsrc) {
    if constexpr (BYTES == 16) {
        cp_async16(smem_dst, gsrc);
    } else {
        unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int M, int KIND, bool TV = false>
__global__ void __launch_bounds__(64) backward_kernel(KArgs<real> g) {
    using C = Cfg<M>;
    using S = BwdSmem<M>;
    constexpr int LPS = C::LPS, LD = C::LD;
    extern __shared__ real smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPS, j = lane % LPS;
    const int series = (blockIdx.x * (blockDim.x >> 5) + warp) * C::SPW + sub;
    const bool valid = series < g.B;
    const int b = valid ? series : g.B - 1;
    real* sm = smem + (size_t)((warp * C::SPW + sub)) * S::TOTAL;
    // Rows j >= M of the shared matrices are never written but are read by the idle lanes of a series (M < LPS), whose
    // values then meet a zero weight: 0 * (whatever the previous kernel left there) must not be 0 * NaN.
    for (int i = j; i < S::TOTAL; i += LPS) sm[i] = real(0.0);
    __syncwarp();
    const int n = g.n;
    const bool act = j < M;

    const real* Tg = param_ptr(g.T, g.shared_mask, BKF_PARAM_T, b, (size_t)M * M);
    const real* Zg = param_ptr(g.Z, g.shared_mask, BKF_PARAM_Z, b, (size_t)M);
    real H0 = *param_ptr(g.H, g.shared_mask, BKF_PARAM_H, b, 1);
    real d0 = *param_ptr(g.d, g.shared_mask, BKF_PARAM_D, b, 1);
    real Tcol[M];  // column j of T
#pragma unroll
    for (int k = 0; k < M; ++k) Tcol[k] = act ? Tg[k * M + j] : real(0.0);
#pragma unroll
    for (int k = 0; k < M; ++k) sm[S::T_rm + k * LD + j] = Tcol[k];  // T[k,j] -> "column" k (row k of T), entry j
    real Zj = act ? Zg[j] : real(0.0);
    sm[S::VZ + j] = Zj;
#pragma unroll
    for (int i = 0; i < LPS; ++i) {
        sm[S::GSUM + j * LD + i] = real(0.0);
        sm[S::GT + j * LD + i] = real(0.0);
    }
    __syncwarp();
    real zv[M];
    load_vec<M>(sm + S::VZ, zv);

    const real* yb = g.y_shared ? g.y : g.y + (size_t)b * n;
    const size_t rec = (size_t)C::DENSE;
    auto prefetch = [&](int t, int buf) {
        const char* src = reinterpret_cast<const char*>(g.traj + ((size_t)b * n + t) * rec);
        char* dst = reinterpret_cast<char*>(sm + (buf ? S::IN1 : S::IN0));
        constexpr int CH = (int)sizeof(real2);  // one pair of values per copy: 16 bytes (double), 8 bytes (float)
        for (int q = j; q < C::DENSE / 2; q += LPS) cp_async_pair<CH>(dst + q * CH, src + q * CH);
        cp_async_commit();
    };
    real abar = real(0.0), gz = real(0.0), gc = real(0.0), gH = real(0.0), gd = real(0.0);
    real Pbar[M];
#pragma unroll
    for (int i = 0; i < M; ++i) Pbar[i] = real(0.0);
    prefetch(n - 1, (n - 1) & 1);
    Upd<M> u;
    for (int t = n - 1; t >= 0; --t) {
        const size_t bt = (size_t)b * n + t;
        const real gcot = g.ll_cot ? g.ll_cot[bt] : real(1.0);
        const real yv = yb[t];
        cp_async_wait<0>();
        __syncwarp();
        if (t > 0) prefetch(t - 1, (t - 1) & 1);
        const real* in = sm + ((t & 1) ? S::IN1 : S::IN0);
        if (TV) {
            const real* Zt = param_ptr_t(g.Z, g.shared_mask, g.tv_mask, BKF_PARAM_Z, b, t, n, (size_t)M);
#pragma unroll
            for (int i = 0; i < M; ++i) zv[i] = Zt[i];
            Zj = act ? Zt[j] : real(0.0);
            H0 = *param_ptr_t(g.H, g.shared_mask, g.tv_mask, BKF_PARAM_H, b, t, n, 1);
            d0 = *param_ptr_t(g.d, g.shared_mask, g.tv_mask, BKF_PARAM_D, b, t, n, 1);
        }
        const real a_j = act ? in[j] : real(0.0);
        real Pc[M];
        real S2c[M];  // column j of P + P^T (standard adjoint only)
        if (t == 0) {
            real Pr[M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                Pc[i] = act ? in[M + i * M + j] : real(0.0);
                Pr[i] = act ? in[M + j * M + i] : real(0.0);
            }
            update<M, KIND, true>(u, Pc, Pr, a_j, j, yv, g.fill, g.jitter, Zj, zv, d0, H0, sm + S::VK, sm + S::VP, sm + S::VQ);
#pragma unroll
            for (int i = 0; i < M; ++i) S2c[i] = Pc[i] + Pr[i];
        } else {
#pragma unroll
            for (int i = 0; i < M; ++i) Pc[i] = act ? in[M + j * M + i] : real(0.0);
            update<M, KIND, false>(u, Pc, Pc, a_j, j, yv, g.fill, g.jitter, Zj, zv, d0, H0, sm + S::VK, sm + S::VP, sm + S::VQ);
#pragma unroll
            for (int i = 0; i < M; ++i) S2c[i] = real(2.0) * Pc[i];
        }
        // ================= predict adjoint =================
        // G = S(Pbar+)
        store_vec<M>(sm + S::GW + j * LD, Pbar);
        if (act) sm[S::VA + j] = abar;
        __syncwarp();
        real Gc[M];
#pragma unroll
        for (int i = 0; i < M; ++i) Gc[i] = real(0.5) * (Pbar[i] + sm[S::GW + i * LD + j]);
        __syncwarp();
        store_vec<M>(sm + S::GW + j * LD, Gc);
        {
            real gs[M];
            load_vec<M>(sm + S::GSUM + j * LD, gs);
#pragma unroll
            for (int i = 0; i < M; ++i) gs[i] += Gc[i];
            store_vec<M>(sm + S::GSUM + j * LD, gs);
        }
        __syncwarp();
        // W = G T
        real W[M];
        prod<M, LD>(sm + S::GW, Tcol, W);
        __syncwarp();
        store_vec<M>(sm + S::GW + j * LD, W);
        __syncwarp();
        // Pfbar = T^T W
        real Pfb[M];
        prod<M, LD>(sm + S::T_rm, W, Pfb);
        // gT += W (Pf + Pf^T) + abar+ af^T      (Pf is symmetric: 2 W Pf)
        real abv[M];
        load_vec<M>(sm + S::VA, abv);
        {
            real tmp[M], gt[M];
            prod<M, LD>(sm + S::GW, u.Pf, tmp);
            load_vec<M>(sm + S::GT + j * LD, gt);
#pragma unroll
            for (int i = 0; i < M; ++i) gt[i] += fma(abv[i], u.af, real(2.0) * tmp[i]);
            store_vec<M>(sm + S::GT + j * LD, gt);
        }
        const real afbar = dot<M>(Tcol, abv);
        gc += abar;
        // ================= update adjoint =================
        if (KIND == K_UNI) {
            // Branch-free on purpose: the series sharing a warp may disagree on `keep` (missing data), and the
            // shuffles / __syncwarp below must be executed by all 32 lanes.  When keep == 0 (K = 0: a and P pass
            // through, no parameter gradient) the computed values are discarded by the selects.
            const bool keep = u.keep != real(0.0);
            const real llbar = -real(0.5) * gcot;
            const real F = u.F, v = u.v;
            const real pk = dot<M>(Pfb, u.kv);  // (Pu_bar K)_j, Pu_bar = S(Pfbar) ~ Pfbar
            const real Kbar = fma(-real(2.0) * F, pk, afbar * v);
            const real kpk = sub_sum(u.K * pk, LPS);
            const real ka = sub_sum(u.K * afbar, LPS);
            const real kk = sub_sum(Kbar * u.K, LPS);
            const real Fbar = -kpk + llbar * (real(1.0) / F - v * v / (F * F)) - kk / F;
            const real vbar = ka + llbar * real(2.0) * v / F;
            const real pzbar = Kbar / F + Fbar * u.zj;
            if (act) sm[S::VP + j] = pzbar;
            __syncwarp();
            real pzv[M];
            load_vec<M>(sm + S::VP, pzv);
            const real zbar = fma(Fbar, u.pz, dot<M>(Pc, pzv)) - vbar * a_j;
            abar = keep ? fma(-vbar, u.zj, afbar) : afbar;
#pragma unroll
            for (int i = 0; i < M; ++i) Pbar[i] = keep ? fma(pzv[i], u.zj, Pfb[i]) : Pfb[i];
            gz += keep ? zbar : real(0.0);
            gH += keep ? Fbar : real(0.0);
            gd -= keep ? vbar : real(0.0);
            __syncwarp();
        } else {
            // G2 = S(Pfbar) ~ Pfbar (symmetric up to rounding): column j doubles as row j
            const real F = u.F, v = u.v, f0 = u.f0;
            const real Hm = u.flag ? real(0.0) : H0;
            const real wv = v / F;
            // xz2 = (P + P^T) z - 2 f0 K ;  u1 = G2 xz2 ;  gk = G2 K
            const real xz2_j = (u.pz + u.zp) - real(2.0) * f0 * u.K;
            if (act) {
                sm[S::VP + j] = xz2_j;
            }
            __syncwarp();
            real xv[M];
            load_vec<M>(sm + S::VP, xv);
            const real u1 = dot<M>(Pfb, xv);
            const real gk = dot<M>(Pfb, u.kv);
            const real kgk = sub_sum(u.K * gk, LPS);
            if (act) sm[S::VQ + j] = gk;
            __syncwarp();
            real gkv[M];
            load_vec<M>(sm + S::VQ, gkv);
            // Zmbar = -K^T Lbar = -(gk^T (P + P^T) - (gk.K) z^T (P + P^T))
            real zmbar = -(dot<M>(gkv, S2c) - kgk * (u.zp + u.pz));
            // Kbar = G2 K (Hm + Hm^T) + afbar v - Lbar z
            real Kbar = fma(real(2.0) * Hm, gk, afbar * v) - u1;
            real Hmbar = kgk;
            real vbar = sub_sum(u.K * afbar, LPS);
            real Fbar = real(0.0);
            if (!u.flag) {
                Fbar = -real(0.5) * gcot * (real(1.0) / F - wv * wv);
                vbar -= gcot * wv;
            }
            // K = PZT / F
            real pztbar = Kbar / F;
            Fbar -= sub_sum(u.K * pztbar, LPS);
            // F = Zm PZT + Hm + eps
            zmbar = fma(Fbar, u.pz, zmbar);
            pztbar = fma(Fbar, u.zj, pztbar);
            Hmbar += Fbar;
            if (act) sm[S::VR + j] = pztbar;
            __syncwarp();
            real pbv[M];
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
                real pb = Pfb[i] - zv[i] * (u.flag ? real(0.0) : real(1.0)) * gk - gkv[i] * u.zj;
                pb = fma(zv[i] * (u.flag ? real(0.0) : real(1.0)) * u.zj, kgk, pb);
                Pbar[i] = fma(pbv[i], u.zj, pb);
            }
            if (!u.flag) {
                gz += zmbar;
                gH += Hmbar;
            }
            __syncwarp();
        }
        if (TV) {  // a gradient with a time axis is written out and cleared after every step
            if ((g.tv_mask >> BKF_PARAM_C) & 1u) {
                if (valid && g.g_c && act) g.g_c[bt * M + j] = gc;
                gc = real(0.0);
            }
            if ((g.tv_mask >> BKF_PARAM_Z) & 1u) {
                if (valid && g.g_Z && act) g.g_Z[bt * M + j] = gz;
                gz = real(0.0);
            }
            if ((g.tv_mask >> BKF_PARAM_D) & 1u) {
                if (valid && g.g_d && j == 0) g.g_d[bt] = gd;
                gd = real(0.0);
            }
            if ((g.tv_mask >> BKF_PARAM_H) & 1u) {
                if (valid && g.g_H && j == 0) g.g_H[bt] = gH;
                gH = real(0.0);
            }
        }
    }
    // ---- gradients out ----
    if (valid) {
        const bool tvC = TV && ((g.tv_mask >> BKF_PARAM_C) & 1u), tvZ = TV && ((g.tv_mask >> BKF_PARAM_Z) & 1u),
                   tvD = TV && ((g.tv_mask >> BKF_PARAM_D) & 1u), tvH = TV && ((g.tv_mask >> BKF_PARAM_H) & 1u);
        if (g.g_a0 && act) g.g_a0[(size_t)b * M + j] = abar;
        if (g.g_c && act && !tvC) g.g_c[(size_t)b * M + j] = gc;
        if (g.g_Z && act && !tvZ) g.g_Z[(size_t)b * M + j] = gz;
        if (g.g_d && j == 0 && !tvD) g.g_d[b] = gd;
        if (g.g_H && j == 0 && !tvH) g.g_H[b] = gH;
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
    const real* Rg = param_ptr(g.Rm, g.shared_mask, BKF_PARAM_R, b, (size_t)M * r);
    const real* Qg = param_ptr(g.Q, g.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    if (valid && g.g_R) {
        for (int idx = j; idx < M * r; idx += LPS) {
            int i = idx / r, jj = idx - i * r;
            real acc = 0;
            for (int k = 0; k < r; ++k) {
                real gr = 0;
                for (int l = 0; l < M; ++l) gr = fma(sm[S::GSUM + l * LD + i], Rg[l * r + k], gr);
                acc = fma(gr, Qg[k * r + jj] + Qg[jj * r + k], acc);
            }
            g.g_R[(size_t)b * M * r + idx] = acc;
        }
    }
    if (valid && g.g_Q) {
        for (int idx = j; idx < r * r; idx += LPS) {
            int i = idx / r, jj = idx - i * r;
            real acc = 0;
            for (int k = 0; k < M; ++k) {
                real gr = 0;
                for (int l = 0; l < M; ++l) gr = fma(sm[S::GSUM + l * LD + k], Rg[l * r + jj], gr);
                acc = fma(Rg[k * r + i], gr, acc);
            }
            g.g_Q[(size_t)b * r * r + idx] = acc;
        }
    }
}

template <int M>
int launch(const KArgs<real>& g, bool backward, cudaStream_t stream, const char** err) {
    using C = Cfg<M>;
    const int warps = 2;
    const int series_per_cta = warps * C::SPW;
    const int grid = (g.B + series_per_cta - 1) / series_per_cta;
    const size_t smem = (size_t)series_per_cta * (backward ? BwdSmem<M>::TOTAL : FwdSmem<M>::TOTAL) * sizeof(real);
    cudaError_t e;
#define BKF_SW_LAUNCH(KERN)                                                                            \
    e = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }                       \
    KERN<<<grid, 32 * warps, smem, stream>>>(g);
#ifdef BKF_SW_F32
    constexpr bool HAS_TV = false;  // the float32 build keeps time-varying inputs on the generic kernels
#else
    constexpr bool HAS_TV = true;
#endif
    const bool tv = g.tv_mask != 0;
    if (tv && !HAS_TV) { *err = "time-varying inputs are not compiled into the float32 sub-warp kernels"; return BKF_ERR_UNSUPPORTED; }
    if (g.kind == BKF_UNIVARIATE) {
        if (tv) {
            if (backward) { BKF_SW_LAUNCH((backward_kernel<M, K_UNI, HAS_TV>)) } else { BKF_SW_LAUNCH((forward_kernel<M, K_UNI, HAS_TV>)) }
        } else {
            if (backward) { BKF_SW_LAUNCH((backward_kernel<M, K_UNI>)) } else { BKF_SW_LAUNCH((forward_kernel<M, K_UNI>)) }
        }
    } else {
        if (tv) {
            if (backward) { BKF_SW_LAUNCH((backward_kernel<M, K_STD, HAS_TV>)) } else { BKF_SW_LAUNCH((forward_kernel<M, K_STD, HAS_TV>)) }
        } else {
            if (backward) { BKF_SW_LAUNCH((backward_kernel<M, K_STD>)) } else { BKF_SW_LAUNCH((forward_kernel<M, K_STD>)) }
        }
    }
#undef BKF_SW_LAUNCH
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace BKF_SW_NS
}  // namespace bkf
#endif  // BKF_SW_BODY
