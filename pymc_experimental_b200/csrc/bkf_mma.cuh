// bkf_mma.cuh -- "warp per series" kernels on the FP64 tensor-core path (mma.sync.m8n8k4.f64, DMMA),
// 9 <= m <= 16 (padded to 16), p = 1, float64, StandardFilter and UnivariateFilter, forward and adjoint.
//
// Mapping (DESIGN.md "Kernels").  One warp owns one time-sequential filter.  Every 16 x 16 state matrix lives in
// registers as four DMMA accumulator fragments ("C layout"): with lane = 4 g + t,
//     X.v[I][J][e] = X[8 I + g][8 J + 2 t + e]                        (8 doubles per lane)
// The only product primitive is  D (+)= X Y^T  (mm_nt): the accumulator registers of X are fed *unchanged* as the
// A operand (the k index is permuted, k-slice (K, e) = columns {8K + 2t + e}), and the accumulator registers of Y
// are fed unchanged as the B operand (B[k][n] = Y[n][k]) with the same permutation -- so chains of products need
// no shuffles and no shared memory:
//     forward   W = T P_f = mm_nt(T, P_f)   (P_f symmetric)        P+ = W T^T = mm_nt(W, T)
//     adjoint   W' = T^T G = mm_nt(T^T, G)  (G symmetric)          Pbar_f = W' T = mm_nt(W', T^T)
//               Tbar += 2 (G T) P_f = 2 mm_nt(V, P_f),  V = W'^T   (one transpose through shared memory)
// Vectors live either in "row form" (x[8I+g], replicated over t) or "column form" (x[8J+2t+e], replicated over
// g); matrix-vector products and rank-1 updates are a handful of FMAs plus quad shuffles.  The prediction
// trajectory (a_t|t-1, P_t|t-1) that the adjoint needs is spilled in fragment order (perfectly coalesced
// 16-byte accesses, REC = 292 doubles per state-step: P, a, and the update scalars P z^T, v, F the adjoint reuses).
//
// Reference formulas: kalman_filter.py:529-539 (step), :594-617 (StandardFilter.update, p = 1),
// :769-820 (UnivariateFilter, p = 1), :407-408 (predict); adjoint: SURVEY.md Appendix B specialised to p = 1.
#pragma once
#include "bkf_common.cuh"

namespace bkf {
namespace tc {

enum { K_STD = 0, K_UNI = 1 };
constexpr int MP = 16;              // padded state dimension
constexpr int REC_A = MP * MP;      // trajectory record: P fragments (256) | a (16) | P z^T (16) | v, F, 1/F-or-0, pad
constexpr int REC_PZ = REC_A + MP;
constexpr int REC_S = REC_PZ + MP;
constexpr int REC = REC_S + 4;      // 292 doubles per state-step
constexpr int SLD = 18;             // leading dimension of the per-warp transpose scratch (conflict-free, 16 B rows)
constexpr int SCRATCH = MP * SLD;   // doubles of scratch per warp
constexpr unsigned FULL = 0xffffffffu;


struct Frag { double v[2][2][2]; };  // v[I][J][e] = X[8I+g][8J+2t+e]
struct RVec { double v[2]; };        // v[I]       = x[8I+g]
struct CVec { double v[2][2]; };     // v[J][e]    = x[8J+2t+e]

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// D (+)= X Y^T, all three 16 x 16 in the C layout; 16 DMMAs, four independent accumulator chains
template <bool ACC>
__device__ __forceinline__ void mm_nt(Frag& D, const Frag& X, const Frag& Y) {
    if (!ACC) {
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J < 2; ++J) D.v[I][J][0] = D.v[I][J][1] = 0.0;
    }
#pragma unroll
    for (int K = 0; K < 2; ++K)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J < 2; ++J) dmma(D.v[I][J][0], D.v[I][J][1], X.v[I][K][e], Y.v[J][K][e]);
}

// Tiles (0,0), (0,1), (1,1) of D += X Y^T (12 DMMAs) for products whose result is symmetric; D is NOT cleared.
__device__ __forceinline__ void mm_nt_upper(Frag& D, const Frag& X, const Frag& Y) {
#pragma unroll
    for (int K = 0; K < 2; ++K)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            dmma(D.v[0][0][0], D.v[0][0][1], X.v[0][K][e], Y.v[0][K][e]);
            dmma(D.v[0][1][0], D.v[0][1][1], X.v[0][K][e], Y.v[1][K][e]);
            dmma(D.v[1][1][0], D.v[1][1][1], X.v[1][K][e], Y.v[1][K][e]);
        }
}
// Stage tiles (0,0), (0,1), (1,1) of X in the scratch and read back (0,0)^T, (1,1)^T and (0,1)^T:
// Xt.v[0][0] = X00^T, Xt.v[1][1] = X11^T, Xt.v[1][0] = X01^T  (Xt.v[0][1] is left untouched).
__device__ __forceinline__ void transpose_upper(Frag& Xt, const Frag& X, double* sc, int g, int t) {
    *reinterpret_cast<double2*>(sc + g * SLD + 2 * t) = make_double2(X.v[0][0][0], X.v[0][0][1]);
    *reinterpret_cast<double2*>(sc + g * SLD + 8 + 2 * t) = make_double2(X.v[0][1][0], X.v[0][1][1]);
    *reinterpret_cast<double2*>(sc + (8 + g) * SLD + 8 + 2 * t) = make_double2(X.v[1][1][0], X.v[1][1][1]);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        Xt.v[0][0][e] = sc[(2 * t + e) * SLD + g];
        Xt.v[1][1][e] = sc[(8 + 2 * t + e) * SLD + 8 + g];
        Xt.v[1][0][e] = sc[(2 * t + e) * SLD + 8 + g];
    }
    __syncwarp();
}

__device__ __forceinline__ double red_t(double x) {  // sum over the 4 lanes of a quad (t)
    x += __shfl_xor_sync(FULL, x, 1);
    x += __shfl_xor_sync(FULL, x, 2);
    return x;
}
__device__ __forceinline__ double red_g(double x) {  // sum over the 8 quads (g)
    x += __shfl_xor_sync(FULL, x, 4);
    x += __shfl_xor_sync(FULL, x, 8);
    x += __shfl_xor_sync(FULL, x, 16);
    return x;
}

// (X x)[8I+g] in row form, x in column form
__device__ __forceinline__ RVec mv(const Frag& X, const CVec& x) {
    RVec r;
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        double s = X.v[I][0][0] * x.v[0][0];
        s = fma(X.v[I][0][1], x.v[0][1], s);
        s = fma(X.v[I][1][0], x.v[1][0], s);
        s = fma(X.v[I][1][1], x.v[1][1], s);
        r.v[I] = red_t(s);
    }
    return r;
}
__device__ __forceinline__ double dot_c(const CVec& a, const CVec& b) {
    double s = a.v[0][0] * b.v[0][0];
    s = fma(a.v[0][1], b.v[0][1], s);
    s = fma(a.v[1][0], b.v[1][0], s);
    s = fma(a.v[1][1], b.v[1][1], s);
    return red_t(s);
}
__device__ __forceinline__ double dot_r(const RVec& a, const RVec& b) {
    return red_g(fma(a.v[1], b.v[1], a.v[0] * b.v[0]));
}
__device__ __forceinline__ CVec row2col(const RVec& r, int t) {
    CVec c;
#pragma unroll
    for (int J = 0; J < 2; ++J)
#pragma unroll
        for (int e = 0; e < 2; ++e) c.v[J][e] = __shfl_sync(FULL, r.v[J], 4 * (2 * t + e) + t);
    return c;
}
// out = in^T through the per-warp scratch (write 16-byte row pieces, read transposed: 2 wavefronts per LDS.64)
__device__ __forceinline__ void transpose(Frag& out, const Frag& in, double* sc, int g, int t) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
            *reinterpret_cast<double2*>(sc + (8 * I + g) * SLD + 8 * J + 2 * t) = make_double2(in.v[I][J][0], in.v[I][J][1]);
    __syncwarp();
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) out.v[I][J][e] = sc[(8 * J + 2 * t + e) * SLD + 8 * I + g];
    __syncwarp();
}

// fragment <- row-major m x m matrix in global memory (zero padded); TR: load the transpose
template <bool TR>
__device__ __forceinline__ void load_frag(Frag& X, const double* __restrict__ M, int m, int g, int t) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * I + g, j = 8 * J + 2 * t + e;
                X.v[I][J][e] = (i < m && j < m) ? (TR ? M[j * m + i] : M[i * m + j]) : 0.0;
            }
}
__device__ __forceinline__ void store_frag(double* __restrict__ M, const Frag& X, int m, int g, int t, double scale) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * I + g, j = 8 * J + 2 * t + e;
                if (i < m && j < m) M[i * m + j] = scale * X.v[I][J][e];
            }
}
__device__ __forceinline__ RVec load_rvec(const double* __restrict__ x, int m, int g) {
    RVec r;
#pragma unroll
    for (int I = 0; I < 2; ++I) r.v[I] = (8 * I + g < m) ? x[8 * I + g] : 0.0;
    return r;
}
__device__ __forceinline__ CVec load_cvec(const double* __restrict__ x, int m, int t) {
    CVec c;
#pragma unroll
    for (int J = 0; J < 2; ++J)
#pragma unroll
        for (int e = 0; e < 2; ++e) c.v[J][e] = (8 * J + 2 * t + e < m) ? x[8 * J + 2 * t + e] : 0.0;
    return c;
}
__device__ __forceinline__ void store_rvec(double* __restrict__ x, const RVec& r, int m, int g, int t) {
    if (t == 0) {
#pragma unroll
        for (int I = 0; I < 2; ++I)
            if (8 * I + g < m) x[8 * I + g] = r.v[I];
    }
}

// trajectory record I/O (fragment order)
__device__ __forceinline__ void store_rec(double* __restrict__ base, const Frag& P, const CVec& a_c, int lane) {
    double2* dst = reinterpret_cast<double2*>(base);
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J) __stcs(dst + (2 * I + J) * 32 + lane, make_double2(P.v[I][J][0], P.v[I][J][1]));
    if (lane < 8) {  // lane L holds chunk L of a in column form: t = L & 3, J = L >> 2
        const int J = lane >> 2;
        __stcs(dst + REC_A / 2 + lane, make_double2(J ? a_c.v[1][0] : a_c.v[0][0], J ? a_c.v[1][1] : a_c.v[0][1]));
    }
}
__device__ __forceinline__ void load_rec(const double* __restrict__ base, Frag& P, RVec& a_r, CVec& a_c, int lane) {
    const double2* src = reinterpret_cast<const double2*>(base);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J) {
            const double2 d = __ldcs(src + (2 * I + J) * 32 + lane);
            P.v[I][J][0] = d.x;
            P.v[I][J][1] = d.y;
        }
#pragma unroll
    for (int J = 0; J < 2; ++J) {
        const double2 d = __ldcs(src + 128 + 4 * J + t);
        a_c.v[J][0] = d.x;
        a_c.v[J][1] = d.y;
    }
#pragma unroll
    for (int I = 0; I < 2; ++I) a_r.v[I] = __ldcs(base + MP * MP + 8 * I + g);
}

// Update quantities of a step that the forward sweep has computed anyway and the adjoint would otherwise recompute
// (a matrix-vector product, two reductions and a division per step): P z^T in both vector forms, v, F and
// rFk = 1/F (univariate: 0 when the update is skipped, kalman_filter.py:777-782).
struct Aux {
    RVec pz_r;
    CVec pz_c;
    double v, F, rFk;
};
__device__ __forceinline__ void store_aux(double* __restrict__ base, const CVec& pz_c, double v, double F, double rFk,
                                          int lane) {
    // one 16-byte store per lane 8..17: lanes 8..15 write chunk (lane - 8) of P z^T (column form: t = lane & 3,
    // J = (lane >> 2) & 1), lane 16 writes (v, F), lane 17 writes (rFk, 0)
    const int J = (lane >> 2) & 1;
    double2 val = make_double2(J ? pz_c.v[1][0] : pz_c.v[0][0], J ? pz_c.v[1][1] : pz_c.v[0][1]);
    if (lane == 16) val = make_double2(v, F);
    if (lane == 17) val = make_double2(rFk, 0.0);
    if (lane >= 8 && lane < 18) __stcs(reinterpret_cast<double2*>(base + REC_PZ) + (lane - 8), val);
}
__device__ __forceinline__ void load_aux(const double* __restrict__ base, Aux& X, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const double2* pc = reinterpret_cast<const double2*>(base + REC_PZ);
#pragma unroll
    for (int J = 0; J < 2; ++J) {
        const double2 d = __ldcs(pc + 4 * J + t);
        X.pz_c.v[J][0] = d.x;
        X.pz_c.v[J][1] = d.y;
    }
    X.pz_r.v[0] = __ldcs(base + REC_PZ + g);
    X.pz_r.v[1] = __ldcs(base + REC_PZ + 8 + g);
    const double2 d = __ldcs(reinterpret_cast<const double2*>(base + REC_S));
    X.v = d.x;
    X.F = d.y;
    X.rFk = __ldcs(base + REC_S + 2);
}

// sym(R Q R^T) in the C layout (once per series)
__device__ __forceinline__ void rqr_frag(Frag& X, const double* __restrict__ Rg, const double* __restrict__ Qg, int m,
                                         int r, int g, int t) {
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * I + g, j = 8 * J + 2 * t + e;
                double s_ij = 0, s_ji = 0;
                if (i < m && j < m) {
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
                X.v[I][J][e] = 0.5 * (s_ij + s_ji);
            }
}

// Quantities of one measurement update that the predict step / the adjoint need.
struct Upd {
    Frag Pf;  // P_f (after + eps I), exactly symmetric
    RVec K_r, af_r, pz_r, zp_r, zm_r;
    CVec K_c, af_c, pz_c, zp_c, zm_c;
    double F, f0, v, keep, ll, yhat, Hm;
    bool flag;
};

// ---------------------------------------------------------------------------------------------------
// Measurement update, p = 1.  P = prediction covariance, Pt = its transpose (only read when ASYM: the user's P0
// at t = 0 may be asymmetric; for t >= 1 every covariance is exactly symmetric).
// ---------------------------------------------------------------------------------------------------
template <int KIND, bool ASYM>
__device__ __forceinline__ void update(Upd& u, const Frag& P, const Frag& Pt, const RVec& a_r, const CVec& a_c,
                                       double yv, double fill, double jitter, const RVec& z_r, const CVec& z_c,
                                       double d0, double H0, int g, int t) {
    const bool miss = (KIND == K_UNI) ? isnan(yv) : (isnan(yv) || yv == fill);  // :792 vs :352 (quirk Q6)
    const double wobs = miss ? 0.0 : 1.0;
    const double ym = miss ? 0.0 : yv;
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        u.zm_r.v[I] = wobs * z_r.v[I];
        u.zm_c.v[I][0] = wobs * z_c.v[I][0];
        u.zm_c.v[I][1] = wobs * z_c.v[I][1];
    }
    const double Hm = wobs * H0;
    u.Hm = Hm;
    u.flag = miss;
    u.yhat = d0 + dot_c(u.zm_c, a_c);  // d + Z a (:594) == Z a + d (:770)
    u.v = ym - u.yhat;
    u.pz_r = mv(P, u.zm_c);  // P z^T
    u.pz_c = row2col(u.pz_r, t);
    if (ASYM && KIND == K_STD) {
        u.zp_r = mv(Pt, u.zm_c);  // (z P)^T
        u.zp_c = row2col(u.zp_r, t);
    } else {
        u.zp_r = u.pz_r;
        u.zp_c = u.pz_c;
    }
    const double f0 = dot_c(u.zm_c, u.pz_c);
    u.f0 = f0;
    if (KIND == K_UNI) {
        const double F0 = f0 + Hm;
        const bool zf = (F0 == 0.0) || miss;  // :777
        const double F = F0 + jitter;
        u.keep = zf ? 0.0 : 1.0;
        u.F = F;
        const double rF = u.keep / F;  // K = PZT / F (1 - zf), :782
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            u.K_r.v[I] = u.pz_r.v[I] * rF;
            u.af_r.v[I] = fma(u.K_r.v[I], u.v, a_r.v[I]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                u.K_c.v[I][e] = u.pz_c.v[I][e] * rF;
                u.af_c.v[I][e] = fma(u.K_c.v[I][e], u.v, a_c.v[I][e]);
            }
        }
        const double lli = zf ? 0.0 : log(F) + u.v * u.v / F;  // :787
        u.ll = -0.5 * ((lli != 0.0 ? kLog2Pi : 0.0) + lli);    // :818
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double ps = ASYM ? 0.5 * (P.v[I][J][e] + Pt.v[I][J][e]) : P.v[I][J][e];  // :815
                    double pu = ps - (u.K_r.v[I] * u.K_c.v[J][e]) * F;                             // :785
                    if (I == J && g == 2 * t + e) pu += jitter;
                    u.Pf.v[I][J][e] = pu;
                }
    } else {
        const double F = f0 + (Hm + jitter);  // :598
        u.F = F;
        u.keep = 1.0;
        const double rF = 1.0 / F;  // K = solve(F^T, PZT^T)^T, p = 1 (:600)
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            u.K_r.v[I] = u.pz_r.v[I] * rF;
            u.af_r.v[I] = fma(u.K_r.v[I], u.v, a_r.v[I]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                u.K_c.v[I][e] = u.pz_c.v[I][e] * rF;
                u.af_c.v[I][e] = fma(u.K_c.v[I][e], u.v, a_c.v[I][e]);
            }
        }
        u.ll = miss ? 0.0 : -0.5 * (kLog2Pi + log(F) + u.v * (u.v / F));  // :611-615
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    // Joseph form with L = I - K z implicit:  Y = (P - K (zP)) - ((P z^T) - K f0) K^T   (:601-604)
                    const double Ki = u.K_r.v[I], Kj = u.K_c.v[J][e];
                    const double Pij = P.v[I][J][e], Pji = ASYM ? Pt.v[I][J][e] : P.v[I][J][e];
                    const double y_ij = fma(-(fma(-Ki, f0, u.pz_r.v[I])), Kj, fma(-Ki, u.zp_c.v[J][e], Pij));
                    const double y_ji = fma(-(fma(-Kj, f0, u.pz_c.v[J][e])), Ki, fma(-Kj, u.zp_r.v[I], Pji));
                    const double khk = 0.5 * ((Ki * Hm) * Kj + (Kj * Hm) * Ki);
                    double pf = 0.5 * (y_ij + y_ji) + khk;
                    if (I == J && g == 2 * t + e) pf += jitter;  // stabilize, :536
                    u.Pf.v[I][J][e] = pf;
                }
    }
}

struct FwdConst {
    Frag Tc;   // T
    Frag C0;   // sym(R Q R^T) + eps sym(T T^T): everything of P+ that does not depend on the step (exactly symmetric)
    Frag RQR;  // sym(R Q R^T) (first step only)
    RVec z_r, c_r;
    CVec z_c;
    double H0, d0;
};

// Log-likelihood terms are deferred: lane (ts mod 32) keeps (F, v) of step ts and every 32 steps all lanes evaluate
// their log / division at once (one log per 32 steps per lane instead of one per step, replicated 32 times).
struct LlBuf {
    double F, v, acc;
    bool skip;
};
__device__ __forceinline__ void ll_capture(LlBuf& L, int lane, int ts, double F, double v, bool skip) {
    if (lane == (ts & 31)) {
        L.F = F;
        L.v = v;
        L.skip = skip;
    }
}
template <int KIND>
__device__ __forceinline__ void ll_flush(LlBuf& L, const KArgs<double>& G, int b, int ts, int lane) {
    const int my = (ts & ~31) + lane;
    if (my <= ts) {
        double ll;
        if (KIND == K_UNI) {
            const double lli = L.skip ? 0.0 : log(L.F) + L.v * L.v / L.F;  // :787
            ll = -0.5 * ((lli != 0.0 ? kLog2Pi : 0.0) + lli);             // :818
        } else {
            ll = L.skip ? 0.0 : -0.5 * (kLog2Pi + log(L.F) + L.v * (L.v / L.F));  // :611-615
        }
        L.acc += ll;
        if (G.ll) G.ll[(size_t)b * G.n + my] = ll;
    }
}

// First step (t = 0): the user's P0 may be asymmetric, so the update is evaluated exactly as the reference writes
// it (update<KIND, true>) and P_f goes through both products.
template <int KIND>
__device__ __forceinline__ void fwd_step_first(const KArgs<double>& G, const FwdConst& K, Frag& P, const Frag& Pt,
                                               RVec& a_r, CVec& a_c, LlBuf& L, int b, double yv, double* sc, int lane) {
    const int g = lane >> 2, t = lane & 3, m = G.m;
    const size_t bt = (size_t)b * G.n;
    if (G.traj) store_rec(G.traj + bt * REC, P, a_c, lane);
    if (G.pred_a) store_rvec(G.pred_a + bt * m, a_r, m, g, t);
    if (G.pred_P) store_frag(G.pred_P + bt * m * m, P, m, g, t, 1.0);
    Upd u;
    update<KIND, true>(u, P, Pt, a_r, a_c, yv, G.fill, G.jitter, K.z_r, K.z_c, K.d0, K.H0, g, t);
    ll_capture(L, lane, 0, u.F, u.v, KIND == K_UNI ? u.keep == 0.0 : u.flag);
    if (lane == 0) {
        if (G.obs_mu) G.obs_mu[bt] = u.yhat;
        if (G.obs_cov) G.obs_cov[bt] = u.F;
    }
    if (G.filt_a) store_rvec(G.filt_a + bt * m, u.af_r, m, g, t);
    if (G.filt_P) store_frag(G.filt_P + bt * m * m, u.Pf, m, g, t, 1.0);
    // predict: a+ = T a_f + c ; P+ = sym(T P_f T^T) + sym(R Q R^T)
    a_r = mv(K.Tc, u.af_c);
#pragma unroll
    for (int I = 0; I < 2; ++I) a_r.v[I] += K.c_r.v[I];
    a_c = row2col(a_r, t);
    Frag W, Y, Yt;
    mm_nt<false>(W, K.Tc, u.Pf);
    mm_nt<false>(Y, W, K.Tc);
    transpose(Yt, Y, sc, g, t);
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) P.v[I][J][e] = 0.5 * (Y.v[I][J][e] + Yt.v[I][J][e]) + K.RQR.v[I][J][e];
}

// Steps t >= 1: P is symmetric, so P_f = P - coef (P z^T)(P z^T)^T + eps I for both filters
//   Univariate: coef = 1/F (0 when skipped)                                        (:782-785, :815)
//   Standard:   coef = (F + eps)/F^2  -- the Joseph form (:601-604) with K = P z^T / F substituted
// and the prediction is evaluated as
//   P+ = sym(T P T^T) - coef (T P z^T)(T P z^T)^T + C0 ,   a+ = T a + c + (T P z^T) v / F
// which takes the two DMMA products off the scalar dependency chain (they only need P, not K or F).
template <int KIND, bool OUT>
__device__ __forceinline__ void fwd_step(const KArgs<double>& G, const FwdConst& K, Frag& P, RVec& a_r, CVec& a_c,
                                         LlBuf& L, int b, int ts, double yv, double* sc, int lane) {
    const int g = lane >> 2, t = lane & 3, m = G.m;
    const size_t bt = (size_t)b * G.n + ts;
    if (G.traj) store_rec(G.traj + bt * REC, P, a_c, lane);
    if (OUT) {
        if (G.pred_a) store_rvec(G.pred_a + bt * m, a_r, m, g, t);
        if (G.pred_P) store_frag(G.pred_P + bt * m * m, P, m, g, t, 1.0);
    }
    Frag W;
    mm_nt<false>(W, K.Tc, P);  // W = T P
    const bool miss = (KIND == K_UNI) ? isnan(yv) : (isnan(yv) || yv == G.fill);  // :792 vs :352 (quirk Q6)
    const double ym = miss ? 0.0 : yv;
    const double Hm = miss ? 0.0 : K.H0;
    RVec zm_r;
    CVec zm_c;
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        zm_r.v[I] = miss ? 0.0 : K.z_r.v[I];
        zm_c.v[I][0] = miss ? 0.0 : K.z_c.v[I][0];
        zm_c.v[I][1] = miss ? 0.0 : K.z_c.v[I][1];
    }
    const double yhat = K.d0 + dot_c(zm_c, a_c);
    const double v = ym - yhat;
    const RVec pz_r = mv(P, zm_c);
    const CVec pz_c = row2col(pz_r, t);
    const double f0 = dot_r(zm_r, pz_r);
    double F, rFk, coef;
    bool skip;
    if (KIND == K_UNI) {
        const double F0 = f0 + Hm;
        skip = (F0 == 0.0) || miss;  // :777
        F = F0 + G.jitter;
        rFk = skip ? 0.0 : 1.0 / F;
        coef = rFk;
    } else {
        F = f0 + (Hm + G.jitter);  // :598
        rFk = 1.0 / F;
        coef = (F + G.jitter) * rFk * rFk;
        skip = miss;
    }
    ll_capture(L, lane, ts, F, v, skip);
    if (G.traj) store_aux(G.traj + bt * REC, pz_c, v, F, rFk, lane);
    const double gain = rFk * v;
    if (OUT) {
        if (lane == 0) {
            if (G.obs_mu) G.obs_mu[bt] = yhat;
            if (G.obs_cov) G.obs_cov[bt] = F;
        }
        if (G.filt_a) {
            RVec af_r;
            af_r.v[0] = fma(pz_r.v[0], gain, a_r.v[0]);
            af_r.v[1] = fma(pz_r.v[1], gain, a_r.v[1]);
            store_rvec(G.filt_a + bt * m, af_r, m, g, t);
        }
        if (G.filt_P) {
            Frag Pf;
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J < 2; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double pf = fma(-coef, pz_r.v[I] * pz_c.v[J][e], P.v[I][J][e]);
                        if (I == J && g == 2 * t + e) pf += G.jitter;
                        Pf.v[I][J][e] = pf;
                    }
            store_frag(G.filt_P + bt * m * m, Pf, m, g, t, 1.0);
        }
    }
    const RVec Ta_r = mv(K.Tc, a_c);
    const RVec tp_r = mv(K.Tc, pz_c);  // T P z^T
    const CVec tp_c = row2col(tp_r, t);
#pragma unroll
    for (int I = 0; I < 2; ++I) a_r.v[I] = fma(tp_r.v[I], gain, Ta_r.v[I] + K.c_r.v[I]);
    a_c = row2col(a_r, t);
    // Y = W T^T, upper tiles; C0 rides in as the accumulator of the off-diagonal tile
    Frag Y, Yt;
    Y.v[0][0][0] = Y.v[0][0][1] = Y.v[1][1][0] = Y.v[1][1][1] = 0.0;
    Y.v[0][1][0] = K.C0.v[0][1][0];
    Y.v[0][1][1] = K.C0.v[0][1][1];
    mm_nt_upper(Y, W, K.Tc);
    transpose_upper(Yt, Y, sc, g, t);
    const double cr0 = coef * tp_r.v[0], cr1 = coef * tp_r.v[1];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        P.v[0][0][e] = fma(-cr0, tp_c.v[0][e], fma(0.5, Y.v[0][0][e] + Yt.v[0][0][e], K.C0.v[0][0][e]));
        P.v[1][1][e] = fma(-cr1, tp_c.v[1][e], fma(0.5, Y.v[1][1][e] + Yt.v[1][1][e], K.C0.v[1][1][e]));
        P.v[0][1][e] = fma(-cr0, tp_c.v[1][e], Y.v[0][1][e]);
        P.v[1][0][e] = fma(-cr1, tp_c.v[0][e], Yt.v[1][0][e]);
    }
}

// TV: c, d, Z and / or H carry a time axis (G.tv_mask; the RegressionComponent pattern Z[t], structural.py:1601-1617, and
// time-varying intercepts / measurement noise): the vectors and scalars of the step are re-read from x[t] at the top of
// every step (a few L1-resident loads; T, R, Q with a time axis stay on the generic kernels).
template <int KIND, bool OUT, int OCC, bool TV = false>
__global__ void __launch_bounds__(64, OCC) forward_kernel(KArgs<double> G) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= G.B) return;  // whole warp; the kernel only uses warp-level synchronisation
    double* sc = smem + warp * SCRATCH;
    const int m = G.m, n = G.n;

    FwdConst K;
    const double* Tg = param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
    load_frag<false>(K.Tc, Tg, m, g, t);
    rqr_frag(K.RQR, param_ptr(G.Rm, G.shared_mask, BKF_PARAM_R, b, (size_t)m * G.r),
             param_ptr(G.Q, G.shared_mask, BKF_PARAM_Q, b, (size_t)G.r * G.r), m, G.r, g, t);
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // (T T^T)[i][j]: the same products in the same order for (i,j) and (j,i) => exactly symmetric
                const int i = 8 * I + g, j = 8 * J + 2 * t + e;
                double s = 0.0;
                if (i < m && j < m)
                    for (int k = 0; k < m; ++k) s = fma(Tg[i * m + k], Tg[j * m + k], s);
                K.C0.v[I][J][e] = fma(G.jitter, s, K.RQR.v[I][J][e]);
            }
    const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)m);
    K.z_r = load_rvec(Zg, m, g);
    K.z_c = load_cvec(Zg, m, t);
    K.c_r = load_rvec(param_ptr(G.c, G.shared_mask, BKF_PARAM_C, b, (size_t)m), m, g);
    K.H0 = *param_ptr(G.H, G.shared_mask, BKF_PARAM_H, b, 1);
    K.d0 = *param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, 1);

    const double* a0g = param_ptr(G.a0, G.shared_mask, BKF_PARAM_A0, b, (size_t)m);
    const double* P0g = param_ptr(G.P0, G.shared_mask, BKF_PARAM_P0, b, (size_t)m * m);
    RVec a_r = load_rvec(a0g, m, g);
    CVec a_c = load_cvec(a0g, m, t);
    Frag P;
    load_frag<false>(P, P0g, m, g, t);
    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n;
    auto refresh = [&](int ts) {  // TV: the observation-side parameters of step ts
        const double* Zt = param_ptr_t(G.Z, G.shared_mask, G.tv_mask, BKF_PARAM_Z, b, ts, n, (size_t)m);
        K.z_r = load_rvec(Zt, m, g);
        K.z_c = load_cvec(Zt, m, t);
        K.c_r = load_rvec(param_ptr_t(G.c, G.shared_mask, G.tv_mask, BKF_PARAM_C, b, ts, n, (size_t)m), m, g);
        K.H0 = *param_ptr_t(G.H, G.shared_mask, G.tv_mask, BKF_PARAM_H, b, ts, n, 1);
        K.d0 = *param_ptr_t(G.d, G.shared_mask, G.tv_mask, BKF_PARAM_D, b, ts, n, 1);
    };
    LlBuf L;
    L.F = 1.0;
    L.v = 0.0;
    L.acc = 0.0;
    L.skip = true;
    {
        Frag Pt;
        load_frag<true>(Pt, P0g, m, g, t);
        if (TV) refresh(0);
        fwd_step_first<KIND>(G, K, P, Pt, a_r, a_c, L, b, yb[0], sc, lane);
        if (n == 1) ll_flush<KIND>(L, G, b, 0, lane);
    }
    double ynext = n > 1 ? yb[1] : 0.0;
#pragma unroll 1
    for (int ts = 1; ts < n; ++ts) {
        const double yv = ynext;
        if (ts + 1 < n) ynext = yb[ts + 1];
        if (TV) refresh(ts);
        fwd_step<KIND, OUT>(G, K, P, a_r, a_c, L, b, ts, yv, sc, lane);
        if ((ts & 31) == 31 || ts == n - 1) ll_flush<KIND>(L, G, b, ts, lane);
    }
    const double logp = warp_sum(L.acc);
    if (lane == 0) {
        if (G.logp) G.logp[b] = logp;
        if (G.status) G.status[b] = isfinite(logp) ? 0 : BKF_STATUS_NONFINITE;
    }
}

// ---------------------------------------------------------------------------------------------------
// adjoint
// ---------------------------------------------------------------------------------------------------
struct BwdConst {
    Frag TTc;  // T^T in the C layout
    RVec z_r;
    CVec z_c;
    double H0, d0;
};
struct BwdState {
    Frag G;     // S(Pbar) of the step processed last (cotangent of the prediction covariance, symmetrised)
    Frag Gsum;  // sum_t G_t  (for Rbar, Qbar)
    Frag Th;    // Tbar / 2
    RVec abar_r, gc_r, gz_r;
    CVec abar_c;
    double gH, gd;
};

// One reverse step.  On exit S.G / S.abar_* hold the cotangents of (P, a) of this step; when ASYM (t = 0) S.G is
// the un-symmetrised Pbar_0 (every entry of P0 is an independent variable).
template <int KIND, bool ASYM>
__device__ __forceinline__ void bwd_step(const KArgs<double>& G, const BwdConst& K, BwdState& S, const Frag& P,
                                         const Frag& Pt, const RVec& a_r, const CVec& a_c, double yv, double gcot,
                                         double* sc, int lane) {
    const int g = lane >> 2, t = lane & 3;
    Upd u;
    update<KIND, ASYM>(u, P, Pt, a_r, a_c, yv, G.fill, G.jitter, K.z_r, K.z_c, K.d0, K.H0, g, t);
    // ================= predict adjoint =================
    Frag W, V, Pfb;
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) S.Gsum.v[I][J][e] += S.G.v[I][J][e];
    mm_nt<false>(W, K.TTc, S.G);    // W' = T^T G
    transpose(V, W, sc, g, t);      // V = G T
    mm_nt<false>(Pfb, W, K.TTc);    // Pbar_f = T^T G T
    mm_nt<true>(S.Th, V, u.Pf);     // Tbar/2 += (G T) P_f   (P_f symmetric:  G T (P_f + P_f^T) = 2 G T P_f)
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        const double h = 0.5 * S.abar_r.v[I];  // Tbar += abar+ a_f^T
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) S.Th.v[I][J][e] = fma(h, u.af_c.v[J][e], S.Th.v[I][J][e]);
    }
    const RVec afbar_r = mv(K.TTc, S.abar_c);  // T^T abar+
#pragma unroll
    for (int I = 0; I < 2; ++I) S.gc_r.v[I] += S.abar_r.v[I];
    // ================= update adjoint =================
    if (KIND == K_UNI) {
        if (u.keep != 0.0) {  // warp-uniform
            const double llbar = -0.5 * gcot;
            const double F = u.F, v = u.v, rF = 1.0 / F;
            const RVec pk_r = mv(Pfb, u.K_c);  // Pbar_u K, Pbar_u = S(Pbar_f) ~ Pbar_f
            const double kpk = dot_r(u.K_r, pk_r);
            const double ka = dot_r(u.K_r, afbar_r);
            const double kk = fma(-2.0 * F, kpk, v * ka);  // Kbar . K
            const double Fbar = -kpk + llbar * (rF - v * v * rF * rF) - kk * rF;
            const double vbar = ka + llbar * 2.0 * v * rF;
            RVec pzbar_r;
#pragma unroll
            for (int I = 0; I < 2; ++I) {
                const double Kbar = fma(-2.0 * F, pk_r.v[I], afbar_r.v[I] * v);
                pzbar_r.v[I] = fma(Fbar, u.zm_r.v[I], Kbar * rF);
            }
            const CVec pzbar_c = row2col(pzbar_r, t);
            const RVec ppz = mv(ASYM ? Pt : P, pzbar_c);  // P^T pzbar
#pragma unroll
            for (int I = 0; I < 2; ++I) {
                S.gz_r.v[I] += fma(Fbar, u.pz_r.v[I], ppz.v[I]) - vbar * a_r.v[I];
                S.abar_r.v[I] = fma(-vbar, u.zm_r.v[I], afbar_r.v[I]);
#pragma unroll
                for (int J = 0; J < 2; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double pij = pzbar_r.v[I] * u.zm_c.v[J][e];
                        const double pji = u.zm_r.v[I] * pzbar_c.v[J][e];
                        S.G.v[I][J][e] = Pfb.v[I][J][e] + (ASYM ? pij : 0.5 * (pij + pji));
                    }
            }
            S.gH += Fbar;
            S.gd -= vbar;
        } else {  // K = 0: a and P pass through, no parameter gradient
            S.abar_r = afbar_r;
            S.G = Pfb;
        }
    } else {
        const double F = u.F, v = u.v, f0 = u.f0, Hm = u.Hm, rF = 1.0 / F;
        const double wv = v * rF;
        // xz2 = (P + P^T) z - 2 f0 K ;  u1 = G2 xz2 ;  gk = G2 K      (G2 = S(Pbar_f) ~ Pbar_f)
        CVec xz2_c;
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) xz2_c.v[J][e] = fma(-2.0 * f0, u.K_c.v[J][e], u.pz_c.v[J][e] + u.zp_c.v[J][e]);
        const RVec u1_r = mv(Pfb, xz2_c);
        const RVec gk_r = mv(Pfb, u.K_c);
        const double kgk = dot_r(u.K_r, gk_r);
        const CVec gk_c = row2col(gk_r, t);
        // Zmbar = -K^T Lbar = -((P + P^T) gk - (gk . K) (zP + P z^T))
        RVec s2g = mv(P, gk_c);
        if (ASYM) {
            const RVec s2 = mv(Pt, gk_c);
            s2g.v[0] += s2.v[0];
            s2g.v[1] += s2.v[1];
        } else {
            s2g.v[0] *= 2.0;
            s2g.v[1] *= 2.0;
        }
        double vbar = dot_r(u.K_r, afbar_r);
        double Fbar = 0.0;
        if (!u.flag) {
            Fbar = -0.5 * gcot * (rF - wv * wv);
            vbar -= gcot * wv;
        }
        // Kbar = G2 K (Hm + Hm^T) + afbar v - Lbar z ;  K = PZT / F
        RVec pztbar_r;
#pragma unroll
        for (int I = 0; I < 2; ++I) pztbar_r.v[I] = (fma(2.0 * Hm, gk_r.v[I], afbar_r.v[I] * v) - u1_r.v[I]) * rF;
        Fbar -= dot_r(u.K_r, pztbar_r);
        const double Hmbar = kgk + Fbar;
        // F = Zm PZT + Hm + eps
        RVec zmbar_r;
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            zmbar_r.v[I] = fma(Fbar, u.pz_r.v[I], -(s2g.v[I] - kgk * (u.zp_r.v[I] + u.pz_r.v[I])));
            pztbar_r.v[I] = fma(Fbar, u.zm_r.v[I], pztbar_r.v[I]);
        }
        const CVec pztbar_c = row2col(pztbar_r, t);
        const RVec ppz = mv(ASYM ? Pt : P, pztbar_c);  // PZT = P Zm^T  ->  Zmbar += P^T PZTbar
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            zmbar_r.v[I] += ppz.v[I];
            zmbar_r.v[I] = fma(-vbar, a_r.v[I], zmbar_r.v[I]);  // v = ym - d - Zm a
            S.abar_r.v[I] = fma(-vbar, u.zm_r.v[I], afbar_r.v[I]);
            // Pbar = L^T G2 L + PZTbar Zm
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double zi = u.zm_r.v[I], zj = u.zm_c.v[J][e];
                    double pb = Pfb.v[I][J][e] - zi * gk_c.v[J][e] - gk_r.v[I] * zj;
                    pb = fma(zi * zj, kgk, pb);
                    const double pij = pztbar_r.v[I] * zj;
                    const double pji = zi * pztbar_c.v[J][e];
                    S.G.v[I][J][e] = pb + (ASYM ? pij : 0.5 * (pij + pji));
                }
        }
        S.gd -= vbar;
        if (!u.flag) {
            S.gz_r.v[0] += zmbar_r.v[0];
            S.gz_r.v[1] += zmbar_r.v[1];
            S.gH += Hmbar;
        }
    }
    S.abar_c = row2col(S.abar_r, t);
}

// Reverse step for t >= 1 (P symmetric): branch-free, the DMMA chain (W', V, Pbar_f) only needs G, the update is
// recomputed from (a, P) concurrently, P_f = P - coef (P z^T)(P z^T)^T + eps I (see fwd_step).
template <int KIND>
__device__ __forceinline__ void bwd_step_sym(const KArgs<double>& G, const BwdConst& K, BwdState& S, const Frag& P,
                                             const RVec& a_r, const CVec& a_c, const Aux& X, double yv, double gcot,
                                             double* sc, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // ================= predict adjoint: products =================
    Frag W, V, Pfb;
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) S.Gsum.v[I][J][e] += S.G.v[I][J][e];
    mm_nt<false>(W, K.TTc, S.G);  // W' = T^T G
    transpose(V, W, sc, g, t);    // V = G T
    Pfb.v[0][0][0] = Pfb.v[0][0][1] = Pfb.v[0][1][0] = Pfb.v[0][1][1] = Pfb.v[1][1][0] = Pfb.v[1][1][1] = 0.0;
    mm_nt_upper(Pfb, W, K.TTc);   // Pbar_f = T^T G T (symmetric): upper tiles, then mirror the off-diagonal one
    *reinterpret_cast<double2*>(sc + g * SLD + 8 + 2 * t) = make_double2(Pfb.v[0][1][0], Pfb.v[0][1][1]);
    __syncwarp();
    Pfb.v[1][0][0] = sc[(2 * t) * SLD + 8 + g];
    Pfb.v[1][0][1] = sc[(2 * t + 1) * SLD + 8 + g];
    __syncwarp();
    // ================= recompute the update =================
    const bool miss = (KIND == K_UNI) ? isnan(yv) : (isnan(yv) || yv == G.fill);
    const double Hm = miss ? 0.0 : K.H0;
    RVec zm_r;
    CVec zm_c;
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        zm_r.v[I] = miss ? 0.0 : K.z_r.v[I];
        zm_c.v[I][0] = miss ? 0.0 : K.z_c.v[I][0];
        zm_c.v[I][1] = miss ? 0.0 : K.z_c.v[I][1];
    }
    // P z^T, v, F, 1/F come from the forward sweep's record (store_aux)
    const RVec pz_r = X.pz_r;
    const CVec pz_c = X.pz_c;
    const double v = X.v, F = X.F, rFk = X.rFk;
    const double f0 = F - (Hm + G.jitter);  // only the standard adjoint reads it
    double rF, coef, keep;
    if (KIND == K_UNI) {
        keep = rFk != 0.0 ? 1.0 : 0.0;
        rF = rFk;  // a skipped step multiplies every use by keep = 0
        coef = rFk;
    } else {
        rF = rFk;
        coef = (F + G.jitter) * rF * rF;
        keep = miss ? 0.0 : 1.0;
    }
    RVec K_r;
    CVec K_c, af_c;
    Frag Pf;
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        K_r.v[I] = pz_r.v[I] * rFk;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            K_c.v[I][e] = pz_c.v[I][e] * rFk;
            af_c.v[I][e] = fma(K_c.v[I][e], v, a_c.v[I][e]);
        }
    }
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        const double cp = coef * pz_r.v[I];
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double pf = fma(-cp, pz_c.v[J][e], P.v[I][J][e]);
                if (I == J && g == 2 * t + e) pf += G.jitter;
                Pf.v[I][J][e] = pf;
            }
    }
    // ================= predict adjoint: T, c =================
    mm_nt<true>(S.Th, V, Pf);  // Tbar/2 += (G T) P_f
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        const double h = 0.5 * S.abar_r.v[I];  // Tbar += abar+ a_f^T
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) S.Th.v[I][J][e] = fma(h, af_c.v[J][e], S.Th.v[I][J][e]);
    }
    const RVec afbar_r = mv(K.TTc, S.abar_c);  // T^T abar+
#pragma unroll
    for (int I = 0; I < 2; ++I) S.gc_r.v[I] += S.abar_r.v[I];
    // ================= update adjoint =================
    if (KIND == K_UNI) {
        const double llbar = -0.5 * gcot;
        const RVec pk_r = mv(Pfb, K_c);  // Pbar_u K, Pbar_u = S(Pbar_f) ~ Pbar_f
        const double kpk = dot_r(K_r, pk_r);
        const double ka = dot_r(K_r, afbar_r);
        const double kk = fma(-2.0 * F, kpk, v * ka);  // Kbar . K
        const double Fbar = keep * (-kpk + llbar * (rF - v * v * rF * rF) - kk * rF);
        const double vbar = keep * (ka + llbar * 2.0 * v * rF);
        RVec hr;  // pzbar / 2
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            const double Kbar = fma(-2.0 * F, pk_r.v[I], afbar_r.v[I] * v);
            hr.v[I] = 0.5 * fma(Fbar, zm_r.v[I], Kbar * rFk);
        }
        const CVec hc = row2col(hr, t);
        const RVec ppz = mv(P, hc);  // P pzbar / 2
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            S.gz_r.v[I] += fma(Fbar, pz_r.v[I], 2.0 * ppz.v[I]) - vbar * a_r.v[I];
            S.abar_r.v[I] = fma(-vbar, zm_r.v[I], afbar_r.v[I]);
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    S.G.v[I][J][e] = fma(hr.v[I], zm_c.v[J][e], fma(zm_r.v[I], hc.v[J][e], Pfb.v[I][J][e]));
        }
        S.gH += Fbar;
        S.gd -= vbar;
    } else {
        const double wv = v * rF;
        // xz2 = (P + P^T) z - 2 f0 K ;  u1 = G2 xz2 ;  gk = G2 K      (G2 = S(Pbar_f) ~ Pbar_f)
        CVec xz2_c;
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) xz2_c.v[J][e] = 2.0 * fma(-f0, K_c.v[J][e], pz_c.v[J][e]);
        const RVec u1_r = mv(Pfb, xz2_c);
        const RVec gk_r = mv(Pfb, K_c);
        const double kgk = dot_r(K_r, gk_r);
        const CVec gk_c = row2col(gk_r, t);
        const RVec pg = mv(P, gk_c);  // (P + P^T) gk = 2 P gk
        double vbar = dot_r(K_r, afbar_r);
        double Fbar = keep * (-0.5 * gcot * (rF - wv * wv));
        vbar -= keep * gcot * wv;
        // Kbar = G2 K (Hm + Hm^T) + afbar v - Lbar z ;  K = PZT / F
        RVec pztbar_r;
#pragma unroll
        for (int I = 0; I < 2; ++I) pztbar_r.v[I] = (fma(2.0 * Hm, gk_r.v[I], afbar_r.v[I] * v) - u1_r.v[I]) * rF;
        Fbar -= dot_r(K_r, pztbar_r);
        const double Hmbar = kgk + Fbar;
        RVec zmbar_r, hr;
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            // Zmbar = -K^T Lbar + Fbar PZT = -(2 P gk - 2 (gk . K) P z^T) + Fbar P z^T
            zmbar_r.v[I] = fma(Fbar + 2.0 * kgk, pz_r.v[I], -2.0 * pg.v[I]);
            hr.v[I] = 0.5 * fma(Fbar, zm_r.v[I], pztbar_r.v[I]);  // PZTbar / 2
        }
        const CVec hc = row2col(hr, t);
        const RVec ppz = mv(P, hc);  // PZT = P Zm^T  ->  Zmbar += P^T PZTbar
        const double hk = 0.5 * kgk;
        CVec qc;
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) qc.v[J][e] = fma(hk, zm_c.v[J][e], hc.v[J][e] - gk_c.v[J][e]);
#pragma unroll
        for (int I = 0; I < 2; ++I) {
            zmbar_r.v[I] = fma(-vbar, a_r.v[I], fma(2.0, ppz.v[I], zmbar_r.v[I]));  // v = ym - d - Zm a
            S.abar_r.v[I] = fma(-vbar, zm_r.v[I], afbar_r.v[I]);
            S.gz_r.v[I] = fma(keep, zmbar_r.v[I], S.gz_r.v[I]);
            // S(Pbar) = S(L^T G2 L + PZTbar Zm) = G2 + z q^T + q z^T,  q = PZTbar/2 - gk + (gk . K) z / 2
            const double qr = fma(hk, zm_r.v[I], hr.v[I] - gk_r.v[I]);
#pragma unroll
            for (int J = 0; J < 2; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    S.G.v[I][J][e] = fma(qr, zm_c.v[J][e], fma(zm_r.v[I], qc.v[J][e], Pfb.v[I][J][e]));
        }
        S.gd -= vbar;
        S.gH = fma(keep, Hmbar, S.gH);
    }
    S.abar_c = row2col(S.abar_r, t);
}

template <int KIND, bool TV = false>
__global__ void __launch_bounds__(64) backward_kernel(KArgs<double> G) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= G.B) return;
    double* sc = smem + warp * SCRATCH;
    const int m = G.m, n = G.n;

    BwdConst K;
    load_frag<true>(K.TTc, param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m), m, g, t);
    const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)m);
    K.z_r = load_rvec(Zg, m, g);
    K.z_c = load_cvec(Zg, m, t);
    K.H0 = *param_ptr(G.H, G.shared_mask, BKF_PARAM_H, b, 1);
    K.d0 = *param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, 1);

    BwdState S;
#pragma unroll
    for (int I = 0; I < 2; ++I) {
        S.abar_r.v[I] = S.gc_r.v[I] = S.gz_r.v[I] = 0.0;
#pragma unroll
        for (int J = 0; J < 2; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                S.G.v[I][J][e] = S.Gsum.v[I][J][e] = S.Th.v[I][J][e] = 0.0;
                S.abar_c.v[J][e] = 0.0;
            }
    }
    S.gH = S.gd = 0.0;

    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n;
    const double* traj = G.traj + (size_t)b * n * REC;
    const double* cot = G.ll_cot ? G.ll_cot + (size_t)b * n : nullptr;
    // TV: c, d, Z, H of step ts are read at the top of the step; the gradient of a parameter with a time axis keeps that
    // axis -- its accumulator is written out and cleared after every step instead of once at the end.
    const bool tvC = TV && ((G.tv_mask >> BKF_PARAM_C) & 1u), tvD = TV && ((G.tv_mask >> BKF_PARAM_D) & 1u),
               tvZ = TV && ((G.tv_mask >> BKF_PARAM_Z) & 1u), tvH = TV && ((G.tv_mask >> BKF_PARAM_H) & 1u);
    auto refresh = [&](int ts) {
        const double* Zt = param_ptr_t(G.Z, G.shared_mask, G.tv_mask, BKF_PARAM_Z, b, ts, n, (size_t)m);
        K.z_r = load_rvec(Zt, m, g);
        K.z_c = load_cvec(Zt, m, t);
        K.H0 = *param_ptr_t(G.H, G.shared_mask, G.tv_mask, BKF_PARAM_H, b, ts, n, 1);
        K.d0 = *param_ptr_t(G.d, G.shared_mask, G.tv_mask, BKF_PARAM_D, b, ts, n, 1);
    };
    auto flush = [&](int ts) {
        const size_t bt = (size_t)b * n + ts;
        if (tvC) {
            if (G.g_c) store_rvec(G.g_c + bt * m, S.gc_r, m, g, t);
            S.gc_r.v[0] = S.gc_r.v[1] = 0.0;
        }
        if (tvZ) {
            if (G.g_Z) store_rvec(G.g_Z + bt * m, S.gz_r, m, g, t);
            S.gz_r.v[0] = S.gz_r.v[1] = 0.0;
        }
        if (tvD) {
            if (G.g_d && lane == 0) G.g_d[bt] = S.gd;
            S.gd = 0.0;
        }
        if (tvH) {
            if (G.g_H && lane == 0) G.g_H[bt] = S.gH;
            S.gH = 0.0;
        }
    };
    Frag P;
    RVec a_r;
    CVec a_c;
    {
        Frag Pn;
        RVec an_r;
        CVec an_c;
        Aux Xn;
        if (n > 1) {
            load_rec(traj + (size_t)(n - 1) * REC, Pn, an_r, an_c, lane);
            load_aux(traj + (size_t)(n - 1) * REC, Xn, lane);
        }
#pragma unroll 1
        for (int ts = n - 1; ts >= 1; --ts) {
            P = Pn;
            a_r = an_r;
            a_c = an_c;
            const Aux X = Xn;
            if (ts > 1) {  // prefetch one step ahead
                load_rec(traj + (size_t)(ts - 1) * REC, Pn, an_r, an_c, lane);
                load_aux(traj + (size_t)(ts - 1) * REC, Xn, lane);
            }
            if (TV) refresh(ts);
            bwd_step_sym<KIND>(G, K, S, P, a_r, a_c, X, yb[ts], cot ? cot[ts] : 1.0, sc, lane);
            if (TV) flush(ts);
        }
    }
    {
        const double* a0g = param_ptr(G.a0, G.shared_mask, BKF_PARAM_A0, b, (size_t)m);
        const double* P0g = param_ptr(G.P0, G.shared_mask, BKF_PARAM_P0, b, (size_t)m * m);
        Frag Pt;
        load_frag<false>(P, P0g, m, g, t);
        load_frag<true>(Pt, P0g, m, g, t);
        a_r = load_rvec(a0g, m, g);
        a_c = load_cvec(a0g, m, t);
        if (TV) refresh(0);
        bwd_step<KIND, true>(G, K, S, P, Pt, a_r, a_c, yb[0], cot ? cot[0] : 1.0, sc, lane);
        if (TV) flush(0);
    }
    // ---- gradients out ----
    if (G.g_a0) store_rvec(G.g_a0 + (size_t)b * m, S.abar_r, m, g, t);
    if (G.g_c && !tvC) store_rvec(G.g_c + (size_t)b * m, S.gc_r, m, g, t);
    if (G.g_Z && !tvZ) store_rvec(G.g_Z + (size_t)b * m, S.gz_r, m, g, t);
    if (lane == 0) {
        if (G.g_d && !tvD) G.g_d[b] = S.gd;
        if (G.g_H && !tvH) G.g_H[b] = S.gH;
    }
    if (G.g_P0) store_frag(G.g_P0 + (size_t)b * m * m, S.G, m, g, t, 1.0);
    if (G.g_T) store_frag(G.g_T + (size_t)b * m * m, S.Th, m, g, t, 2.0);
    // Rbar = Gsum R (Q + Q^T) ; Qbar = R^T Gsum R   (sym(R Q R^T) is time-invariant: cotangents summed first)
    if (G.g_R || G.g_Q) {
        const int r = G.r;
        const double* Rg = param_ptr(G.Rm, G.shared_mask, BKF_PARAM_R, b, (size_t)m * r);
        const double* Qg = param_ptr(G.Q, G.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J < 2; ++J)
                *reinterpret_cast<double2*>(sc + (8 * I + g) * SLD + 8 * J + 2 * t) =
                    make_double2(S.Gsum.v[I][J][0], S.Gsum.v[I][J][1]);
        __syncwarp();
        if (G.g_R) {
            for (int idx = lane; idx < m * r; idx += 32) {
                const int i = idx / r, jj = idx - i * r;
                double acc = 0;
                for (int k = 0; k < r; ++k) {
                    double gr = 0;
                    for (int l = 0; l < m; ++l) gr = fma(sc[i * SLD + l], Rg[l * r + k], gr);
                    acc = fma(gr, Qg[k * r + jj] + Qg[jj * r + k], acc);
                }
                G.g_R[(size_t)b * m * r + idx] = acc;
            }
        }
        if (G.g_Q) {
            for (int idx = lane; idx < r * r; idx += 32) {
                const int i = idx / r, jj = idx - i * r;
                double acc = 0;
                for (int k = 0; k < m; ++k) {
                    double gr = 0;
                    for (int l = 0; l < m; ++l) gr = fma(sc[k * SLD + l], Rg[l * r + jj], gr);
                    acc = fma(Rg[k * r + i], gr, acc);
                }
                G.g_Q[(size_t)b * r * r + idx] = acc;
            }
        }
    }
}

}  // namespace tc
}  // namespace bkf
