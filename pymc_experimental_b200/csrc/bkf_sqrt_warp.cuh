// bkf_sqrt_warp.cuh -- SquareRootFilter (kalman_filter.py:620-750) on FP64 tensor-core tiles, ONE WARP PER SERIES.
//
// Same arithmetic as bkf_sqrt.cu (CTA per series) -- Householder QR with LAPACK's dgeqr2/dlarfg sign rules, the Q-free
// QR pull-back -- but organised so that a warp never waits for another one:
//
//   * every matrix lives in the warp's shared-memory arena as 8 x 8 tiles in DMMA accumulator order ("C layout":
//     lane 4g+t owns elements [g][2t], [g][2t+1] of a tile, one 16-byte word), so a tile load/store is one conflict-free
//     512-byte access and a loaded tile is directly an operand of   D += X Y^T   (mma_nt: two DMMA.8x8x4 whose A / B
//     operands are the accumulator registers of X / Y, see bkf_mma.cuh);
//   * the arrays that are factored are stored TRANSPOSED (tile row = column block of the array): the 8-column panel of
//     a Householder sweep is then one tile row, a lane group g owns column g of the panel, the reflector is broadcast
//     with one shuffle per element and all inner products of a reflector step are formed in one reduction window; the
//     compact-WY factor T is built on the fly from the same window (dlarft), and the trailing update
//     A^T <- A^T - (A^T V) T V^T  is three chains of mma_nt per column block;
//   * the update array [[Hc, Zm Pc], [0, Pc]] and the predict array [T Pcf, R Qc] share one tile array: the factor of
//     one is written where the next one reads it (Pc never moves).
//
// Shapes: float64, m, p, r multiples of 8 with r <= p (instantiated for the shapes in bkf_sqrt_warp.cu).  Everything else
// takes the CTA-per-series kernels.  The forward record kept for the adjoint is the one of bkf_sqrt.cu
// ([a | Pc | triu(Rf) | triu(Rp)]), so either family's adjoint can read either family's forward sweep.
#pragma once
#include "bkf_common.cuh"

namespace bkf {
namespace sqw {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
// D += X Y^T on 8 x 8 tiles in C layout
__device__ __forceinline__ void mma_nt(double2& d, const double2 x, const double2 y) {
    dmma(d.x, d.y, x.x, y.x);
    dmma(d.x, d.y, x.y, y.y);
}
// Accumulator tile of the adjoint's product chains.  (Splitting it into one accumulator per k-half -- two independent
// DMMAs instead of a dependent pair -- was measured: no gain at one warp per scheduler, more spills.)
struct Acc2 {
    double2 a;
};
__device__ __forceinline__ Acc2 acc2_zero() { return Acc2{make_double2(0.0, 0.0)}; }
__device__ __forceinline__ Acc2 acc2_from(double2 v) { return Acc2{v}; }
__device__ __forceinline__ void mma_nt(Acc2& d, const double2 x, const double2 y) {
    dmma(d.a.x, d.a.y, x.x, y.x);
    dmma(d.a.x, d.a.y, x.y, y.y);
}
__device__ __forceinline__ double2 fin(const Acc2& d) { return d.a; }
__device__ __forceinline__ double2 ldt(const double* S, int tile, int lane) {
    return reinterpret_cast<const double2*>(S)[tile * 32 + lane];
}
__device__ __forceinline__ void stt(double* S, int tile, int lane, double2 v) {
    reinterpret_cast<double2*>(S)[tile * 32 + lane] = v;
}
__device__ __forceinline__ double2 neg(double2 v) { return make_double2(-v.x, -v.y); }
__device__ __forceinline__ double2 zero2() { return make_double2(0.0, 0.0); }

// transpose of a tile in C layout: two shuffles (each lane sends one of its two values per round)
__device__ __forceinline__ double2 ttrans(double2 v, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const bool odd = g & 1;
    const int s1 = 4 * (2 * t + (odd ? 1 : 0)) + (g >> 1);
    const int s2 = 4 * (2 * t + (odd ? 0 : 1)) + (g >> 1);
    const double r1 = __shfl_sync(FULL, odd ? v.y : v.x, s1);
    const double r2 = __shfl_sync(FULL, odd ? v.x : v.y, s2);
    return odd ? make_double2(r2, r1) : make_double2(r1, r2);
}
// sum over the four lanes of a group (same g)
__device__ __forceinline__ double red_t(double x) {
    x += __shfl_xor_sync(FULL, x, 1);
    x += __shfl_xor_sync(FULL, x, 2);
    return x;
}
// sum over the eight groups (same t)
__device__ __forceinline__ double red_g(double x) {
    x += __shfl_xor_sync(FULL, x, 4);
    x += __shfl_xor_sync(FULL, x, 8);
    x += __shfl_xor_sync(FULL, x, 16);
    return x;
}
__device__ __forceinline__ double warp_sum_d(double x) { return red_g(red_t(x)); }

// tile (I, J) of a row-major global matrix (leading dimension ld, even) in C layout, and of its transpose
__device__ __forceinline__ double2 ldg_tile(const double* __restrict__ M, int ld, int I, int J, int lane) {
    const int g = lane >> 2, t = lane & 3;
    return *reinterpret_cast<const double2*>(M + (size_t)(8 * I + g) * ld + 8 * J + 2 * t);
}
__device__ __forceinline__ double2 ldg_tile_t(const double* __restrict__ M, int ld, int I, int J, int lane) {
    const int g = lane >> 2, t = lane & 3;  // out[g][2t+e] = M[8J + 2t + e][8I + g]
    const double* q = M + (size_t)(8 * J + 2 * t) * ld + 8 * I + g;
    return make_double2(q[0], q[ld]);
}
// M[tile] += v without reading it back (RED: the add happens in L2, nothing waits for it; every element has one owner)
__device__ __forceinline__ void red_tile(double* M, int ld, int I, int J, int lane, double2 v) {
    const int g = lane >> 2, t = lane & 3;
    double* q = M + (size_t)(8 * I + g) * ld + 8 * J + 2 * t;
    atomicAdd(q, v.x);
    atomicAdd(q + 1, v.y);
}
__device__ __forceinline__ void stg_tile(double* __restrict__ M, int ld, int I, int J, int lane, double2 v) {
    const int g = lane >> 2, t = lane & 3;
    *reinterpret_cast<double2*>(M + (size_t)(8 * I + g) * ld + 8 * J + 2 * t) = v;
}
// address of element (i, j) of a tile matrix with WT tile columns
__device__ __forceinline__ int elem_addr(int i, int j, int WT) {
    return ((i >> 3) * WT + (j >> 3)) * 64 + ((i & 7) * 4 + ((j & 7) >> 1)) * 2 + (j & 1);
}

// ---------------------------------------------------------------------------------------------------------------
// One panel (8 columns, NQ active row blocks, the diagonal block first) of the Householder sweep and its trailing update.
// tile(jt, rt) -> tile index.  The reflector loop is rolled (runtime k): see qr_tiles.
template <int NQ, int WT, class ColF>
__device__ __forceinline__ void qr_panel(double* __restrict__ S, double* __restrict__ tq, const int jt, const int np,
                                         const int row0, ColF colf, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    // tile (column block j, active row block q) = (row0 + j) * WT + col[q]: the column part is fixed for the panel
    int col[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) col[q] = colf(jt + q);
    auto tile = [&](int j, int rt) { return (row0 + j) * WT + col[rt - jt]; };
    double2 pa[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) pa[q] = ldt(S, tile(jt, jt + q), lane);
    double tauv = 0.0, scalv = 1.0;  // tau and 1 / (alpha - beta) of the reflector of column g (tau = 0: H = I)
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
        // x = column k of the panel below its diagonal, handed to every group (lane (k, t) -> lanes (*, t))
        double2 xb[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            xb[q].x = __shfl_sync(FULL, pa[q].x, 4 * k + t);
            xb[q].y = __shfl_sync(FULL, pa[q].y, 4 * k + t);
        }
        double2 xm = xb[0];
        if (2 * t <= k) xm.x = 0.0;
        if (2 * t + 1 <= k) xm.y = 0.0;
        double d0 = xm.x * pa[0].x, d1 = xm.y * pa[0].y;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            d0 = fma(xb[q].x, pa[q].x, d0);
            d1 = fma(xb[q].y, pa[q].y, d1);
        }
        // one reduction window: group k -> ss = x.x; group c > k -> x.a_c
        const double dot = red_t(d0 + d1);
        const double ss = __shfl_sync(FULL, dot, 4 * k);
        const double akc = __shfl_sync(FULL, (k & 1) ? pa[0].y : pa[0].x, 4 * g + (k >> 1));  // [row k][col g]
        const double alpha = __shfl_sync(FULL, akc, 4 * k);
        if (ss == 0.0) continue;  // dlarfg: tau = 0, H = I
        // beta = -sign(alpha) nrm,  tau = (beta - alpha) / beta = (|alpha| + nrm) / nrm,  scal = 1 / (alpha - beta).
        // 1 / sqrt and 1 / x come from the 20-bit hardware seeds and two Newton steps each (a quarter of the instructions
        // and half the latency of the IEEE sequences: measured dependent latencies sqrt 91, div 71 cycles).
        // Serial chain of the reflector, kept as short as it goes: the reciprocal of |alpha| + nrm starts from a seed
        // taken at the rsqrt SEED's accuracy (both seeds carry ~20 bits, two Newton steps each reach full precision), so
        // the second hardware approximation overlaps the Newton steps of the first instead of waiting for nrm.
        const double x2 = fma(alpha, alpha, ss);  // (x2 subnormal, a column of magnitude < 1e-154, is out of range)
        const double aa = fabs(alpha);
        double rn, r0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rn) : "d"(x2));
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(fma(x2, rn, aa)));
        const double hx = 0.5 * x2;
        rn = fma(rn, fma(-hx * rn, rn, 0.5), rn);
        rn = fma(rn, fma(-hx * rn, rn, 0.5), rn);
        const double nrm = x2 * rn;
        const double den = aa + nrm;
        r0 = fma(r0, fma(-den, r0, 1.0), r0);
        const double rden = fma(r0, fma(-den, r0, 1.0), r0);
        const double beta = copysign(nrm, -alpha);
        const double tau = den * rn;
        const double scal = copysign(rden, alpha);
        const double ts = tau * scal;
        const double zz = fma(scal, dot, akc);  // v_k . a_c  (c > k)
        // Update with the UNSCALED reflector u = [alpha - beta; x] (v = u scal):  a_c -= (tau zz scal) u  for c > k.
        // The pivot column keeps x until the sweep of the panel is over (nothing reads v before) and is scaled once.
        const bool piv = g == k;
        const double w = g > k ? ts * zz : 0.0;
        double2 ud = xm;  // u on my rows of the diagonal block
        if (2 * t == k) ud.x = alpha - beta;
        if (2 * t + 1 == k) ud.y = alpha - beta;
        {
            const double nx = fma(-w, ud.x, pa[0].x), ny = fma(-w, ud.y, pa[0].y);
            pa[0].x = (piv && 2 * t == k) ? beta : nx;
            pa[0].y = (piv && 2 * t + 1 == k) ? beta : ny;
        }
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            pa[q].x = fma(-w, xb[q].x, pa[q].x);
            pa[q].y = fma(-w, xb[q].y, pa[q].y);
        }
        if (piv) {
            tauv = tau;
            scalv = scal;
        }
    }
    // R goes back to the array (zeros below the diagonal); V^T = (x scal)^T with unit diagonal stays in registers (in pa)
    {
        const double2 r = pa[0];
        stt(S, tile(jt, jt), lane, make_double2(2 * t <= g ? r.x : 0.0, 2 * t + 1 <= g ? r.y : 0.0));
        pa[0].x = 2 * t > g ? r.x * scalv : (2 * t == g ? 1.0 : 0.0);
        pa[0].y = 2 * t + 1 > g ? r.y * scalv : (2 * t + 1 == g ? 1.0 : 0.0);
    }
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
        pa[q].x *= scalv;
        pa[q].y *= scalv;
        stt(S, tile(jt, jt + q), lane, zero2());
    }
    if (jt + 1 < np) {
        // Compact-WY factor (dlarft) from the Gram matrix of the reflectors:  T^{-1} = striu(V^T V) + diag(1 / tau), so
        // column j of T solves  x_i = tau_i (delta_ij - sum_{l > i} G_il x_l)  by back substitution -- no division, a
        // reflector with tau = 0 gives a zero row and column.  G = V^T V is two short DMMA chains; group j solves column j
        // with broadcast reads of G (all lanes read the same word).  Doing this once per panel instead of one
        // three-stage shuffle reduction per reflector takes the T factor off the serial reflector chain.
        double2 tt;  // T^T in C layout: lane (g = j, t) holds T[2t][j], T[2t+1][j]
        {
            double2 ga = zero2(), gb = zero2();
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                dmma(ga.x, ga.y, pa[q].x, pa[q].x);
                dmma(gb.x, gb.y, pa[q].y, pa[q].y);
            }
            reinterpret_cast<double2*>(tq)[lane] = make_double2(ga.x + gb.x, ga.y + gb.y);  // G, row-major 8 x 8
            if (t == 0) tq[64 + g] = tauv;
            __syncwarp();
            double x[8];
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                double acc = 0.0;
#pragma unroll
                for (int l = 7; l > i; --l) acc = fma(tq[i * 8 + l], x[l], acc);
                const double ti = tq[64 + i];
                x[i] = i == g ? ti : (i < g ? -ti * acc : 0.0);
            }
            tt.x = t == 0 ? x[0] : (t == 1 ? x[2] : (t == 2 ? x[4] : x[6]));
            tt.y = t == 0 ? x[1] : (t == 1 ? x[3] : (t == 2 ? x[5] : x[7]));
            __syncwarp();
        }
        double2 vv[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) vv[q] = ttrans(pa[q], lane);
#pragma unroll 1
        for (int j2 = jt + 1; j2 < np; ++j2) {
            double2 c[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) c[q] = ldt(S, tile(j2, jt + q), lane);
            // (A^T V) block: independent accumulator chains (even / odd q, first / second k-half)
            double2 wa = zero2(), wb = zero2(), wc = zero2(), wd = zero2();
#pragma unroll
            for (int q = 0; q < NQ; q += 2) {
                dmma(wa.x, wa.y, c[q].x, pa[q].x);
                if (q + 1 < NQ) dmma(wc.x, wc.y, c[q + 1 < NQ ? q + 1 : q].x, pa[q + 1 < NQ ? q + 1 : q].x);
                dmma(wb.x, wb.y, c[q].y, pa[q].y);
                if (q + 1 < NQ) dmma(wd.x, wd.y, c[q + 1 < NQ ? q + 1 : q].y, pa[q + 1 < NQ ? q + 1 : q].y);
            }
            const double2 wt = make_double2((wa.x + wb.x) + (wc.x + wd.x), (wa.y + wb.y) + (wc.y + wd.y));
            double2 u0 = zero2(), u1 = zero2();  // (A^T V) T
            dmma(u0.x, u0.y, wt.x, tt.x);
            dmma(u1.x, u1.y, wt.y, tt.y);
            const double2 ut = make_double2(-(u0.x + u1.x), -(u0.y + u1.y));
            // A^T -= (A^T V T) V^T : first k-halves of every tile, then the second halves
#pragma unroll
            for (int q = 0; q < NQ; ++q) dmma(c[q].x, c[q].y, ut.x, vv[q].x);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                dmma(c[q].x, c[q].y, ut.y, vv[q].y);
                stt(S, tile(j2, jt + q), lane, c[q]);
            }
        }
    }
    __syncwarp();
}

// Householder R factor of an array stored transposed in tiles (element [col][row] in C layout).  Two arrays share the
// tile grid (WT tile columns):
//   update  (predict = false): column block jt -> tile row jt,      row block rt -> tile column rt
//   predict (predict = true ): column block jt -> tile row PT + jt, row block rt -> tile column PT + rt (rt < MT),
//                              rt - MT (the R Qc rows) -- so that R^T (= the next Pc) lands on the Pc block.
// np column blocks, nr row blocks (nr <= NRMAX).  On return the tiles hold R^T, zeros elsewhere.
// One copy of the code per number of active row blocks, panel and reflector loops rolled: a fully unrolled sweep is
// ~200 KB of SASS, which the 32 KB instruction cache cannot hold with one warp per scheduler.
template <int NRMAX, int MT, int PT, int WT>
__device__ __noinline__ void qr_tiles(double* __restrict__ S, double* __restrict__ tq, const int np, const int nr,
                                      const bool predict, const int lane) {
    const int row0 = predict ? PT : 0;
    auto colf = [&](int rt) { return predict ? (rt < MT ? PT + rt : rt - MT) : rt; };
#pragma unroll 1
    for (int jt = 0; jt < np; ++jt) {
        switch (nr - jt) {
            case 1: qr_panel<1, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 2: if constexpr (NRMAX >= 2) qr_panel<2, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 3: if constexpr (NRMAX >= 3) qr_panel<3, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 4: if constexpr (NRMAX >= 4) qr_panel<4, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 5: if constexpr (NRMAX >= 5) qr_panel<5, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 6: if constexpr (NRMAX >= 6) qr_panel<6, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 7: if constexpr (NRMAX >= 7) qr_panel<7, WT>(S, tq, jt, np, row0, colf, lane); break;
            case 8: if constexpr (NRMAX >= 8) qr_panel<8, WT>(S, tq, jt, np, row0, colf, lane); break;
            default: break;
        }
    }
}

// Doubles per (series, step) record: identical to sq::traj_record_elems (bkf_sqrt.cu)
__host__ __device__ inline size_t traj_record_elems(int m, int p) {
    const size_t N1 = (size_t)p + m;
    return (size_t)m + (size_t)m * m + N1 * (N1 + 1) / 2 + (size_t)m * (m + 1) / 2;
}

template <int MT, int PT, int RT>
struct Dims {
    static constexpr int mt = MT, pt = PT, rt = RT;
    static constexpr int m = 8 * MT, p = 8 * PT, r = 8 * RT, N1 = m + p;
    static constexpr int NT1 = MT + PT;  // tile rows / columns of the shared array
    static constexpr int WT = NT1;       // tile columns (RT <= PT so that the predict array fits)
    static_assert(RT <= PT, "predict array [T Pcf, R Qc] must fit in the update array");
    // update array A^T = [[Hc, Zm Pc], [0, Pc]]: column block jt -> tile row jt, row block rt -> tile column rt
    struct MapU {
        __host__ __device__ static constexpr int tile(int jt, int rt) { return jt * WT + rt; }
    };
    // predict array M^T = [T Pcf, R Qc]: its R^T (= next Pc) must land on the Pc block -> tile rows PT + jt,
    // row blocks 0..MT-1 on tile columns PT.., row blocks MT.. (the R Qc part) on tile columns 0..RT-1
    struct MapP {
        __host__ __device__ static constexpr int tile(int jt, int rt) {
            return (PT + jt) * WT + (rt < MT ? PT + rt : rt - MT);
        }
    };
    static constexpr int pc_tile(int I, int J) { return (PT + I) * WT + PT + J; }
    // forward arena (doubles)
    static constexpr int SA = 0;
    static constexpr int RQ = SA + NT1 * WT * 64;   // R Qc, MT x RT tiles
    static constexpr int HC = RQ + MT * RT * 64;    // Hc, PT x PT tiles
    static constexpr int VA = HC + PT * PT * 64;    // vectors
    static constexpr int V_a = VA, V_af = V_a + m, V_c = V_af + m, V_d = V_c + m, V_v = V_d + p, V_sv = V_v + p,
                         V_in = V_sv + p, V_w = V_in + p, V_ym = V_w + p, V_tq = V_ym + p;  // V_tq: Gram tile + taus (72)
    static constexpr int FWD_ELEMS = V_tq + 72;
};

// In-place lower Cholesky of a k x k row-major matrix in shared memory by one warp; the upper triangle is zeroed.
__device__ inline bool chol_lower_warp(double* A, int k, int lane) {
    bool ok = true;
    for (int j = 0; j < k; ++j) {
        double s = A[j * k + j];
        for (int q = 0; q < j; ++q) s -= A[j * k + q] * A[j * k + q];
        if (!(s > 0)) { ok = false; s = NAN; }
        const double ljj = sqrt(s);
        __syncwarp();
        for (int i = j + 1 + lane; i < k; i += 32) {
            double tv = A[i * k + j];
            for (int q = 0; q < j; ++q) tv -= A[i * k + q] * A[j * k + q];
            A[i * k + j] = tv / ljj;
        }
        if (lane == 0) A[j * k + j] = ljj;
        for (int i = lane; i < j; i += 32) A[i * k + j] = 0.0;
        __syncwarp();
    }
    return ok;
}

// Per-series constants: Hc = chol(H) (or H when all zero, :666) and R Qc as tiles.  scratch: >= p*p + r*r + m*r doubles.
template <class D>
__device__ inline int setup_consts(const KArgs<double>& G, int b, double* HCt, double* RQt, double* Qc_rowmajor,
                                   double* scratch, bool* Hzero_out, int lane) {
    constexpr int m = D::m, p = D::p, r = D::r;
    const double* Hg = param_ptr(G.H, G.shared_mask, BKF_PARAM_H, b, (size_t)p * p);
    const double* Qg = param_ptr(G.Q, G.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
    const double* Rg = param_ptr(G.Rm, G.shared_mask, BKF_PARAM_R, b, (size_t)m * r);
    double* Hc = scratch;
    double* Qc = Qc_rowmajor ? Qc_rowmajor : scratch + p * p;
    bool nz = false;
    for (int i = lane; i < p * p; i += 32) {
        const double h = Hg[i];
        Hc[i] = h;
        nz = nz || !(h == 0.0);
    }
    for (int i = lane; i < r * r; i += 32) Qc[i] = Qg[i];
    const bool Hzero = !__any_sync(FULL, nz);
    __syncwarp();
    int st = 0;
    if (!Hzero && !chol_lower_warp(Hc, p, lane)) st |= BKF_STATUS_NOT_PD;
    if (!chol_lower_warp(Qc, r, lane)) st |= BKF_STATUS_NOT_PD;
    for (int idx = lane; idx < p * p; idx += 32) HCt[elem_addr(idx / p, idx % p, D::pt)] = Hc[idx];
    if (RQt)
        for (int idx = lane; idx < m * r; idx += 32) {
            const int i = idx / r, k = idx - i * r;
            double acc = 0;
            for (int l = k; l < r; ++l) acc = fma(Rg[i * r + l], Qc[l * r + k], acc);  // Qc lower: l >= k
            RQt[elem_addr(i, k, D::rt)] = acc;
        }
    __syncwarp();
    *Hzero_out = Hzero;
    return st;
}

// Packed upper triangle (row i holds n - i entries) <- R, where the tiles hold R^T: tile (I, J), I >= J, element
// [g][2t+e] = R[8J + 2t + e][8I + g].
template <int NT, class TileOf>
__device__ __forceinline__ void store_packed_upper(double* __restrict__ dst, const double* __restrict__ S, int lane,
                                                   TileOf tile_of) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int n = 8 * NT;
#pragma unroll
    for (int I = 0; I < NT; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            const double2 v = ldt(S, tile_of(I, J), lane);
            const int j = 8 * I + g;
            int i = 8 * J + 2 * t;
            if (j >= i) dst[i * n - (i * (i - 1)) / 2 + (j - i)] = v.x;
            ++i;
            if (j >= i) dst[i * n - (i * (i - 1)) / 2 + (j - i)] = v.y;
        }
}

// dst (row-major n x n, n = 8 NT) = L L^T for a factor given as tiles L(I, K)  (_postprocess_scan_results, :733-740)
template <int NT, class TileF>
__device__ __forceinline__ void store_square(double* __restrict__ dst, TileF L, int lane) {
#pragma unroll
    for (int I = 0; I < NT; ++I) {
        double2 li[NT];
#pragma unroll
        for (int K = 0; K < NT; ++K) li[K] = L(I, K);
#pragma unroll
        for (int J = 0; J < NT; ++J) {
            double2 acc = zero2();
#pragma unroll
            for (int K = 0; K < NT; ++K) mma_nt(acc, li[K], L(J, K));
            stg_tile(dst, 8 * NT, I, J, lane, acc);
        }
    }
}

template <int MT, int PT, int RT>
__global__ void __launch_bounds__(32) forward_kernel(KArgs<double> G) {
    using D = Dims<MT, PT, RT>;
    constexpr int m = D::m, p = D::p, N1 = D::N1, WT = D::WT;
    extern __shared__ double smem[];
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    const int n = G.n;
    double* SA = smem + D::SA;
    double* RQ = smem + D::RQ;
    double* HC = smem + D::HC;
    double *va = smem + D::V_a, *vaf = smem + D::V_af, *vc = smem + D::V_c, *vd = smem + D::V_d, *vv = smem + D::V_v,
           *vsv = smem + D::V_sv, *vin = smem + D::V_in, *vw = smem + D::V_w, *vym = smem + D::V_ym;

    bool Hzero;
    int st = setup_consts<D>(G, b, HC, RQ, nullptr, SA, &Hzero, lane);
    const double* Tg = param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
    const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)p * m);
    {
        const double* a0 = param_ptr(G.a0, G.shared_mask, BKF_PARAM_A0, b, (size_t)m);
        const double* P0 = param_ptr(G.P0, G.shared_mask, BKF_PARAM_P0, b, (size_t)m * m);
        const double* cg = param_ptr(G.c, G.shared_mask, BKF_PARAM_C, b, (size_t)m);
        const double* dg = param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, (size_t)p);
        for (int i = lane; i < m; i += 32) { va[i] = a0[i]; vc[i] = cg[i]; }
        for (int i = lane; i < p; i += 32) vd[i] = dg[i];
        __syncwarp();  // the Cholesky scratch in SA is dead from here on
#pragma unroll
        for (int I = 0; I < MT; ++I)
#pragma unroll
            for (int J = 0; J < MT; ++J) stt(SA, D::pc_tile(I, J), lane, ldg_tile(P0, m, I, J, lane));
        __syncwarp();
    }
    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n * p;
    const size_t tstride = traj_record_elems(m, p);
    const size_t off_rf = (size_t)m + (size_t)m * m, off_rp = off_rf + (size_t)N1 * (N1 + 1) / 2;
    double logp = 0.0;
    for (int ts = 0; ts < n; ++ts) {
        const size_t bt = (size_t)b * n + ts;
        double* rec = G.traj ? G.traj + bt * tstride : nullptr;
        // ---- mask, innovation ----
        bool miss = true;
        if (lane < p) {
            const double yv = yb[(size_t)ts * p + lane];
            miss = is_missing(yv, G.fill, true);
            vw[lane] = miss ? 0.0 : 1.0;
            vym[lane] = miss ? 0.0 : yv;
        }
        const int nobs = __popc(__ballot_sync(FULL, !miss));
        const bool flag = nobs == 0, partial = !flag && nobs != p;
        __syncwarp();
        if (rec) {
            for (int i = lane; i < m; i += 32) rec[i] = va[i];
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < MT; ++J) stg_tile(rec + m, m, I, J, lane, ldt(SA, D::pc_tile(I, J), lane));
        }
        if (G.pred_a)
            for (int i = lane; i < m; i += 32) G.pred_a[bt * m + i] = va[i];
        if (G.pred_P)
            store_square<MT>(G.pred_P + bt * m * m, [&](int I, int K) { return ldt(SA, D::pc_tile(I, K), lane); }, lane);
        double2 zf[PT][MT];
#pragma unroll
        for (int I = 0; I < PT; ++I) {
            const double wI = vw[8 * I + g];
            double acc = 0.0;
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                double2 z = ldg_tile(Zg, m, I, J, lane);
                z.x *= wI;
                z.y *= wI;
                zf[I][J] = z;
                acc = fma(z.x, va[8 * J + 2 * t], acc);
                acc = fma(z.y, va[8 * J + 2 * t + 1], acc);
            }
            acc = red_t(acc) + vd[8 * I + g];
            if (t == 0) {
                vsv[8 * I + g] = acc;                  // y_hat
                vv[8 * I + g] = vym[8 * I + g] - acc;  // v
            }
        }
        __syncwarp();
        if (G.obs_mu && lane < p) G.obs_mu[bt * p + lane] = vsv[lane];
        if (G.obs_cov && flag)  // every observation missing: the masked array has zero columns there, F_chol = 0
            for (int i = lane; i < p * p; i += 32) G.obs_cov[bt * p * p + i] = 0.0;
        double ll = 0.0;
        if (!flag) {
            // ---- update array (transposed): [[Hc, Zm Pc], [0, Pc]] ----
#pragma unroll
            for (int I = 0; I < PT; ++I)
#pragma unroll
                for (int J = 0; J < PT; ++J) {
                    double2 h = ldt(HC, I * PT + J, lane);
                    if (Hzero) h = zero2();
                    else if (partial) h = make_double2(NAN, NAN);  // the reference's cholesky(W H) fails here
                    stt(SA, I * WT + J, lane, h);
                }
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                double2 acc[PT];
#pragma unroll
                for (int I = 0; I < PT; ++I) acc[I] = zero2();
#pragma unroll
                for (int K = 0; K < MT; ++K) {
                    const double2 pt = ttrans(ldt(SA, D::pc_tile(K, J), lane), lane);  // (Pc[K][J])^T
#pragma unroll
                    for (int I = 0; I < PT; ++I) mma_nt(acc[I], zf[I][K], pt);
                }
#pragma unroll
                for (int I = 0; I < PT; ++I) stt(SA, I * WT + PT + J, lane, acc[I]);
            }
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < PT; ++J) stt(SA, (PT + I) * WT + J, lane, zero2());
            __syncwarp();
            qr_tiles<D::NT1, MT, PT, WT>(SA, smem + D::V_tq, D::NT1, D::NT1, false, lane);
            if (rec)
                store_packed_upper<D::NT1>(rec + off_rf, SA, lane, [](int I, int J) { return I * WT + J; });
            if (G.obs_cov)  // F = Fc Fc^T, Fc = B[:p,:p]
                store_square<PT>(G.obs_cov + bt * p * p, [&](int I, int K) { return ldt(SA, I * WT + K, lane); }, lane);
            // ---- s = Fc^{-1} v, inner = Fc^{-1} s  (Fc = B[:p,:p], lower), lane q owns x[q] ----
            {
                const int q = lane < p ? lane : p - 1;
                double frow[p];
#pragma unroll
                for (int i = 0; i < p; ++i) frow[i] = SA[elem_addr(q, i, WT)];
                const double dq = SA[elem_addr(q, q, WT)];
                const double rdiag = 1.0 / dq;
                double x = lane < p ? vv[lane] : 0.0;
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
                    for (int i = 0; i < p; ++i) {
                        const double xi = __shfl_sync(FULL, x, i) * __shfl_sync(FULL, rdiag, i);
                        if (lane == i) x = xi;
                        else if (lane > i) x = fma(-frow[i], xi, x);
                    }
                    if (lane < p) (rep == 0 ? vsv : vin)[lane] = x;
                }
                const double quad = warp_sum_d(lane < p ? vv[lane] * x : 0.0);
                const double logdet = warp_sum_d(lane < p ? log(fabs(dq)) : 0.0);
                ll = -0.5 * ((double)p * (kLog2Pi + 2.0 * logdet) + quad);  // :696 (quirk Q2)
            }
            __syncwarp();
            // ---- a_f = a + KF s ;  jitter on the factor (Q4) ----
#pragma unroll
            for (int I = 0; I < MT; ++I) {
                double acc = 0.0;
#pragma unroll
                for (int J = 0; J < PT; ++J) {
                    const double2 kf = ldt(SA, (PT + I) * WT + J, lane);
                    acc = fma(kf.x, vsv[8 * J + 2 * t], acc);
                    acc = fma(kf.y, vsv[8 * J + 2 * t + 1], acc);
                }
                acc = red_t(acc);
                if (t == 0) vaf[8 * I + g] = va[8 * I + g] + acc;
            }
        } else {
            for (int i = lane; i < m; i += 32) vaf[i] = va[i];
        }
        if ((g >> 1) == t) {  // lanes holding a diagonal element: [g][2t+e] with 2t+e == g
#pragma unroll
            for (int I = 0; I < MT; ++I) SA[D::pc_tile(I, I) * 64 + lane * 2 + (g & 1)] += G.jitter;
        }
        __syncwarp();
        if (G.filt_a)
            for (int i = lane; i < m; i += 32) G.filt_a[bt * m + i] = vaf[i];
        if (G.filt_P)
            store_square<MT>(G.filt_P + bt * m * m, [&](int I, int K) { return ldt(SA, D::pc_tile(I, K), lane); }, lane);
        if (G.ll && lane == 0) G.ll[bt] = ll;
        logp += ll;
        // ---- predict array (transposed) [T Pcf, R Qc]; a+ = T a_f + c ----
        {
            double2 tp[MT][MT];
            double an[MT];
#pragma unroll
            for (int I = 0; I < MT; ++I) {
                an[I] = 0.0;
#pragma unroll
                for (int J = 0; J < MT; ++J) tp[I][J] = zero2();
            }
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                double2 pt[MT];
#pragma unroll
                for (int J = 0; J < MT; ++J) pt[J] = ttrans(ldt(SA, D::pc_tile(K, J), lane), lane);
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    const double2 tf = ldg_tile(Tg, m, I, K, lane);
                    an[I] = fma(tf.x, vaf[8 * K + 2 * t], an[I]);
                    an[I] = fma(tf.y, vaf[8 * K + 2 * t + 1], an[I]);
#pragma unroll
                    for (int J = 0; J < MT; ++J) mma_nt(tp[I][J], tf, pt[J]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int I = 0; I < MT; ++I) {
                const double s = red_t(an[I]) + vc[8 * I + g];
                if (t == 0) va[8 * I + g] = s;
#pragma unroll
                for (int J = 0; J < MT; ++J) stt(SA, D::pc_tile(I, J), lane, tp[I][J]);
#pragma unroll
                for (int J = 0; J < RT; ++J) stt(SA, (PT + I) * WT + J, lane, ldt(RQ, I * RT + J, lane));
            }
            __syncwarp();
        }
        qr_tiles<D::NT1, MT, PT, WT>(SA, smem + D::V_tq, MT, MT + RT, true, lane);
        if (rec) store_packed_upper<MT>(rec + off_rp, SA, lane, [](int I, int J) { return D::pc_tile(I, J); });
    }
    if (lane == 0) {
        if (!isfinite(logp)) st |= BKF_STATUS_NONFINITE;
        if (G.logp) G.logp[b] = logp;
        if (G.status) G.status[b] = st;
    }
}

// ===============================================================================================================
// Adjoint sweep.
//
// For A = Q R and an upper-triangular cotangent Rbar:  Abar = A Psi,  Psi = R^{-1} copyltu(R Rbar^T) R^{-T}  (symmetric).
// On tiles:   S0 = R Rbar^T = mma_nt(R, Rbar)  (lower tiles only), S = copyltu(S0),
//             G = S R^{-T}   (right-solve by tile columns, diagonal tiles inverted once),
//             Psi = G^T R^{-T}  (same right-solve after an in-place transposition).
// The R factors come from the forward record (no QR here).  Cotangents of the factor are carried TRANSPOSED
// (PBT = Pbar^T) because that is what both the pull-back (Rbar_p = triu(Pbar^T)) and the products below produce.

template <int NT>
__host__ __device__ constexpr int tri(int I, int K) { return I * NT - (I * (I - 1)) / 2 + (K - I); }  // I <= K

// tiles of the upper triangle of an n x n factor (n = 8 NT) from its packed rows (row i holds n - i entries)
template <int NT>
__device__ __forceinline__ void load_packed_upper(double* __restrict__ RF, const double* __restrict__ src, int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int n = 8 * NT;
#pragma unroll
    for (int I = 0; I < NT; ++I)
#pragma unroll
        for (int K = I; K < NT; ++K) {
            const int i = 8 * I + g, j = 8 * K + 2 * t;
            const double* q = src + (i * n - (i * (i - 1)) / 2 - i + j);
            double2 v;
            v.x = j >= i ? q[0] : 0.0;
            v.y = j + 1 >= i ? q[1] : 0.0;
            stt(RF, tri<NT>(I, K), lane, v);
        }
}

// Psi (NT x NT tiles, row stride NT, in W) from the factor tiles RF (upper, tri<NT> order) and a provider of the
// cotangent tiles rbar(J, K), K >= J, with exact zeros below the diagonal.  RINV: NT tiles of scratch.
template <int NT, class RbF>
__device__ __forceinline__ void psi_tiles(double* __restrict__ W, const double* __restrict__ RF, double* __restrict__ RINV,
                                          RbF rbar, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    // S = copyltu(R Rbar^T)
#pragma unroll
    for (int I = 0; I < NT; ++I) {
        Acc2 acc[NT];
#pragma unroll
        for (int J = 0; J < NT; ++J) acc[J] = acc2_zero();
#pragma unroll
        for (int K = I; K < NT; ++K) {
            const double2 rk = ldt(RF, tri<NT>(I, K), lane);
#pragma unroll
            for (int J = 0; J <= I; ++J) mma_nt(acc[J], rk, rbar(J, K));
        }
#pragma unroll
        for (int J = 0; J < I; ++J) {
            const double2 v = fin(acc[J]);
            stt(W, I * NT + J, lane, v);
            stt(W, J * NT + I, lane, ttrans(v, lane));
        }
        const double2 dv = fin(acc[I]);
        const double2 dt = ttrans(dv, lane);
        stt(W, I * NT + I, lane, make_double2(g >= 2 * t ? dv.x : dt.x, g >= 2 * t + 1 ? dv.y : dt.y));
    }
    __syncwarp();
    // inverses of the diagonal tiles: lane 8q + c solves column c of tile q (four tiles per round)
#pragma unroll
    for (int base = 0; base < NT; base += 4) {
        const int I = base + (lane >> 3), c = lane & 7;
        if (I < NT) {
            const double* Dg = RF + (I * NT - (I * (I - 1)) / 2) * 64;
            double x[8], rd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) rd[i] = 1.0 / Dg[i * 9];
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                double tv = i == c ? 1.0 : 0.0;
#pragma unroll
                for (int q = 7; q > i; --q) tv = fma(-Dg[i * 8 + q], x[q], tv);
                x[i] = tv * rd[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) RINV[I * 64 + i * 8 + c] = x[i];
        }
    }
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        // X <- X R^{-T}, tile column by tile column (all tile rows at once)
#pragma unroll
        for (int I = NT - 1; I >= 0; --I) {
            Acc2 acc[NT];
#pragma unroll
            for (int J = 0; J < NT; ++J) acc[J] = acc2_from(ldt(W, J * NT + I, lane));
#pragma unroll
            for (int K = I + 1; K < NT; ++K) {
                const double2 rk = neg(ldt(RF, tri<NT>(I, K), lane));
#pragma unroll
                for (int J = 0; J < NT; ++J) mma_nt(acc[J], ldt(W, J * NT + K, lane), rk);
            }
            const double2 ri = ldt(RINV, I, lane);
            Acc2 o[NT];
#pragma unroll
            for (int J = 0; J < NT; ++J) {
                o[J] = acc2_zero();
                mma_nt(o[J], fin(acc[J]), ri);
            }
#pragma unroll
            for (int J = 0; J < NT; ++J) stt(W, J * NT + I, lane, fin(o[J]));
            __syncwarp();
        }
        if (pass == 0) {
#pragma unroll
            for (int I = 0; I < NT; ++I) {
#pragma unroll
                for (int J = 0; J < I; ++J) {
                    const double2 lo = ldt(W, I * NT + J, lane), up = ldt(W, J * NT + I, lane);
                    stt(W, I * NT + J, lane, ttrans(up, lane));
                    stt(W, J * NT + I, lane, ttrans(lo, lane));
                }
                stt(W, I * NT + I, lane, ttrans(ldt(W, I * NT + I, lane), lane));
            }
            __syncwarp();
        }
    }
}

// C[M,N] (row-major, ldc) (op)= A B with strided operands, one warp (epilogue only)
template <bool ADD>
__device__ inline void wmm(double* C, int ldc, const double* A, int as0, int as1, const double* Bm, int bs0, int bs1,
                           int M, int N, int K, int lane) {
    for (int idx = lane; idx < M * N; idx += 32) {
        const int i = idx / N, j = idx - i * N;
        double acc = 0;
        for (int k = 0; k < K; ++k) acc = fma(A[i * as0 + k * as1], Bm[k * bs0 + j * bs1], acc);
        C[i * ldc + j] = ADD ? C[i * ldc + j] + acc : acc;
    }
    __syncwarp();
}

// Pull a lower-triangular factor cotangent Lbar back through S = L L^T (PyTensor lower convention, as
// sq::chol_pullback):  Sbar = tril(2 X) - diag(X),  X = 0.5 L^{-T} (P + P^T) L^{-1},  P = Phi(L^T Lbar).  tmp: 2 k k.
__device__ inline void chol_pullback_warp(double* out, const double* L, const double* Lbar, int k, double* tmp, int lane) {
    double* P = tmp;
    double* X = tmp + k * k;
    for (int idx = lane; idx < k * k; idx += 32) {
        const int i = idx / k, j = idx - i * k;
        double acc = 0;
        for (int q = (i > j ? i : j); q < k; ++q) acc = fma(L[q * k + i], Lbar[q * k + j], acc);
        P[idx] = (i > j) ? acc : (i == j ? 0.5 * acc : 0.0);
    }
    __syncwarp();
    for (int idx = lane; idx < k * k; idx += 32) X[idx] = P[idx] + P[(idx % k) * k + idx / k];
    __syncwarp();
    for (int c = lane; c < k; c += 32)  // X <- L^{-T} X, column c
        for (int i = k - 1; i >= 0; --i) {
            double tv = X[i * k + c];
            for (int q = i + 1; q < k; ++q) tv -= L[q * k + i] * X[q * k + c];
            X[i * k + c] = tv / L[i * k + i];
        }
    __syncwarp();
    for (int rw = lane; rw < k; rw += 32)  // X <- X L^{-1}, row rw
        for (int j = k - 1; j >= 0; --j) {
            double tv = X[rw * k + j];
            for (int q = j + 1; q < k; ++q) tv -= X[rw * k + q] * L[q * k + j];
            X[rw * k + j] = tv / L[j * k + j];
        }
    __syncwarp();
    for (int idx = lane; idx < k * k; idx += 32) {
        const int i = idx / k, j = idx - i * k;
        const double sv = 0.5 * X[idx];
        out[idx] = (i > j) ? 2.0 * sv : (i == j ? sv : 0.0);
    }
    __syncwarp();
}

template <int MT, int PT, int RT>
struct BDims {
    using D = Dims<MT, PT, RT>;
    static constexpr int m = D::m, p = D::p, NT1 = D::NT1;
    static constexpr int NTRI = NT1 * (NT1 + 1) / 2;
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    // tile counts: the arrays double as scratch for the products of the predict / update adjoints (see the kernel)
    static constexpr int WTILES = cmax(NT1 * NT1, 2 * MT * MT);                       // Psi | Psi_p + TP^T / U1
    static constexpr int RFTILES = cmax(NTRI, MT * PT);                               // factor | Abar21
    static constexpr int RBTILES = cmax(cmax(NTRI, MT * MT), NT1 + MT * PT);          // Rbar | RINV + A21 | U1^T
    static constexpr int W = 0;                           // Psi work array
    static constexpr int RF = W + WTILES * 64;            // factor tiles (upper)
    static constexpr int RB = RF + RFTILES * 64;          // cotangent tiles (upper); also RINV / scratch
    static constexpr int PBT = RB + RBTILES * 64;         // carried Pbar^T, MT x MT tiles
    static constexpr int HC = PBT + MT * MT * 64;         // Hc tiles
    static constexpr int HBT = HC + PT * PT * 64;         // sum of Abar11 (= Hcbar^T)
    static constexpr int VA = HBT + PT * PT * 64;
    static constexpr int V_a = VA, V_af = V_a + m, V_ab = V_af + m, V_afb = V_ab + m, V_gc = V_afb + m, V_v = V_gc + m,
                         V_s = V_v + p, V_in = V_s + p, V_w = V_in + p, V_ym = V_w + p, V_vb = V_ym + p, V_sb = V_vb + p,
                         V_t1 = V_sb + p, V_t2 = V_t1 + p, V_gd = V_t2 + p, V_d = V_gd + p;
    static constexpr int ELEMS = V_d + p;
    // epilogue scratch (row-major) laid over W, RF, RB
    static_assert((WTILES + RFTILES + RBTILES) * 64 >= cmax(5 * p * p, 6 * D::r * D::r + 2 * m * D::r), "epilogue scratch");
};

template <int MT, int PT, int RT>
__global__ void __launch_bounds__(32) backward_kernel(KArgs<double> G) {
    using D = Dims<MT, PT, RT>;
    using BD = BDims<MT, PT, RT>;
    constexpr int m = D::m, p = D::p, r = D::r, N1 = D::N1, NT1 = D::NT1;
    extern __shared__ double smem[];
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    const int n = G.n;
    double *W = smem + BD::W, *RF = smem + BD::RF, *RB = smem + BD::RB, *PBT = smem + BD::PBT, *HC = smem + BD::HC,
           *HBT = smem + BD::HBT;
    double *va = smem + BD::V_a, *vaf = smem + BD::V_af, *vab = smem + BD::V_ab, *vafb = smem + BD::V_afb,
           *vgc = smem + BD::V_gc, *vv = smem + BD::V_v, *vs = smem + BD::V_s, *vin = smem + BD::V_in, *vw = smem + BD::V_w,
           *vym = smem + BD::V_ym, *vvb = smem + BD::V_vb, *vsb = smem + BD::V_sb, *vt1 = smem + BD::V_t1,
           *vt2 = smem + BD::V_t2, *vgd = smem + BD::V_gd, *vd = smem + BD::V_d;

    bool Hzero;
    setup_consts<D>(G, b, HC, nullptr, nullptr, W, &Hzero, lane);
    const double* Tg = param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
    const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)p * m);
    {
        const double* dg = param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, (size_t)p);
        for (int i = lane; i < p; i += 32) { vd[i] = dg[i]; vgd[i] = 0.0; }
        for (int i = lane; i < m; i += 32) { vab[i] = 0.0; vgc[i] = 0.0; }
        for (int i = lane; i < MT * MT * 32; i += 32) reinterpret_cast<double2*>(PBT)[i] = zero2();
        for (int i = lane; i < PT * PT * 32; i += 32) reinterpret_cast<double2*>(HBT)[i] = zero2();
    }
    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n * p;
    const size_t tstride = traj_record_elems(m, p);
    const size_t off_rf = (size_t)m + (size_t)m * m, off_rp = off_rf + (size_t)N1 * (N1 + 1) / 2;
    // accumulators that are only added to live in global memory (L2), updated with fire-and-forget reductions (RED.ADD:
    // no load on the step's critical path): the gradient buffers themselves when requested, else the scratch record
    // behind the trajectory
    double* scr = G.traj + (size_t)G.B * n * tstride + (size_t)b * tstride;
    double* gT = G.g_T ? G.g_T + (size_t)b * m * m : scr;
    double* gZ = G.g_Z ? G.g_Z + (size_t)b * p * m : scr + m * m;
    double* PS = scr + m * m + p * m;
    const bool want_ps = G.g_R || G.g_Q;
    for (int i = lane; i < m * m; i += 32) { gT[i] = 0.0; PS[i] = 0.0; }
    for (int i = lane; i < p * m; i += 32) gZ[i] = 0.0;
    __syncwarp();

    for (int ts = n - 1; ts >= 0; --ts) {
        const size_t bt = (size_t)b * n + ts;
        const double* rec = G.traj + bt * tstride;
        const double* Pcg = rec + m;
        const double gcot = G.ll_cot ? G.ll_cot[bt] : 1.0;
        const bool dense_pc = ts == 0;  // P0 is whatever the caller passed; later factors are lower triangular
        if (ts > 0) {  // the next record this warp will read: pull it towards L2 while this step computes
            const char* nxt = reinterpret_cast<const char*>(rec - tstride);
            for (int off = lane * 128; off < (int)(tstride * sizeof(double)); off += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
        }
        // ---- mask, innovation (as the forward sweep) ----
        bool miss = true;
        if (lane < p) {
            const double yv = yb[(size_t)ts * p + lane];
            miss = is_missing(yv, G.fill, true);
            vw[lane] = miss ? 0.0 : 1.0;
            vym[lane] = miss ? 0.0 : yv;
        }
        const int nobs = __popc(__ballot_sync(FULL, !miss));
        const bool flag = nobs == 0, partial = !flag && nobs != p;
        for (int i = lane; i < m; i += 32) va[i] = rec[i];
        load_packed_upper<MT>(RF, rec + off_rp, lane);
        __syncwarp();
#pragma unroll
        for (int I = 0; I < PT; ++I) {
            const double wI = vw[8 * I + g];
            double acc = 0.0;
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                const double2 z = ldg_tile(Zg, m, I, J, lane);
                acc = fma(z.x * wI, va[8 * J + 2 * t], acc);
                acc = fma(z.y * wI, va[8 * J + 2 * t + 1], acc);
            }
            acc = red_t(acc) + vd[8 * I + g];
            if (t == 0) vv[8 * I + g] = vym[8 * I + g] - acc;
        }
        // ================= predict adjoint =================
        psi_tiles<MT>(W, RF, RB, [&](int J, int K) {
            double2 v = ldt(PBT, J * MT + K, lane);
            if (J == K) {
                if (2 * t < g) v.x = 0.0;
                if (2 * t + 1 < g) v.y = 0.0;
            }
            return v;
        }, lane);
        if (want_ps) {
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    red_tile(PS, m, I, J, lane, ldt(W, I * MT + J, lane));
                }
        }
        // ---- forward quantities of this step from the record: s, inner, a_f, Pcf ----
        if (!flag) {
            load_packed_upper<NT1>(RF, rec + off_rf, lane);
            __syncwarp();
            const int q = lane < p ? lane : p - 1;
            double frow[p];
#pragma unroll
            for (int i = 0; i < p; ++i)  // Fc[q][i] = Rf[i][q], i <= q
                frow[i] = (i >> 3) <= (q >> 3) ? RF[((i >> 3) * NT1 - ((i >> 3) * ((i >> 3) - 1)) / 2 + ((q >> 3) - (i >> 3))) * 64 +
                                                    (i & 7) * 8 + (q & 7)]
                                               : 0.0;
            const double rdiag = 1.0 / RF[((q >> 3) * NT1 - ((q >> 3) * ((q >> 3) - 1)) / 2) * 64 + (q & 7) * 9];
            double x = lane < p ? vv[lane] : 0.0;
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
                for (int i = 0; i < p; ++i) {
                    const double xi = __shfl_sync(FULL, x, i) * __shfl_sync(FULL, rdiag, i);
                    if (lane == i) x = xi;
                    else if (lane > i) x = fma(-frow[i], xi, x);
                }
                if (lane < p) (rep == 0 ? vs : vin)[lane] = x;
            }
            __syncwarp();
            // a_f = a + KF s,  KF[i][o] = Rf[o][p + i]
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int I = 0; I < PT; ++I) {
                    const double2 rk = ldt(RF, tri<NT1>(I, PT + J), lane);
                    const double sI = vs[8 * I + g];
                    c0 = fma(rk.x, sI, c0);
                    c1 = fma(rk.y, sI, c1);
                }
                c0 = red_g(c0);
                c1 = red_g(c1);
                if (g == 0) {
                    vaf[8 * J + 2 * t] = va[8 * J + 2 * t] + c0;
                    vaf[8 * J + 2 * t + 1] = va[8 * J + 2 * t + 1] + c1;
                }
            }
        } else {
            for (int i = lane; i < m; i += 32) vaf[i] = va[i];
        }
        __syncwarp();
        // (Pcf + jitter I)^T tile (I, K): upper triangular from the record's factor, or Pc^T when the step was skipped
        auto pft = [&](int I, int K) {
            double2 v;
            if (flag) v = ldg_tile_t(Pcg, m, I, K, lane);
            else v = K >= I ? ldt(RF, tri<NT1>(PT + I, PT + (K >= I ? K : I)), lane) : zero2();
            if (I == K) {
                if (2 * t == g) v.x += G.jitter;
                if (2 * t + 1 == g) v.y += G.jitter;
            }
            return v;
        };
        const bool dense_pf = flag && dense_pc;  // (a skipped step keeps Pc, which is dense only at t = 0)
        constexpr int TPT = MT * MT;  // tile offset of the second scratch block inside W
        // TP^T = (T Pcf)^T = Pcf^T T^T :  [I][J] = sum_K pft(I,K) (T[J][K])^T   (each T tile is fetched once)
        {
            double2 acc[MT][MT];
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < MT; ++J) acc[I][J] = zero2();
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                double2 tk[MT];
#pragma unroll
                for (int J = 0; J < MT; ++J) tk[J] = ldg_tile(Tg, m, J, K, lane);
#pragma unroll
                for (int I = 0; I < MT; ++I)
                    if (dense_pf || K >= I) {
                        const double2 pf = pft(I, K);
#pragma unroll
                        for (int J = 0; J < MT; ++J) mma_nt(acc[I][J], pf, tk[J]);
                    }
            }
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < MT; ++J) stt(W, TPT + I * MT + J, lane, acc[I][J]);
        }
        __syncwarp();
        // U1^T = TP^T Psi_p  -> RB[0 .. MT*MT)
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            Acc2 acc[MT];
#pragma unroll
            for (int J = 0; J < MT; ++J) acc[J] = acc2_zero();
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                const double2 tpk = ldt(W, TPT + I * MT + K, lane);
#pragma unroll
                for (int J = 0; J < MT; ++J) mma_nt(acc[J], tpk, ldt(W, J * MT + K, lane));
            }
#pragma unroll
            for (int J = 0; J < MT; ++J) stt(RB, I * MT + J, lane, fin(acc[J]));
        }
        __syncwarp();
        // U1 -> W[TPT ..)
#pragma unroll
        for (int I = 0; I < MT; ++I)
#pragma unroll
            for (int J = 0; J < MT; ++J) stt(W, TPT + I * MT + J, lane, ttrans(ldt(RB, J * MT + I, lane), lane));
        __syncwarp();
        // gT += U1 Pcf^T + abar af^T ;  Pcf[J][K] = (pft(K, J))^T
        if (G.g_T) {
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                Acc2 acc[MT];
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    const double ab = vab[8 * I + g];
                    acc[I] = acc2_from(make_double2(ab * vaf[8 * J + 2 * t], ab * vaf[8 * J + 2 * t + 1]));
                }
#pragma unroll
                for (int K = 0; K < MT; ++K)
                    if (dense_pf || K <= J) {
                        const double2 pk = ttrans(pft(K, J), lane);
#pragma unroll
                        for (int I = 0; I < MT; ++I) mma_nt(acc[I], ldt(W, TPT + I * MT + K, lane), pk);
                    }
#pragma unroll
                for (int I = 0; I < MT; ++I) red_tile(gT, m, I, J, lane, fin(acc[I]));
            }
        }
        // Pcfbar^T = U1^T T :  [I][J] = sum_K U1T[I][K] ((T^T)[J][K])^T  -> W[0 .. MT*MT)   (Psi_p is dead)
        // and, from the same tiles of T^T,  afbar = T^T abar
        {
            double afp[MT];
#pragma unroll
            for (int J = 0; J < MT; ++J) afp[J] = 0.0;
            double2 acc[MT][MT];
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < MT; ++J) acc[I][J] = zero2();
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                double2 tk[MT];
                const double ab0 = vab[8 * K + 2 * t], ab1 = vab[8 * K + 2 * t + 1];
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    tk[J] = ldg_tile_t(Tg, m, J, K, lane);  // [g][2t+e] = T[8K + 2t + e][8J + g]
                    afp[J] = fma(tk[J].x, ab0, fma(tk[J].y, ab1, afp[J]));
                }
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    const double2 uk = ldt(RB, I * MT + K, lane);
#pragma unroll
                    for (int J = 0; J < MT; ++J) mma_nt(acc[I][J], uk, tk[J]);
                }
            }
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < MT; ++J) stt(W, I * MT + J, lane, acc[I][J]);
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                const double v = red_t(afp[J]);
                if (t == 0) vafb[8 * J + g] = v;
            }
        }
        // gc += abar
        for (int i = lane; i < m; i += 32) vgc[i] += vab[i];
        __syncwarp();
        if (flag) {  // degenerate branch (:700-705): (a_f, Pcf, ll) = (a, Pc, 0)
            for (int i = lane; i < m; i += 32) vab[i] = vafb[i];
#pragma unroll
            for (int i = 0; i < MT * MT; ++i) stt(PBT, i, lane, ldt(W, i, lane));
            __syncwarp();
            continue;
        }
        // ================= update adjoint =================
        {
            const double lossbar = -0.5 * gcot, logdetbar = -0.5 * gcot * (double)p;
            // sbar = KF^T afbar
#pragma unroll
            for (int I = 0; I < PT; ++I) {
                double acc = 0.0;
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    const double2 rk = ldt(RF, tri<NT1>(I, PT + J), lane);
                    acc = fma(rk.x, vafb[8 * J + 2 * t], acc);
                    acc = fma(rk.y, vafb[8 * J + 2 * t + 1], acc);
                }
                acc = red_t(acc);
                if (t == 0) vsb[8 * I + g] = acc;
            }
            __syncwarp();
            // back substitution with U = Rf[:p,:p] = Fc^T : lane q owns x[q] and row q of U
            const int q = lane < p ? lane : p - 1;
            double urow[p];
#pragma unroll
            for (int i = 0; i < p; ++i)
                urow[i] = (i >> 3) >= (q >> 3) ? RF[((q >> 3) * NT1 - ((q >> 3) * ((q >> 3) - 1)) / 2 + ((i >> 3) - (q >> 3))) * 64 +
                                                    (q & 7) * 8 + (i & 7)]
                                               : 0.0;
            const double dq = RF[((q >> 3) * NT1 - ((q >> 3) * ((q >> 3) - 1)) / 2) * 64 + (q & 7) * 9];
            const double rdiag = 1.0 / dq;
            auto usolve = [&](double x) {
#pragma unroll
                for (int i = p - 1; i >= 0; --i) {
                    const double xi = __shfl_sync(FULL, x, i) * __shfl_sync(FULL, rdiag, i);
                    if (lane == i) x = xi;
                    else if (lane < i) x = fma(-urow[i], xi, x);
                }
                return x;
            };
            double vbar = lane < p ? lossbar * vin[lane] : 0.0;
            const double t1 = usolve(lane < p ? lossbar * vv[lane] : 0.0);          // Fc^{-T} innerbar
            const double t2 = usolve(lane < p ? vsb[lane] + t1 : 0.0);              // Fc^{-T} (sbar + t1)
            vbar += t2;
            if (lane < p) {
                vt1[lane] = t1;
                vt2[lane] = t2;
                vvb[lane] = vbar;
                vsb[lane] = logdetbar * 2.0 * rdiag;  // diagonal term of Fcbar
            }
            __syncwarp();
            // Rbar = Bbar^T (upper):  [Fcbar^T, KFbar^T; 0, triu(Pcfbar^T)]
#pragma unroll
            for (int I = 0; I < PT; ++I) {
                const int i = 8 * I + g;
                const double ini = vin[i], si = vs[i];
#pragma unroll
                for (int K = I; K < PT; ++K) {
                    const int k = 8 * K + 2 * t;
                    double2 v;
                    v.x = k >= i ? fma(-vt1[k], ini, -vt2[k] * si) + (k == i ? vsb[i] : 0.0) : 0.0;
                    v.y = k + 1 >= i ? fma(-vt1[k + 1], ini, -vt2[k + 1] * si) + (k + 1 == i ? vsb[i] : 0.0) : 0.0;
                    stt(RB, tri<NT1>(I, K), lane, v);
                }
#pragma unroll
                for (int K = 0; K < MT; ++K)
                    stt(RB, tri<NT1>(I, PT + K), lane,
                        make_double2(vafb[8 * K + 2 * t] * si, vafb[8 * K + 2 * t + 1] * si));
            }
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int K = I; K < MT; ++K) {
                    double2 v = ldt(W, I * MT + K, lane);
                    if (I == K) {
                        if (2 * t < g) v.x = 0.0;
                        if (2 * t + 1 < g) v.y = 0.0;
                    }
                    stt(RB, tri<NT1>(PT + I, PT + K), lane, v);
                }
            __syncwarp();
        }
        // RINV must not overlap the cotangent tiles while S0 is formed: psi_tiles reads rbar only in its first phase
        psi_tiles<NT1>(W, RF, RB, [&](int J, int K) { return ldt(RB, tri<NT1>(J, K), lane); }, lane);
        // ---- Abar = A Psi, A = [[Hc^T, 0], [A21, Pc^T]],  A21 = (Zm Pc)^T = Pc^T Zm^T -> RB[NT1 ..) ----
        constexpr int A21T = NT1;  // behind RINV
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            Acc2 acc[PT];
#pragma unroll
            for (int J = 0; J < PT; ++J) acc[J] = acc2_zero();
#pragma unroll
            for (int K = 0; K < MT; ++K)
                if (dense_pc || K >= I) {
                    const double2 pk = ldg_tile_t(Pcg, m, I, K, lane);
#pragma unroll
                    for (int J = 0; J < PT; ++J) {
                        double2 z = ldg_tile(Zg, m, J, K, lane);
                        const double wJ = vw[8 * J + g];
                        z.x *= wJ;
                        z.y *= wJ;
                        mma_nt(acc[J], pk, z);
                    }
                }
#pragma unroll
            for (int J = 0; J < PT; ++J) stt(RB, A21T + I * PT + J, lane, fin(acc[J]));
        }
        __syncwarp();
        // Abar11 = Hc^T Psi11  (accumulated: Hcbar^T)
#pragma unroll
        for (int I = 0; I < PT; ++I) {
            Acc2 acc[PT];
#pragma unroll
            for (int J = 0; J < PT; ++J) acc[J] = acc2_from(ldt(HBT, I * PT + J, lane));
#pragma unroll
            for (int K = I; K < PT; ++K) {
                double2 h = ttrans(ldt(HC, K * PT + I, lane), lane);  // (Hc^T)[I][K]
                if (Hzero) h = zero2();
                else if (partial) h = make_double2(NAN, NAN);
#pragma unroll
                for (int J = 0; J < PT; ++J) mma_nt(acc[J], h, ldt(W, J * NT1 + K, lane));
            }
#pragma unroll
            for (int J = 0; J < PT; ++J) stt(HBT, I * PT + J, lane, fin(acc[J]));
        }
        // rows of [Abar21, Abar22]; Abar21 -> RF[0 ..) (the factor is dead), Pbar^T = Abar22 + Abar21 Zm -> PBT
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            Acc2 acc2[NT1];
#pragma unroll
            for (int J = 0; J < NT1; ++J) acc2[J] = acc2_zero();
#pragma unroll
            for (int K = 0; K < NT1; ++K)
                if (K < PT || dense_pc || K - PT >= I) {
                    const double2 ak = K < PT ? ldt(RB, A21T + I * PT + (K < PT ? K : 0), lane)
                                              : ldg_tile_t(Pcg, m, I, K >= PT ? K - PT : 0, lane);
#pragma unroll
                    for (int J = 0; J < NT1; ++J) mma_nt(acc2[J], ak, ldt(W, J * NT1 + K, lane));
                }
            double2 acc[NT1];
#pragma unroll
            for (int J = 0; J < NT1; ++J) acc[J] = fin(acc2[J]);
#pragma unroll
            for (int J = 0; J < PT; ++J) stt(RF, I * PT + J, lane, acc[J]);
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                Acc2 pb = acc2_from(acc[PT + J]);
#pragma unroll
                for (int K = 0; K < PT; ++K) {
                    double2 zt = ldg_tile_t(Zg, m, J, K, lane);  // (Zm^T)[J][K]: [g][2t+e] = Zm[8K+2t+e][8J+g]
                    zt.x *= vw[8 * K + 2 * t];
                    zt.y *= vw[8 * K + 2 * t + 1];
                    mma_nt(pb, acc[K], zt);
                }
                stt(PBT, I * MT + J, lane, fin(pb));
            }
        }
        __syncwarp();
        // gZ += W (ZPbar Pc^T - vbar a^T),  ZPbar = Abar21^T
        if (G.g_Z) {
            double2 zp[PT][MT];
#pragma unroll
            for (int Io = 0; Io < PT; ++Io)
#pragma unroll
                for (int K = 0; K < MT; ++K) zp[Io][K] = ttrans(ldt(RF, K * PT + Io, lane), lane);
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                double2 acc[PT];
#pragma unroll
                for (int Io = 0; Io < PT; ++Io) acc[Io] = zero2();
#pragma unroll
                for (int K = 0; K < MT; ++K)
                    if (dense_pc || K <= J) {
                        const double2 pc = ldg_tile(Pcg, m, J, K, lane);
#pragma unroll
                        for (int Io = 0; Io < PT; ++Io) mma_nt(acc[Io], zp[Io][K], pc);
                    }
#pragma unroll
                for (int Io = 0; Io < PT; ++Io) {
                    const double wo = vw[8 * Io + g], vb = vvb[8 * Io + g];
                    red_tile(gZ, m, Io, J, lane, make_double2(wo * fma(-vb, va[8 * J + 2 * t], acc[Io].x),
                                                              wo * fma(-vb, va[8 * J + 2 * t + 1], acc[Io].y)));
                }
            }
        }
        // abar = afbar - Zm^T vbar ; gd -= vbar
#pragma unroll
        for (int J = 0; J < MT; ++J) {
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int I = 0; I < PT; ++I) {
                const double2 z = ldg_tile(Zg, m, I, J, lane);
                const double f = vw[8 * I + g] * vvb[8 * I + g];
                c0 = fma(z.x, f, c0);
                c1 = fma(z.y, f, c1);
            }
            c0 = red_g(c0);
            c1 = red_g(c1);
            if (g == 0) {
                vab[8 * J + 2 * t] = vafb[8 * J + 2 * t] - c0;
                vab[8 * J + 2 * t + 1] = vafb[8 * J + 2 * t + 1] - c1;
            }
        }
        if (lane < p) vgd[lane] -= vvb[lane];
        __syncwarp();
    }
    // ---- gradients out ----
    for (int i = lane; i < m; i += 32) {
        if (G.g_a0) G.g_a0[(size_t)b * m + i] = vab[i];
        if (G.g_c) G.g_c[(size_t)b * m + i] = vgc[i];
    }
    if (G.g_d && lane < p) G.g_d[(size_t)b * p + lane] = vgd[lane];
    if (G.g_P0) {
        double* dst = G.g_P0 + (size_t)b * m * m;
#pragma unroll
        for (int I = 0; I < MT; ++I)
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                const double2 v = ldt(PBT, I * MT + J, lane);  // Pbar[8J+2t+e][8I+g]
                dst[(8 * J + 2 * t) * m + 8 * I + g] = v.x;
                dst[(8 * J + 2 * t + 1) * m + 8 * I + g] = v.y;
            }
    }
    __syncwarp();
    double* E = smem + BD::W;  // W, RF, RB are dead: one row-major scratch block
    if (G.g_H) {
        double* dst = G.g_H + (size_t)b * p * p;
        if (Hzero) {
            for (int idx = lane; idx < p * p; idx += 32) dst[idx] = HBT[elem_addr(idx % p, idx / p, PT)];
        } else {
            double *L = E, *Lb = E + p * p, *out = E + 2 * p * p, *tmp = E + 3 * p * p;
            for (int idx = lane; idx < p * p; idx += 32) {
                const int i = idx / p, j = idx - i * p;
                L[idx] = HC[elem_addr(i, j, PT)];
                Lb[idx] = j <= i ? HBT[elem_addr(j, i, PT)] : 0.0;
            }
            __syncwarp();
            chol_pullback_warp(out, L, Lb, p, tmp, lane);
            for (int idx = lane; idx < p * p; idx += 32) dst[idx] = out[idx];
            __syncwarp();
        }
    }
    if (want_ps) {
        const double* Rg = param_ptr(G.Rm, G.shared_mask, BKF_PARAM_R, b, (size_t)m * r);
        const double* Qg = param_ptr(G.Q, G.shared_mask, BKF_PARAM_Q, b, (size_t)r * r);
        double *Qc = E, *X1 = Qc + r * r, *QQ = X1 + m * r, *Y = QQ + r * r, *Qb = Y + m * r, *out = Qb + r * r,
               *tmp = out + r * r;
        for (int i = lane; i < r * r; i += 32) Qc[i] = Qg[i];
        __syncwarp();
        chol_lower_warp(Qc, r, lane);
        __threadfence();  // the reductions into PS were issued by other lanes of this warp
        __syncwarp();
        wmm<false>(X1, r, PS, m, 1, Rg, r, 1, m, r, m, lane);  // X1 = Psum R
        if (G.g_R) {
            wmm<false>(QQ, r, Qc, r, 1, Qc, 1, r, r, r, r, lane);  // Qc Qc^T
            wmm<false>(Y, r, X1, r, 1, QQ, r, 1, m, r, r, lane);
            for (int i = lane; i < m * r; i += 32) G.g_R[(size_t)b * m * r + i] = Y[i];
            __syncwarp();
        }
        if (G.g_Q) {
            wmm<false>(Y, r, X1, r, 1, Qc, r, 1, m, r, r, lane);   // X1 Qc
            wmm<false>(Qb, r, Rg, 1, r, Y, r, 1, r, r, m, lane);   // R^T X1 Qc
            for (int idx = lane; idx < r * r; idx += 32)
                if ((idx % r) > (idx / r)) Qb[idx] = 0.0;
            __syncwarp();
            chol_pullback_warp(out, Qc, Qb, r, tmp, lane);
            for (int i = lane; i < r * r; i += 32) G.g_Q[(size_t)b * r * r + i] = out[i];
        }
    }
}

}  // namespace sqw
}  // namespace bkf
