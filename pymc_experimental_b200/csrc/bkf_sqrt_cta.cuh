// bkf_sqrt_cta.cuh -- SquareRootFilter (kalman_filter.py:620-750), FOUR WARPS PER SERIES with a look-ahead Householder
// sweep.  Same arithmetic as bkf_sqrt_warp.cuh (dgeqr2 / dlarfg sign rules, Q-free QR pull-back), reorganised for the
// regime the VARMAX config lives in (a few hundred series per GPU, i.e. fewer series than warp slots):
//
//   * The serial part of a step is the chain of 80 Householder reflectors (48 of the update array, 32 of the predict
//     array).  ONE warp ("chain warp", warp 0) owns that chain and nothing else: the 8-column panel it factors and the
//     next column block live in its registers (column g of a block in lane group g, as in bkf_sqrt_warp.cuh), a
//     reflector costs one broadcast, one 4-lane reduction and a division-free scalar chain
//     (kappa = 1 / (nrm (|alpha| + nrm)) from overlapped rsqrt / rcp seeds), and the panel is never blocked: there is
//     no compact-WY factor, no V to store.
//   * Every reflector is PUBLISHED (u and kappa, 8 slots of shared memory, one named barrier per slot: the chain warp
//     only arrives, it never waits) and the helper warps apply it to the column blocks they own, reflector by reflector,
//     one reflector behind the chain.  A column block changes hands once: its owner stores it when the panel before it
//     is finished and the chain warp picks it up as its look-ahead block.  The chain warp therefore never executes a
//     trailing update, and the helpers never see the chain.
//   * Everything that is not the chain runs beside it: the products Z Pc and T Pcf (DMMA tiles, split over the four
//     warps), and -- during the predict sweep, on the fourth warp -- the two triangular solves, the log-likelihood,
//     a_f, a+, the innovation of the next step and the filter outputs.
//   * R factors go to the record for the adjoint as they are finished (tile by tile, 512-byte coalesced stores, already
//     in the orientation the adjoint multiplies with).  The record is [a | R_update tiles | R_predict tiles]: Pc is not
//     stored (it is the previous step's R_predict, or P0).
//
// Shapes: as bkf_sqrt_warp.cuh (float64; m, p, r multiples of 8; r <= p).
#pragma once
#include "bkf_sqrt_warp.cuh"

namespace bkf {
namespace sqc {

using namespace sqw;

constexpr int NWARP = 4;
constexpr int NTHR = 32 * NWARP;
constexpr int BAR_REFL = 1;  // named barriers 1..8: reflector k of the current panel is published
constexpr int BAR_HAND = 9;  // the helpers have finished the panel (their column blocks are back in shared memory)

__device__ __forceinline__ void bar_sync(int id, int nthr) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthr) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthr) : "memory");
}
// mbarriers (one per reflector slot, one arrival -- lane 0 of the chain warp -- per phase): the arrive is one more
// operation in the chain warp's in-order shared-memory queue behind the stores it publishes (no drain, no stall: a
// bar.arrive costs the chain warp ~130 cycles per reflector), the helpers sleep in try_wait.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned addr, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(addr), "r"(parity) : "memory");
}
// sum over the four lanes of a group with three independent shuffles (one shuffle latency instead of two); every lane
// of the group gets the same bits
__device__ __forceinline__ double red4(double x) {
    const double a = __shfl_xor_sync(FULL, x, 1), b = __shfl_xor_sync(FULL, x, 2), c = __shfl_xor_sync(FULL, x, 3);
    return (x + a) + (b + c);
}
// The same sum on the FP64 tensor pipe: in the m8n8k4 fragments lane (g, t) holds A[g][t], so A x ones gives every lane of
// group g the sum over the group's four lanes -- one DMMA instead of six shuffles and three adds.
__device__ __forceinline__ double red4_mma(double x) {
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, x, 1.0);
    return c0;
}
__device__ __forceinline__ void stg_rec(double* __restrict__ rec, int tile, int lane, double2 v) {
    reinterpret_cast<double2*>(rec)[tile * 32 + lane] = v;
}
__device__ __forceinline__ double2 ldg_rec(const double* __restrict__ rec, int tile, int lane) {
    return reinterpret_cast<const double2*>(rec)[tile * 32 + lane];
}
// helpers at work on nb column blocks behind a panel: helper 0 takes the first, the others two each
__host__ __device__ constexpr int nhelpers(int nb, int nhelp) {
    const int want = nb <= 0 ? 0 : 1 + (nb >> 1);
    return want < nhelp ? want : nhelp;
}
__host__ __device__ constexpr int trin(int NT, int I, int K) { return I * NT - (I * (I - 1)) / 2 + (K - I); }  // I <= K

template <int MT, int PT, int RT>
struct Dims {
    using W = sqw::Dims<MT, PT, RT>;
    static constexpr int mt = MT, pt = PT, rt = RT;
    static constexpr int m = 8 * MT, p = 8 * PT, r = 8 * RT;
    static constexpr int NT1 = MT + PT;  // column / row blocks of the update array [[Hc^T, 0], [Pc^T Zm^T, Pc^T]]
    static constexpr int NRP = MT + RT;  // row blocks of the predict array [T Pcf, R Qc]^T (MT column blocks)
    static constexpr int NRMAX = NT1 > NRP ? NT1 : NRP;
    static_assert(RT <= PT, "r <= p");
    static_assert(NT1 - 2 <= 2 * (NWARP - 2) && MT - 2 <= 2 * (NWARP - 3), "a helper owns at most two column blocks");
    static constexpr int NTRI1 = NT1 * (NT1 + 1) / 2, NTRIP = MT * (MT + 1) / 2;
    // record (doubles): a | R_update (upper tiles, trin(NT1) order) | R_predict (upper tiles, trin(MT) order)
    static constexpr int OFF_RF = m, OFF_RP = m + NTRI1 * 64, RSTRIDE = m + (NTRI1 + NTRIP) * 64;
    // forward arena (doubles)
    static constexpr int UA = 0;                      // update array, tile (column block j, row block q) at j NT1 + q
    static constexpr int PA = UA + NT1 * NT1 * 64;    // predict array, tile (j, q) at j NRP + q
    static constexpr int RQ = PA + MT * NRP * 64;     // R Qc, MT x RT tiles
    static constexpr int HC = RQ + MT * RT * 64;      // Hc, PT x PT tiles
    static constexpr int TT = HC + PT * PT * 64;      // T, MT x MT tiles
    static constexpr int ZZ = TT + MT * MT * 64;      // Z, PT x MT tiles
    static constexpr int UBS = 8 * NRMAX;             // one published reflector
    static constexpr int UB = ZZ + PT * MT * 64;      // 8 slots
    static constexpr int KB = UB + 8 * UBS;           // 8 x (kappa, u_k)
    static constexpr int VA = KB + 32;                // (+16: debug counters of BKF_SQC_PROFILE builds)
    static constexpr int V_a = VA, V_af = V_a + m, V_c = V_af + m, V_d = V_c + m, V_v = V_d + p, V_sv = V_v + p,
                         V_in = V_sv + p, V_w = V_in + p, V_ym = V_w + p, V_yh = V_ym + p, SI = V_yh + p;
    static constexpr int MB = SI + 4;          // 8 ints: flag[2], partial[2], Hzero, status; then 8 mbarriers
    static constexpr int FWD_ELEMS = MB + 8;
    static_assert(NT1 * NT1 * 64 >= p * p + r * r + m * r, "Cholesky scratch of the set-up fits the update array");
};

// ---------------------------------------------------------------------------------------------------------------
// Chain warp: one panel (8 reflectors) of the sweep.  pa = the panel (column block jt, row blocks jt .. jt+NQ-1, the
// diagonal block first).  A lone warp issues one instruction every ~4 cycles on this code, so the chain warp executes
// as few instructions as possible: the pivot column goes to the reflector slot in shared memory and comes back to every
// lane group with broadcast loads (that store is also the publication for the helpers), the norm is the pivot group's
// own inner product, and no other column block is touched here.
template <int NQ>
__device__ __forceinline__ void w0_panel(double* __restrict__ A, const int WT, const int jt, const int np,
                                         double* __restrict__ UBp, double* __restrict__ KBp, const int ubs,
                                         double* __restrict__ recR, const int nhelp, const unsigned mb,
                                         const int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr bool carry = false;  // (a look-ahead block in the chain warp costs it ~105 cycles per reflector)
    // helpers at work in this panel / in the previous one: helper 0 owns the next column block, the others two each
#ifdef BKF_SQC_NOPUBLISH
    const int nact = 0, nprev = 0;
#else
    const int nact = nhelpers(np - jt - 1, nhelp), nprev = nhelpers(np - jt, nhelp);
#endif
#ifdef BKF_SQC_PROFILE
    const long long h1 = clock64();
#endif
    double2 pa[NQ], cb[NQ];  // the panel and the look-ahead block jt + 1
#ifdef BKF_SQC_NOHELP
    if (false) {
#else
    if (jt > 0 && nprev > 0) {  // this column block is back from its helper, the reflector slots are free
#endif
#ifdef BKF_SQC_PROFILE
        const long long h0 = clock64();
#endif
        bar_sync(BAR_HAND, 32 * (1 + nprev));
#ifdef BKF_SQC_PROFILE
        if (lane == 0) { KBp[16] += (double)(clock64() - h0); KBp[17] += 1.0; }
#endif
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        pa[q] = ldt(A, jt * WT + jt + q, lane);
        cb[q] = zero2();
    }
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
        // column k of the panel to every lane group (lane (k, t) -> lanes (*, t))
        double2 xb[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            xb[q].x = __shfl_sync(FULL, pa[q].x, 4 * k + t);
            xb[q].y = __shfl_sync(FULL, pa[q].y, 4 * k + t);
        }
        const int srck = (lane & ~3) | (k >> 1);
        const double akc = __shfl_sync(FULL, (k & 1) ? pa[0].y : pa[0].x, srck);    // A[row k][col g]
        const double alpha = __shfl_sync(FULL, (k & 1) ? xb[0].y : xb[0].x, srck);  // A[row k][col k]
        double2 xm = xb[0];  // strictly below the diagonal
        if (2 * t <= k) xm.x = 0.0;
        if (2 * t + 1 <= k) xm.y = 0.0;
        // this lane's share of x . a_g and of x . x (every group forms the same x . x: no broadcast of the norm)
        double dA = xm.x * pa[0].x, dB = xm.y * pa[0].y, sA = xm.x * xm.x, sB = xm.y * xm.y;
        double dC = 0.0, dD = 0.0, sC = 0.0, sD = 0.0;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            if (q & 1) {
                dC = fma(xb[q].x, pa[q].x, dC);
                dD = fma(xb[q].y, pa[q].y, dD);
                sC = fma(xb[q].x, xb[q].x, sC);
                sD = fma(xb[q].y, xb[q].y, sD);
            } else {
                dA = fma(xb[q].x, pa[q].x, dA);
                dB = fma(xb[q].y, pa[q].y, dB);
                sA = fma(xb[q].x, xb[q].x, sA);
                sB = fma(xb[q].y, xb[q].y, sB);
            }
        }
        const double ss = red4_mma(NQ > 1 ? (sA + sB) + (sC + sD) : sA + sB);   // x . x
        const double dot = red4_mma(NQ > 1 ? (dA + dB) + (dC + dD) : dA + dB);  // x . a_g
        // dlarfg: beta = -sign(alpha) nrm, u = [alpha - beta; x], H = I - kappa u u^T with
        // kappa = tau / (alpha - beta)^2 = 1 / (nrm (|alpha| + nrm)).  ss == 0: H = I.
        // nrm = sqrt(x2) from the 20-bit rsqrt seed and ONE cubic (Halley) step: nrm = t + t e (1/2 + 3/8 e), t = x2 y0,
        // e = 1 - t y0; kappa from the 20-bit reciprocal seed of the first-order denominator and one cubic step
        // s (1 + e (1 + e)).  Four dependent operations each instead of six.
        double kap = 0.0, upv = 0.0, bet = alpha;
        if (ss != 0.0) {
            const double x2 = fma(alpha, alpha, ss);  // (x2 subnormal, a column of magnitude < 1e-154, is out of range)
            const double aa = fabs(alpha);
            double y0;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x2));
            const double tt = x2 * y0;
            const double e = fma(-tt, y0, 1.0);
            double s0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s0) : "d"(fma(aa, tt, x2)));
            const double nrm = fma(tt * e, fma(0.375, e, 0.5), tt);
            const double nd = fma(aa, nrm, x2);
            const double e2 = fma(-nd, s0, 1.0);
            kap = fma(s0 * e2, 1.0 + e2, s0);
            bet = copysign(nrm, -alpha);
            upv = copysign(aa + nrm, alpha);
        }
        if (nact > 0) {  // publish x (rows above the diagonal included: readers mask them) and (kappa, u_k)
            if (g == 0) {
                double2* ub = reinterpret_cast<double2*>(UBp + k * ubs);
#pragma unroll
                for (int q = 0; q < NQ; ++q) ub[q * 4 + t] = xb[q];
            }
            if (lane == 0) {
                reinterpret_cast<double2*>(KBp)[k] = make_double2(kap, upv);
                mbar_arrive(mb + 8 * k);
            }
        }
        const double w = g > k ? kap * fma(upv, akc, dot) : 0.0;  // kappa (u . a_g)
        double2 ud = xm;  // u on the rows of the diagonal block
        if (2 * t == k) ud.x = upv;
        if (2 * t + 1 == k) ud.y = upv;
        {
            const bool piv = g == k;
            const double nx = fma(-w, ud.x, pa[0].x), ny = fma(-w, ud.y, pa[0].y);
            pa[0].x = (piv && 2 * t == k) ? bet : nx;
            pa[0].y = (piv && 2 * t + 1 == k) ? bet : ny;
        }
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            pa[q].x = fma(-w, xb[q].x, pa[q].x);
            pa[q].y = fma(-w, xb[q].y, pa[q].y);
        }
        if (carry) {  // look-ahead block: off the chain
            double eA = ud.x * cb[0].x, eB = ud.y * cb[0].y;
#pragma unroll
            for (int q = 1; q < NQ; ++q) {
                eA = fma(xb[q].x, cb[q].x, eA);
                eB = fma(xb[q].y, cb[q].y, eB);
            }
            const double w2 = kap * red4_mma(eA + eB);
            cb[0].x = fma(-w2, ud.x, cb[0].x);
            cb[0].y = fma(-w2, ud.y, cb[0].y);
#pragma unroll
            for (int q = 1; q < NQ; ++q) {
                cb[q].x = fma(-w2, xb[q].x, cb[q].x);
                cb[q].y = fma(-w2, xb[q].y, cb[q].y);
            }
        }
    }
    // R^T of the panel goes back to the array (zeros below the diagonal of R) and, transposed, to the record
    {
        const double2 rd = make_double2(2 * t <= g ? pa[0].x : 0.0, 2 * t + 1 <= g ? pa[0].y : 0.0);
        stt(A, jt * WT + jt, lane, rd);
#pragma unroll
        for (int q = 1; q < NQ; ++q) stt(A, jt * WT + jt + q, lane, zero2());
        if (recR) stg_rec(recR, trin(np, jt, jt), lane, ttrans(rd, lane));
    }
    if (carry) {
        stt(A, (jt + 1) * WT + jt, lane, cb[0]);  // rows of R finished by this panel
        if (recR) stg_rec(recR, trin(np, jt, jt + 1), lane, ttrans(cb[0], lane));
#pragma unroll
        for (int q = 1; q < NQ; ++q) stt(A, (jt + 1) * WT + jt + q, lane, cb[q]);
    }
    __syncwarp();
#ifdef BKF_SQC_PROFILE
    if (lane == 0) KBp[18 + NQ] += (double)(clock64() - h1);
#endif
}

// Helper warp: applies the published reflectors of panel jt to the NB column blocks it owns in this panel (c0 and, for
// NB = 2, c1).  nbar: threads taking part in this panel's barriers (the chain warp and the helpers that own a block).
template <int NQ, int NB>
__device__ __forceinline__ void helper_panel(double* __restrict__ A, const int WT, const int jt, const int np,
                                             const double* __restrict__ UBp, const double* __restrict__ KBp, const int ubs,
                                             double* __restrict__ recR, const int nbar, const int c0, const int c1,
                                             const unsigned mb, const unsigned parity, const int lane) {
    const int t = lane & 3;
    double2 bb[NB][NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        bb[0][q] = ldt(A, c0 * WT + jt + q, lane);
        if (NB > 1) bb[NB - 1][q] = ldt(A, c1 * WT + jt + q, lane);
    }
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
        mbar_wait(mb + 8 * k, parity);
        const double2 ku = reinterpret_cast<const double2*>(KBp)[k];  // kappa, u_k
        double2 u[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) u[q] = reinterpret_cast<const double2*>(UBp + k * ubs)[q * 4 + t];
        u[0].x = 2 * t < k ? 0.0 : (2 * t == k ? ku.y : u[0].x);
        u[0].y = 2 * t + 1 < k ? 0.0 : (2 * t + 1 == k ? ku.y : u[0].y);
        double e[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            double eA = u[0].x * bb[i][0].x, eB = u[0].y * bb[i][0].y;
#pragma unroll
            for (int q = 1; q < NQ; ++q) {
                eA = fma(u[q].x, bb[i][q].x, eA);
                eB = fma(u[q].y, bb[i][q].y, eB);
            }
            e[i] = eA + eB;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) e[i] = ku.x * red4_mma(e[i]);
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                bb[i][q].x = fma(-e[i], u[q].x, bb[i][q].x);
                bb[i][q].y = fma(-e[i], u[q].y, bb[i][q].y);
            }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int c = i == 0 ? c0 : c1;
        stt(A, c * WT + jt, lane, bb[i][0]);  // rows of R finished by this panel
#pragma unroll
        for (int q = 1; q < NQ; ++q) stt(A, c * WT + jt + q, lane, bb[i][q]);
    }
    __syncwarp();
    bar_sync(BAR_HAND, nbar);  // the next panel (chain warp) and the next owners of these blocks may read them
    if (recR) {
#pragma unroll
        for (int i = 0; i < NB; ++i) stg_rec(recR, trin(np, jt, i == 0 ? c0 : c1), lane, ttrans(bb[i][0], lane));
    }
}

// Householder R factor of an array of np column blocks x nr row blocks (nr >= np) stored transposed in tiles (tile
// (j, q) at j WT + q holds element [col][row] in C layout).  Role 0 is the chain warp, roles 1 .. nhelp the helpers;
// other warps do not take part.  In panel jt helper 0 owns the column block the chain warp needs next (jt + 1) and
// nothing else; the blocks behind it go to the other helpers two by two.  On return (after a CTA barrier) the tiles hold
// R^T, zeros elsewhere; recR (nullable) gets the upper tiles of R in trin(np) order.  pub: panels published so far in
// the kernel (every one flips the phase of the eight slot barriers); np - 1 more when this returns.
template <int NRMAX>
__device__ __noinline__ void qr_cta(double* __restrict__ A, const int WT, const int np, const int nr,
                                    double* __restrict__ UBp, double* __restrict__ KBp, double* __restrict__ recR,
                                    const int nhelp, const int role, const unsigned mb, const int pub, const int lane) {
    constexpr int ubs = 8 * NRMAX;
    if (role == 0) {
#pragma unroll 1
        for (int jt = 0; jt < np; ++jt) {
            switch (nr - jt) {
                case 1: w0_panel<1>(A, WT, jt, np, UBp, KBp, ubs, recR, nhelp, mb, lane); break;
                case 2: if constexpr (NRMAX >= 2) w0_panel<2>(A, WT, jt, np, UBp, KBp, ubs, recR, nhelp, mb, lane); break;
                case 3: if constexpr (NRMAX >= 3) w0_panel<3>(A, WT, jt, np, UBp, KBp, ubs, recR, nhelp, mb, lane); break;
                case 4: if constexpr (NRMAX >= 4) w0_panel<4>(A, WT, jt, np, UBp, KBp, ubs, recR, nhelp, mb, lane); break;
                case 5: if constexpr (NRMAX >= 5) w0_panel<5>(A, WT, jt, np, UBp, KBp, ubs, recR, nhelp, mb, lane); break;
                case 6: if constexpr (NRMAX >= 6) w0_panel<6>(A, WT, jt, np, UBp, KBp, ubs, recR, nhelp, mb, lane); break;
                default: break;
            }
        }
    } else if (role <= nhelp) {
#if defined(BKF_SQC_NOPUBLISH) || defined(BKF_SQC_NOHELP)
        return;
#endif
        const int h = role - 1;
#pragma unroll 1
        for (int jt = 0; jt < np; ++jt) {
            const int nb = np - jt - 1;  // column blocks behind the panel
            const int nact = nhelpers(nb, nhelp);
            if (h >= nact) break;
            const int nbar = 32 * (1 + nact);
            const unsigned parity = (unsigned)(pub + jt) & 1u;
            // block i behind the panel: i = 0 -> helper 0; i = 2h - 1, 2h -> helper h >= 1
            const int c0 = h == 0 ? jt + 1 : jt + 2 * h, c1 = (h > 0 && 2 * h < nb) ? c0 + 1 : np;
            if (c1 < np) {
                switch (nr - jt) {
                    case 3: if constexpr (NRMAX >= 3) helper_panel<3, 2>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                    case 4: if constexpr (NRMAX >= 4) helper_panel<4, 2>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                    case 5: if constexpr (NRMAX >= 5) helper_panel<5, 2>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                    case 6: if constexpr (NRMAX >= 6) helper_panel<6, 2>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                    default: __trap();
                }
                continue;
            }
            switch (nr - jt) {
                case 2: if constexpr (NRMAX >= 2) helper_panel<2, 1>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                case 3: if constexpr (NRMAX >= 3) helper_panel<3, 1>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                case 4: if constexpr (NRMAX >= 4) helper_panel<4, 1>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                case 5: if constexpr (NRMAX >= 5) helper_panel<5, 1>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                case 6: if constexpr (NRMAX >= 6) helper_panel<6, 1>(A, WT, jt, np, UBp, KBp, ubs, recR, nbar, c0, c1, mb, parity, lane); break;
                default: break;
            }
        }
    }
}

template <int MT, int PT, int RT>
__global__ void __launch_bounds__(NTHR, 4) forward_kernel(KArgs<double> G) {
    using D = Dims<MT, PT, RT>;
    constexpr int m = D::m, p = D::p, NT1 = D::NT1, NRP = D::NRP;
    static_assert(D::NRMAX <= 6, "panel code exists for up to six row blocks");
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    const int n = G.n;
    const int warp = threadIdx.x >> 5;
    double *UA = smem + D::UA, *PA = smem + D::PA, *RQ = smem + D::RQ, *HC = smem + D::HC, *TT = smem + D::TT,
           *ZZ = smem + D::ZZ, *UBp = smem + D::UB, *KBp = smem + D::KB;
    double *va = smem + D::V_a, *vaf = smem + D::V_af, *vc = smem + D::V_c, *vd = smem + D::V_d, *vv = smem + D::V_v,
           *vsv = smem + D::V_sv, *vin = smem + D::V_in, *vw = smem + D::V_w, *vym = smem + D::V_ym, *vyh = smem + D::V_yh;
    volatile int* si = reinterpret_cast<volatile int*>(smem + D::SI);

    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n * p;
    // ---- set-up: warp 0 factors H and Q, the others stage T, Z, P0 and the vectors ----
    if (warp == 0) {
        bool Hz;
        const int st0 = setup_consts<typename D::W>(G, b, HC, RQ, nullptr, UA, &Hz, lane);
        if (lane == 0) {
            si[4] = Hz ? 1 : 0;
            si[5] = st0;
        }
    } else {
        const double* Tg = param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
        const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)p * m);
        const double* P0 = param_ptr(G.P0, G.shared_mask, BKF_PARAM_P0, b, (size_t)m * m);
        for (int idx = warp - 1; idx < MT * MT; idx += NWARP - 1) {
            stt(TT, idx, lane, ldg_tile(Tg, m, idx / MT, idx % MT, lane));
            stt(PA, (idx / MT) * NRP + idx % MT, lane, ldg_tile(P0, m, idx / MT, idx % MT, lane));
        }
        for (int idx = warp - 1; idx < PT * MT; idx += NWARP - 1) stt(ZZ, idx, lane, ldg_tile(Zg, m, idx / MT, idx % MT, lane));
        if (warp == 1) {
            const double* a0 = param_ptr(G.a0, G.shared_mask, BKF_PARAM_A0, b, (size_t)m);
            const double* cg = param_ptr(G.c, G.shared_mask, BKF_PARAM_C, b, (size_t)m);
            const double* dg = param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, (size_t)p);
            for (int i = lane; i < m; i += 32) { va[i] = a0[i]; vc[i] = cg[i]; }
            for (int i = lane; i < p; i += 32) vd[i] = dg[i];
        }
    }
#ifdef BKF_SQC_PROFILE
    if (threadIdx.x < 16) KBp[16 + threadIdx.x] = 0.0;
#endif
    const unsigned mb = smem_u32(smem + D::MB);
    if (threadIdx.x < 8) mbar_init(mb + 8 * threadIdx.x, 1);
    __syncthreads();
    int pub = 0;  // panels published so far (phase of the slot barriers)
    const bool Hzero = si[4] != 0;

    // mask, y_hat and innovation of step ts from va (warp 3); the record gets a
    auto innovation = [&](const int ts) {
        const size_t bt = (size_t)b * n + ts;
        bool miss = true;
        if (lane < p) {
            const double yv = yb[(size_t)ts * p + lane];
            miss = is_missing(yv, G.fill, true);
            vw[lane] = miss ? 0.0 : 1.0;
            vym[lane] = miss ? 0.0 : yv;
        }
        const int nobs = __popc(__ballot_sync(FULL, !miss));
        if (lane == 0) {
            si[ts & 1] = nobs == 0 ? 1 : 0;
            si[2 + (ts & 1)] = (nobs != 0 && nobs != p) ? 1 : 0;
        }
        __syncwarp();
#pragma unroll
        for (int I = 0; I < PT; ++I) {
            const double wI = vw[8 * I + g];
            double acc = 0.0;
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                const double2 z = ldt(ZZ, I * MT + J, lane);
                acc = fma(z.x * wI, va[8 * J + 2 * t], acc);
                acc = fma(z.y * wI, va[8 * J + 2 * t + 1], acc);
            }
            acc = red_t(acc) + vd[8 * I + g];
            if (t == 0) {
                vyh[8 * I + g] = acc;
                vv[8 * I + g] = vym[8 * I + g] - acc;
            }
        }
        __syncwarp();
        if (G.obs_mu && lane < p) G.obs_mu[bt * p + lane] = vyh[lane];
        if (G.pred_a)
            for (int i = lane; i < m; i += 32) G.pred_a[bt * m + i] = va[i];
        if (G.traj) {
            double* rec = G.traj + bt * D::RSTRIDE;
            for (int i = lane; i < m; i += 32) rec[i] = va[i];
        }
    };
    if (warp == 3 && n > 0) innovation(0);
    __syncthreads();

    double logp = 0.0;  // warp 3
#ifdef BKF_SQC_PROFILE
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tc0 = clock64();
#define BKF_TICK(i) { const long long tc1 = clock64(); tacc[i] += tc1 - tc0; tc0 = tc1; }
#else
#define BKF_TICK(i)
#endif
    for (int ts = 0; ts < n; ++ts) {
        const size_t bt = (size_t)b * n + ts;
        const bool flag = si[ts & 1] != 0, partial = si[2 + (ts & 1)] != 0;
        double* rec = G.traj ? G.traj + bt * D::RSTRIDE : nullptr;
        const bool dense_pc = ts == 0;  // P0 is whatever the caller passed; later factors are lower triangular
        // ---- S1: update array (transposed) [[Hc, Zm Pc], [0, Pc]] from Pc (= predict array's R^T) ----
        if (G.pred_P) {
            double* dst = G.pred_P + bt * m * m;
            for (int I = warp; I < MT; I += NWARP)
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    double2 acc = zero2();
#pragma unroll
                    for (int K = 0; K < MT; ++K) mma_nt(acc, ldt(PA, I * NRP + K, lane), ldt(PA, J * NRP + K, lane));
                    stg_tile(dst, m, I, J, lane, acc);
                }
        }
        if (!flag) {
            for (int idx = warp; idx < PT * PT; idx += NWARP) {
                double2 h = ldt(HC, idx, lane);
                if (Hzero) h = zero2();
                else if (partial) h = make_double2(NAN, NAN);  // the reference's cholesky(W H) fails here
                stt(UA, (idx / PT) * NT1 + idx % PT, lane, h);
            }
            for (int idx = warp; idx < PT * MT; idx += NWARP) {
                const int I = idx / MT, J = idx % MT;
                const double wI = vw[8 * I + g];
                double2 acc = zero2();
                for (int K = dense_pc ? 0 : J; K < MT; ++K) {
                    double2 z = ldt(ZZ, I * MT + K, lane);
                    z.x *= wI;
                    z.y *= wI;
                    mma_nt(acc, z, ttrans(ldt(PA, K * NRP + J, lane), lane));  // Zm[I][K] Pc[K][J]
                }
                stt(UA, I * NT1 + PT + J, lane, acc);
            }
            for (int idx = warp; idx < MT * PT; idx += NWARP) stt(UA, (PT + idx / PT) * NT1 + idx % PT, lane, zero2());
        }
        for (int idx = warp; idx < MT * MT; idx += NWARP)  // (a skipped update keeps Pc: Pcf := Pc)
            stt(UA, (PT + idx / MT) * NT1 + PT + idx % MT, lane, ldt(PA, (idx / MT) * NRP + idx % MT, lane));
        BKF_TICK(0)
        __syncthreads();
        BKF_TICK(1)
        // ---- S2: R factor of the update array ----
        if (!flag) {
            qr_cta<D::NRMAX>(UA, NT1, NT1, NT1, UBp, KBp, rec ? rec + D::OFF_RF : nullptr, NWARP - 1, warp, mb, pub, lane);
            pub += NT1 - 1;
        }
        BKF_TICK(2)
        __syncthreads();
        BKF_TICK(3)
        // ---- S3: predict array (transposed) [T (Pcf + jitter I), R Qc] ----
        {
            const bool dense_pf = flag && dense_pc;
            for (int idx = warp; idx < MT * MT; idx += NWARP) {
                const int I = idx % MT, J = idx / MT;
                double2 acc = zero2();
                for (int K = dense_pf ? 0 : J; K < MT; ++K) {
                    double2 pf = ldt(UA, (PT + K) * NT1 + PT + J, lane);
                    if (K == J && (g >> 1) == t) {  // jitter on the factor (quirk Q4)
                        if (g & 1) pf.y += G.jitter;
                        else pf.x += G.jitter;
                    }
                    mma_nt(acc, ldt(TT, I * MT + K, lane), ttrans(pf, lane));
                }
                stt(PA, I * NRP + J, lane, acc);
            }
            for (int idx = warp; idx < MT * RT; idx += NWARP) stt(PA, (idx / RT) * NRP + MT + idx % RT, lane, ldt(RQ, idx, lane));
        }
        BKF_TICK(4)
        __syncthreads();
        BKF_TICK(5)
        // ---- S4: R factor of the predict array (warps 0-2)  ||  everything else of the step (warp 3) ----
        if (warp < NWARP - 1) {
            qr_cta<D::NRMAX>(PA, NRP, MT, NRP, UBp, KBp, rec ? rec + D::OFF_RP : nullptr, NWARP - 2, warp, mb, pub, lane);
        } else {
            auto pcf = [&](int I, int K) {  // (Pcf + jitter I) tile
                double2 v = ldt(UA, (PT + I) * NT1 + PT + K, lane);
                if (I == K && (g >> 1) == t) {
                    if (g & 1) v.y += G.jitter;
                    else v.x += G.jitter;
                }
                return v;
            };
            double ll = 0.0;
            if (!flag) {
                if (G.obs_cov)  // F = Fc Fc^T, Fc = B[:p,:p]
                    store_square<PT>(G.obs_cov + bt * p * p, [&](int I, int K) { return ldt(UA, I * NT1 + K, lane); }, lane);
                // ---- s = Fc^{-1} v, inner = Fc^{-1} s  (Fc lower), lane q owns x[q] ----
                {
                    const int q = lane < p ? lane : p - 1;
                    double frow[p];
#pragma unroll
                    for (int i = 0; i < p; ++i) frow[i] = UA[elem_addr(q, i, NT1)];
                    const double dq = UA[elem_addr(q, q, NT1)];
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
                // a_f = a + KF s
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    double acc = 0.0;
#pragma unroll
                    for (int J = 0; J < PT; ++J) {
                        const double2 kf = ldt(UA, (PT + I) * NT1 + J, lane);
                        acc = fma(kf.x, vsv[8 * J + 2 * t], acc);
                        acc = fma(kf.y, vsv[8 * J + 2 * t + 1], acc);
                    }
                    acc = red_t(acc);
                    if (t == 0) vaf[8 * I + g] = va[8 * I + g] + acc;
                }
            } else {
                for (int i = lane; i < m; i += 32) vaf[i] = va[i];
                if (G.obs_cov)  // every observation missing: the masked array has zero columns there, F_chol = 0
                    for (int i = lane; i < p * p; i += 32) G.obs_cov[bt * p * p + i] = 0.0;
            }
            __syncwarp();
            if (G.filt_a)
                for (int i = lane; i < m; i += 32) G.filt_a[bt * m + i] = vaf[i];
            if (G.filt_P) store_square<MT>(G.filt_P + bt * m * m, pcf, lane);
            if (G.ll && lane == 0) G.ll[bt] = ll;
            logp += ll;
            // a+ = T a_f + c
            {
                double an[MT];
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    an[I] = 0.0;
#pragma unroll
                    for (int K = 0; K < MT; ++K) {
                        const double2 tf = ldt(TT, I * MT + K, lane);
                        an[I] = fma(tf.x, vaf[8 * K + 2 * t], an[I]);
                        an[I] = fma(tf.y, vaf[8 * K + 2 * t + 1], an[I]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    const double s = red_t(an[I]) + vc[8 * I + g];
                    if (t == 0) va[8 * I + g] = s;
                }
                __syncwarp();
            }
            if (ts + 1 < n) innovation(ts + 1);
        }
        pub += MT - 1;
        BKF_TICK(6)
        __syncthreads();
        BKF_TICK(7)
    }
#ifdef BKF_SQC_PROFILE
    __syncthreads();
    if (b == 0 && lane == 0 && G.ll) {
        for (int i = 0; i < 8; ++i) G.ll[warp * 8 + i] = (double)tacc[i];
        if (warp == 0)
            for (int i = 0; i < 9; ++i) G.ll[32 + i] = KBp[16 + i];
    }
#endif
    if (warp == NWARP - 1 && lane == 0) {
        int st = si[5];
        if (!isfinite(logp)) st |= BKF_STATUS_NONFINITE;
        if (G.logp) G.logp[b] = logp;
        if (G.status) G.status[b] = st;
    }
}


// ===============================================================================================================
// Adjoint sweep, four warps per series.  Same mathematics and the same record as sqw::backward_kernel
// (bkf_sqrt_warp.cuh: Abar = A Psi, Psi = R^{-1} copyltu(R Rbar^T) R^{-T}; [a | Pc | triu(R_update) | triu(R_predict)]
// from either forward family); every tile phase is split over the warps by tile ROWS (warp (row + b) mod 4, so that
// co-resident series load different sub-partitions), with a CTA barrier between phases.  The right-solves of Psi are
// independent per tile row, so a pass needs no barrier inside.

template <int NT, class RbF>
__device__ __forceinline__ void psi_tiles_cta(double* __restrict__ W, const double* __restrict__ RF,
                                              double* __restrict__ RINV, RbF rbar, const int warp, const int rot,
                                              const int lane) {
    const int g = lane >> 2, t = lane & 3;
    auto own = [&](int I) { return ((I + rot) & (NWARP - 1)) == warp; };
    // S = copyltu(R Rbar^T), row by row; the inverse of the row's diagonal tile of R on lanes 0..7
#pragma unroll
    for (int I = 0; I < NT; ++I) {
        if (!own(I)) continue;
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
        if (lane < 8) {
            const int c = lane;
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
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        // X <- X R^{-T}: tile row J, tile column by tile column (the rows are independent)
#pragma unroll
        for (int J = 0; J < NT; ++J) {
            if (!own(J)) continue;
#pragma unroll
            for (int I = NT - 1; I >= 0; --I) {
                Acc2 acc = acc2_from(ldt(W, J * NT + I, lane));
#pragma unroll
                for (int K = I + 1; K < NT; ++K) mma_nt(acc, ldt(W, J * NT + K, lane), neg(ldt(RF, tri<NT>(I, K), lane)));
                Acc2 o = acc2_zero();
                mma_nt(o, fin(acc), ldt(RINV, I, lane));
                stt(W, J * NT + I, lane, fin(o));
                __syncwarp();
            }
        }
        __syncthreads();
        if (pass == 0) {
#pragma unroll
            for (int I = 0; I < NT; ++I) {
                if (!own(I)) continue;
#pragma unroll
                for (int J = 0; J < I; ++J) {
                    const double2 lo = ldt(W, I * NT + J, lane), up = ldt(W, J * NT + I, lane);
                    stt(W, I * NT + J, lane, ttrans(up, lane));
                    stt(W, J * NT + I, lane, ttrans(lo, lane));
                }
                stt(W, I * NT + I, lane, ttrans(ldt(W, I * NT + I, lane), lane));
            }
            __syncthreads();
        }
    }
}

// tiles of the upper triangle of an n x n factor from its packed rows, split over the warps (see load_packed_upper)
template <int NT>
__device__ __forceinline__ void load_packed_upper_cta(double* __restrict__ RF, const double* __restrict__ src,
                                                      const int warp, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int n = 8 * NT;
    int idx = 0;
#pragma unroll
    for (int I = 0; I < NT; ++I)
#pragma unroll
        for (int K = I; K < NT; ++K, ++idx) {
            if ((idx & (NWARP - 1)) != warp) continue;
            const int i = 8 * I + g, j = 8 * K + 2 * t;
            const double* q = src + (i * n - (i * (i - 1)) / 2 - i + j);
            double2 v;
            v.x = j >= i ? q[0] : 0.0;
            v.y = j + 1 >= i ? q[1] : 0.0;
            stt(RF, tri<NT>(I, K), lane, v);
        }
}

template <int MT, int PT, int RT>
struct BDimsC : sqw::BDims<MT, PT, RT> {
    using B0 = sqw::BDims<MT, PT, RT>;
    static constexpr int SI = B0::ELEMS;      // 4 ints: flag, partial, Hzero
    static constexpr int ELEMS = B0::ELEMS + 2;
    static constexpr int RINV2 = B0::PBT;     // inverses of the diagonal tiles of the update factor: Pbar^T is dead then
    static constexpr int AB21 = B0::RF;       // Abar21: the factor is dead then
    static_assert(B0::NT1 <= MT * MT && MT * PT <= B0::RFTILES && MT * PT <= B0::RBTILES, "scratch blocks fit");
};

template <int MT, int PT, int RT>
__global__ void __launch_bounds__(NTHR, 4) backward_kernel(KArgs<double> G) {
    using D = sqw::Dims<MT, PT, RT>;
    using BD = BDimsC<MT, PT, RT>;
    constexpr int m = D::m, p = D::p, r = D::r, N1 = D::N1, NT1 = D::NT1;
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    const int n = G.n;
    const int rot = b & (NWARP - 1);
    auto own = [&](int I) { return ((I + rot) & (NWARP - 1)) == warp; };
    double *W = smem + BD::W, *RF = smem + BD::RF, *RB = smem + BD::RB, *PBT = smem + BD::PBT, *HC = smem + BD::HC,
           *HBT = smem + BD::HBT;
    double *va = smem + BD::V_a, *vaf = smem + BD::V_af, *vab = smem + BD::V_ab, *vafb = smem + BD::V_afb,
           *vgc = smem + BD::V_gc, *vv = smem + BD::V_v, *vs = smem + BD::V_s, *vin = smem + BD::V_in, *vw = smem + BD::V_w,
           *vym = smem + BD::V_ym, *vvb = smem + BD::V_vb, *vsb = smem + BD::V_sb, *vt1 = smem + BD::V_t1,
           *vt2 = smem + BD::V_t2, *vgd = smem + BD::V_gd, *vd = smem + BD::V_d;
    volatile int* si = reinterpret_cast<volatile int*>(smem + BD::SI);

    const double* Tg = param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
    const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)p * m);
    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n * p;
    const size_t tstride = sqw::traj_record_elems(m, p);
    const size_t off_rf = (size_t)m + (size_t)m * m, off_rp = off_rf + (size_t)N1 * (N1 + 1) / 2;
    // add-only accumulators in global memory (L2), as in sqw::backward_kernel
    double* scr = G.traj + (size_t)G.B * n * tstride + (size_t)b * tstride;
    double* gT = G.g_T ? G.g_T + (size_t)b * m * m : scr;
    double* gZ = G.g_Z ? G.g_Z + (size_t)b * p * m : scr + m * m;
    double* PS = scr + m * m + p * m;
    const bool want_ps = G.g_R || G.g_Q;

    if (warp == 0) {
        bool Hz;
        setup_consts<D>(G, b, HC, nullptr, nullptr, W, &Hz, lane);
        if (lane == 0) si[2] = Hz ? 1 : 0;
    } else {
        const int tid = threadIdx.x - 32;
        const double* dg = param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, (size_t)p);
        for (int i = tid; i < p; i += NTHR - 32) { vd[i] = dg[i]; vgd[i] = 0.0; }
        for (int i = tid; i < m; i += NTHR - 32) { vab[i] = 0.0; vgc[i] = 0.0; }
        for (int i = tid; i < MT * MT * 32; i += NTHR - 32) reinterpret_cast<double2*>(PBT)[i] = zero2();
        for (int i = tid; i < PT * PT * 32; i += NTHR - 32) reinterpret_cast<double2*>(HBT)[i] = zero2();
        for (int i = tid; i < m * m; i += NTHR - 32) { gT[i] = 0.0; PS[i] = 0.0; }
        for (int i = tid; i < p * m; i += NTHR - 32) gZ[i] = 0.0;
    }
    __syncthreads();
    const bool Hzero = si[2] != 0;

#ifdef BKF_SQC_PROFILE
    long long tb[16], tcb = clock64();
    for (int i = 0; i < 16; ++i) tb[i] = 0;
#define BKF_TOCK(i) { const long long tc1 = clock64(); tb[i] += tc1 - tcb; tcb = tc1; }
#else
#define BKF_TOCK(i)
#endif
    for (int ts = n - 1; ts >= 0; --ts) {
        const size_t bt = (size_t)b * n + ts;
        const double* rec = G.traj + bt * tstride;
        const double* Pcg = rec + m;
        const double gcot = G.ll_cot ? G.ll_cot[bt] : 1.0;
        const bool dense_pc = ts == 0;  // P0 is whatever the caller passed; later factors are lower triangular
        if (ts > 0) {  // the next record: pull it towards L2 while this step computes
            const char* nxt = reinterpret_cast<const char*>(rec - tstride);
            for (int off = threadIdx.x * 128; off < (int)(tstride * sizeof(double)); off += NTHR * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
        }
        // ---- mask (warp 0), a, predict factor ----
        if (warp == 0) {
            bool miss = true;
            if (lane < p) {
                const double yv = yb[(size_t)ts * p + lane];
                miss = is_missing(yv, G.fill, true);
                vw[lane] = miss ? 0.0 : 1.0;
                vym[lane] = miss ? 0.0 : yv;
            }
            const int nobs = __popc(__ballot_sync(FULL, !miss));
            if (lane == 0) {
                si[0] = nobs == 0 ? 1 : 0;
                si[1] = (nobs != 0 && nobs != p) ? 1 : 0;
            }
        } else if (warp == 1) {
            for (int i = lane; i < m; i += 32) va[i] = rec[i];
        }
        load_packed_upper_cta<MT>(RF, rec + off_rp, warp, lane);
        __syncthreads(); BKF_TOCK(0)
        const bool flag = si[0] != 0, partial = si[1] != 0;
        if (warp == NWARP - 1) {  // innovation v = y - (Zm a + d): needed by the solves further down
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
        }
        // ================= predict adjoint =================
        psi_tiles_cta<MT>(W, RF, RB, [&](int J, int K) {
            double2 v = ldt(PBT, J * MT + K, lane);
            if (J == K) {
                if (2 * t < g) v.x = 0.0;
                if (2 * t + 1 < g) v.y = 0.0;
            }
            return v;
        }, warp, rot, lane);
        BKF_TOCK(14)
        if (want_ps) {
#pragma unroll
            for (int I = 0; I < MT; ++I)
                if (own(I))
#pragma unroll
                    for (int J = 0; J < MT; ++J) red_tile(PS, m, I, J, lane, ldt(W, I * MT + J, lane));
        }
        // ---- forward quantities of this step from the record: s, inner, a_f, Pcf ----
        if (!flag) load_packed_upper_cta<NT1>(RF, rec + off_rf, warp, lane);  // (the predict factor is dead)
        __syncthreads(); BKF_TOCK(1)
        if (warp == NWARP - 1) {
            if (!flag) {
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
        }
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
        constexpr int TPT = MT * MT;             // tile offset of the second scratch block inside W
        // TP^T = (T Pcf)^T = Pcf^T T^T :  [I][J] = sum_K pft(I,K) (T[J][K])^T, tile column J by warp own(J)
#pragma unroll
        for (int J = 0; J < MT; ++J) {
            if (!own(J)) continue;
            double2 acc[MT];
#pragma unroll
            for (int I = 0; I < MT; ++I) acc[I] = zero2();
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                const double2 tk = ldg_tile(Tg, m, J, K, lane);
#pragma unroll
                for (int I = 0; I < MT; ++I)
                    if (dense_pf || K >= I) mma_nt(acc[I], pft(I, K), tk);
            }
#pragma unroll
            for (int I = 0; I < MT; ++I) stt(W, TPT + I * MT + J, lane, acc[I]);
        }
        __syncthreads(); BKF_TOCK(2)
        // U1^T = TP^T Psi_p  -> RB[0 .. MT*MT), tile row I by warp own(I); then U1 -> W[TPT ..)
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (!own(I)) continue;
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
        __syncthreads(); BKF_TOCK(3)
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (!own(I)) continue;
#pragma unroll
            for (int J = 0; J < MT; ++J) stt(W, TPT + I * MT + J, lane, ttrans(ldt(RB, J * MT + I, lane), lane));
        }
        __syncthreads(); BKF_TOCK(4)
        // gT += U1 Pcf^T + abar af^T (tile row I by warp own(I));  Pcf[J][K] = (pft(K, J))^T
        if (G.g_T) {
#pragma unroll
            for (int I = 0; I < MT; ++I) {
                if (!own(I)) continue;
                const double ab = vab[8 * I + g];
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    Acc2 acc = acc2_from(make_double2(ab * vaf[8 * J + 2 * t], ab * vaf[8 * J + 2 * t + 1]));
#pragma unroll
                    for (int K = 0; K < MT; ++K)
                        if (dense_pf || K <= J) mma_nt(acc, ldt(W, TPT + I * MT + K, lane), ttrans(pft(K, J), lane));
                    red_tile(gT, m, I, J, lane, fin(acc));
                }
            }
        }
        // Pcfbar^T = U1^T T :  [I][J] = sum_K U1T[I][K] ((T^T)[J][K])^T  -> W[0 .. MT*MT)   (Psi_p is dead), row I by own(I)
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (!own(I)) continue;
            Acc2 acc[MT];
#pragma unroll
            for (int J = 0; J < MT; ++J) acc[J] = acc2_zero();
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                const double2 uk = ldt(RB, I * MT + K, lane);
#pragma unroll
                for (int J = 0; J < MT; ++J) mma_nt(acc[J], uk, ldg_tile_t(Tg, m, J, K, lane));
            }
#pragma unroll
            for (int J = 0; J < MT; ++J) stt(W, I * MT + J, lane, fin(acc[J]));
        }
        // afbar = T^T abar (tile column J of T by warp own(J)); gc += abar afterwards
#pragma unroll
        for (int J = 0; J < MT; ++J) {
            if (!own(J)) continue;
            double afp = 0.0;
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                const double2 tk = ldg_tile_t(Tg, m, J, K, lane);  // [g][2t+e] = T[8K + 2t + e][8J + g]
                afp = fma(tk.x, vab[8 * K + 2 * t], fma(tk.y, vab[8 * K + 2 * t + 1], afp));
            }
            afp = red_t(afp);
            if (t == 0) vafb[8 * J + g] = afp;
        }
        __syncthreads(); BKF_TOCK(5)
        for (int i = threadIdx.x; i < m; i += NTHR) vgc[i] += vab[i];
        if (flag) {  // degenerate branch (:700-705): (a_f, Pcf, ll) = (a, Pc, 0)
            for (int i = threadIdx.x; i < m; i += NTHR) vab[i] = vafb[i];
            for (int i = warp; i < MT * MT; i += NWARP) stt(PBT, i, lane, ldt(W, i, lane));
            __syncthreads(); BKF_TOCK(6)
            continue;
        }
        // ================= update adjoint =================
        if (warp == NWARP - 1) {
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
            const double t1 = usolve(lane < p ? lossbar * vv[lane] : 0.0);  // Fc^{-T} innerbar
            const double t2 = usolve(lane < p ? vsb[lane] + t1 : 0.0);      // Fc^{-T} (sbar + t1)
            vbar += t2;
            if (lane < p) {
                vt1[lane] = t1;
                vt2[lane] = t2;
                vvb[lane] = vbar;
                vsb[lane] = logdetbar * 2.0 * rdiag;  // diagonal term of Fcbar
            }
        }
        __syncthreads(); BKF_TOCK(7)
        // Rbar = Bbar^T (upper):  [Fcbar^T, KFbar^T; 0, triu(Pcfbar^T)], tile row by warp own(row)
#pragma unroll
        for (int I = 0; I < PT; ++I) {
            if (!own(I)) continue;
            const int i = 8 * I + g;
            const double ini = vin[i], sv = vs[i];
#pragma unroll
            for (int K = I; K < PT; ++K) {
                const int k = 8 * K + 2 * t;
                double2 v;
                v.x = k >= i ? fma(-vt1[k], ini, -vt2[k] * sv) + (k == i ? vsb[i] : 0.0) : 0.0;
                v.y = k + 1 >= i ? fma(-vt1[k + 1], ini, -vt2[k + 1] * sv) + (k + 1 == i ? vsb[i] : 0.0) : 0.0;
                stt(RB, tri<NT1>(I, K), lane, v);
            }
#pragma unroll
            for (int K = 0; K < MT; ++K)
                stt(RB, tri<NT1>(I, PT + K), lane, make_double2(vafb[8 * K + 2 * t] * sv, vafb[8 * K + 2 * t + 1] * sv));
        }
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (!own(PT + I)) continue;
#pragma unroll
            for (int K = I; K < MT; ++K) {
                double2 v = ldt(W, I * MT + K, lane);
                if (I == K) {
                    if (2 * t < g) v.x = 0.0;
                    if (2 * t + 1 < g) v.y = 0.0;
                }
                stt(RB, tri<NT1>(PT + I, PT + K), lane, v);
            }
        }
        __syncthreads(); BKF_TOCK(8)
        // Psi of the update array.  RINV must not overlap the cotangent tiles while S0 is formed: psi_tiles reads rbar only
        // in its first phase, and RINV sits in a block of its own here (behind the vectors would be too small): W2 below.
        psi_tiles_cta<NT1>(W, RF, smem + BD::RINV2, [&](int J, int K) { return ldt(RB, tri<NT1>(J, K), lane); }, warp, rot, lane);
        BKF_TOCK(15)
        // ---- Abar = A Psi, A = [[Hc^T, 0], [A21, Pc^T]],  A21 = (Zm Pc)^T = Pc^T Zm^T -> RB[0 ..), row I by own(I) ----
        constexpr int A21T = 0;  // the cotangent tiles are dead
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (!own(I)) continue;
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
        // Abar11 = Hc^T Psi11  (accumulated: Hcbar^T), row I by own(I)
#pragma unroll
        for (int I = 0; I < PT; ++I) {
            if (!own(I)) continue;
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
        __syncthreads();  // (a row of [Abar21, Abar22] needs its own row of A21 only, but RF is overwritten below)
        // rows of [Abar21, Abar22]; Abar21 -> RF[0 ..) (the factor is dead), Pbar^T = Abar22 + Abar21 Zm -> PBT
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (!own(I)) continue;
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
            for (int J = 0; J < PT; ++J) stt(smem + BD::AB21, I * PT + J, lane, acc[J]);
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
        __syncthreads(); BKF_TOCK(9)
        // gZ += W (ZPbar Pc^T - vbar a^T),  ZPbar = Abar21^T, tile column J by own(J)
        if (G.g_Z) {
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                if (!own(J)) continue;
                double2 acc[PT];
#pragma unroll
                for (int Io = 0; Io < PT; ++Io) acc[Io] = zero2();
#pragma unroll
                for (int K = 0; K < MT; ++K)
                    if (dense_pc || K <= J) {
                        const double2 pc = ldg_tile(Pcg, m, J, K, lane);
#pragma unroll
                        for (int Io = 0; Io < PT; ++Io)
                            mma_nt(acc[Io], ttrans(ldt(smem + BD::AB21, K * PT + Io, lane), lane), pc);
                    }
#pragma unroll
                for (int Io = 0; Io < PT; ++Io) {
                    const double wo = vw[8 * Io + g], vb = vvb[8 * Io + g];
                    red_tile(gZ, m, Io, J, lane, make_double2(wo * fma(-vb, va[8 * J + 2 * t], acc[Io].x),
                                                              wo * fma(-vb, va[8 * J + 2 * t + 1], acc[Io].y)));
                }
            }
        }
        // abar = afbar - Zm^T vbar (tile column J by own(J)); gd -= vbar
#pragma unroll
        for (int J = 0; J < MT; ++J) {
            if (!own(J)) continue;
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
        if (warp == 0 && lane < p) vgd[lane] -= vvb[lane];
        __syncthreads(); BKF_TOCK(10)
    }
    // ---- gradients out ----
    for (int i = threadIdx.x; i < m; i += NTHR) {
        if (G.g_a0) G.g_a0[(size_t)b * m + i] = vab[i];
        if (G.g_c) G.g_c[(size_t)b * m + i] = vgc[i];
    }
    if (G.g_d && threadIdx.x < p) G.g_d[(size_t)b * p + threadIdx.x] = vgd[threadIdx.x];
    if (G.g_P0) {
        double* dst = G.g_P0 + (size_t)b * m * m;
        for (int idx = warp; idx < MT * MT; idx += NWARP) {
            const int I = idx / MT, J = idx % MT;
            const double2 v = ldt(PBT, I * MT + J, lane);  // Pbar[8J+2t+e][8I+g]
            dst[(8 * J + 2 * t) * m + 8 * I + g] = v.x;
            dst[(8 * J + 2 * t + 1) * m + 8 * I + g] = v.y;
        }
    }
    __threadfence();  // the reductions into PS / gT / gZ were issued by all warps; the epilogue below reads PS
    __syncthreads();
#ifdef BKF_SQC_PROFILE
    if (b == 0 && lane == 0 && G.g_T)
        for (int i = 0; i < 16; ++i) G.g_T[(size_t)b * m * m + 64 + warp * 16 + i] = (double)tb[i];
    __syncthreads();
#endif
    if (warp != 0) return;
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

}  // namespace sqc
}  // namespace bkf
