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
// takes the CTA-per-series kernels.  This file holds the forward sweep and the device helpers shared with its adjoint,
// the four-warps-per-series kernel of bkf_sqrt_bwd4.cuh, which reads the record the sweep leaves: per (series, step)
// [a | a_f | v | s | inner | tiles of Rf^T | tiles of Rp^T] with the tiles exactly as they lie in the arena.
#pragma once
#include <cstdio>

#include "bkf_common.cuh"

namespace bkf {
namespace sqw {

constexpr unsigned FULL = 0xffffffffu;
struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

#ifndef BKF_PANEL  // the reflector loop of a panel: panel_reflectors (plain), panel_reflectors_la (look-ahead)
#define BKF_PANEL panel_reflectors_la
#endif

#ifdef BKF_SQW_PROFILE  // clock64 phase timers of block 0 (tools/build_variant.py sqwprof "-DBKF_SQW_PROFILE" ...)
__device__ unsigned long long g_prof[16];
#define BKF_PMARK(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { const long long _n = clock64(); g_prof[i] += (unsigned long long)(_n - _pm); _pm = _n; } } while (0)
#define BKF_PTICK() const long long _pt0 = clock64()
#define BKF_PTOCK(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_prof[i] += (unsigned long long)(clock64() - _pt0); } while (0)
#else
#define BKF_PTICK() do { } while (0)
#define BKF_PTOCK(i) do { } while (0)
#define BKF_PMARK(i) do { } while (0)
#endif

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
// Sum over the four lanes of a group on the FP64 tensor pipe: in the m8n8k4 fragments lane (g, t) holds A[g][t], so
// A x ones hands every lane of group g the group's sum -- one DMMA (26 cycles) instead of two shuffle + add rounds (70).
__device__ __forceinline__ double red4_mma(double x) {
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, x, 1.0);
    return c0;
}

// Scalars of one Householder reflector (dlarfg) from x2 = alpha^2 + |x|^2:
//   beta = -sign(alpha) nrm,  tau = (beta - alpha) / beta = (|alpha| + nrm) / nrm,  scal = 1 / (alpha - beta).
// 1 / sqrt and 1 / x come from the 20-bit hardware seeds and two Newton steps each (a quarter of the instructions and half
// the latency of the IEEE sequences: measured dependent latencies sqrt 91, div 71 cycles).  The reciprocal of
// |alpha| + nrm starts from a seed taken at the rsqrt SEED's accuracy (both seeds carry ~20 bits, two Newton steps each
// reach full precision), so the second hardware approximation overlaps the Newton steps of the first.
// (x2 subnormal, a column of magnitude < 1e-154, is out of range.)
struct Refl {
    double beta, tau, scal, ts;
};
__device__ __forceinline__ Refl reflector(const double x2, const double alpha) {
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
    Refl R;
    R.beta = copysign(nrm, -alpha);
    R.tau = den * rn;
    R.scal = copysign(rden, alpha);
    R.ts = R.tau * R.scal;
    return R;
}

// The eight reflectors of one panel (NQ active row blocks, the diagonal block first; lane group g owns column g, lane
// (g, t) its rows 2t, 2t+1 of every block).  Per reflector: the pivot column goes through 64 NQ bytes of shared memory
// to every group (NQ stores by four lanes + NQ broadcast loads: a quarter of the MIO slots of 4 NQ shuffles, which cost
// 4.3 cycles each), ONE reduction window -- a DMMA against ones -- yields the pivot norm, the eight v^T a_c AND, in the
// groups whose reflector is finished, the inner products v_i^T v_k of the compact-WY Gram matrix (same formula, free).
// On return pa holds V^T (unit diagonal, scaled), Rd the diagonal tile of R, tq[8 i + l] = v_i^T v_l (i < l).
// scratch: xb = 8 NQ doubles.
template <int NQ>
__device__ __forceinline__ void panel_reflectors(double2 (&pa)[NQ], double2& Rd, double& tauv, double* tq, double* xbuf,
                                                 const int lane) {
    // (no __restrict__ on xbuf: the compiler would forward a lane's own conditional stores to its loads across the
    //  __syncwarp and never see the pivot group's)
    const int g = lane >> 2, t = lane & 3;
    tauv = 0.0;
    double scalv = 1.0;  // tau and 1 / (alpha - beta) of the reflector of column g (tau = 0: H = I)
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
        // x = column k of the panel below its diagonal, handed to every group
        if (g == k) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) reinterpret_cast<double2*>(xbuf)[q * 4 + t] = pa[q];
        }
        __syncwarp();
        double2 xb[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) xb[q] = reinterpret_cast<const double2*>(xbuf)[q * 4 + t];
        const double alpha = xbuf[k];  // [row k][col k]
        const double akc = __shfl_sync(FULL, (k & 1) ? pa[0].y : pa[0].x, 4 * g + (k >> 1));  // [row k][col g]
        double2 xm = xb[0];
        if (2 * t <= k) xm.x = 0.0;
        if (2 * t + 1 <= k) xm.y = 0.0;
        double d0 = xm.x * pa[0].x, d1 = xm.y * pa[0].y;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            d0 = fma(xb[q].x, pa[q].x, d0);
            d1 = fma(xb[q].y, pa[q].y, d1);
        }
        const double dsum = d0 + d1;
        // |x|^2 != 0 (a sum of squares in group k); the vote also orders the loads of xbuf before its next stores
        const bool nz = __any_sync(FULL, g == k && dsum != 0.0);
        if (!nz) {  // dlarfg: tau = 0, H = I
            if (g < k && t == 0) tq[g * 8 + k] = 0.0;
            continue;
        }
        // one reduction window: group k -> ss = x.x; group c > k -> x.a_c; group i < k -> x_i.x_k
        const double dot = red4_mma(dsum);
        const double ss = __shfl_sync(FULL, dot, 4 * k);
        const Refl R = reflector(fma(alpha, alpha, ss), alpha);
        const double zz = fma(R.scal, dot, akc);  // v_k . a_c (c > k);  scal_i zz = v_i . v_k (i < k)
        if (g < k && t == 0) tq[g * 8 + k] = scalv * zz;
        // Update with the UNSCALED reflector u = [alpha - beta; x] (v = u scal):  a_c -= (tau zz scal) u  for c > k.
        // The pivot column keeps x until the sweep of the panel is over (nothing reads v before) and is scaled once.
        const bool piv = g == k;
        const double w = g > k ? R.ts * zz : 0.0;
        double2 ud = xm;  // u on my rows of the diagonal block
        if (2 * t == k) ud.x = alpha - R.beta;
        if (2 * t + 1 == k) ud.y = alpha - R.beta;
        {
            const double nx = fma(-w, ud.x, pa[0].x), ny = fma(-w, ud.y, pa[0].y);
            pa[0].x = (piv && 2 * t == k) ? R.beta : nx;
            pa[0].y = (piv && 2 * t + 1 == k) ? R.beta : ny;
        }
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            pa[q].x = fma(-w, xb[q].x, pa[q].x);
            pa[q].y = fma(-w, xb[q].y, pa[q].y);
        }
        if (piv) {
            tauv = R.tau;
            scalv = R.scal;
        }
    }
    // R (zeros below the diagonal) and V^T = (x scal)^T with unit diagonal
    const double2 r = pa[0];
    Rd = make_double2(2 * t <= g ? r.x : 0.0, 2 * t + 1 <= g ? r.y : 0.0);
    pa[0].x = 2 * t > g ? r.x * scalv : (2 * t == g ? 1.0 : 0.0);
    pa[0].y = 2 * t + 1 > g ? r.y * scalv : (2 * t + 1 == g ? 1.0 : 0.0);
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
        pa[q].x *= scalv;
        pa[q].y *= scalv;
    }
}

// The panel with BOTH serial pieces taken off the data path (measured on the C4 shape: broadcast of the pivot column
// 137, reflector scalars 117, zero-column vote + branch 70, norm shuffle 34 of 525 cycles per reflector):
//   * every lane keeps a private copy x of the CURRENT pivot column (its rows 2t, 2t+1).  The column after it travels
//     through shared memory one reflector EARLY -- group k + 2 stores its column right after update k, everybody loads it
//     during window k + 1 -- and turns into the next x by the same rank-one update the owning group applies
//     (x_{k+1} = y - w_{k+1} u_k, bit-identical to the owner's), so no store -> load round trip sits between an update and
//     the inner products of the next reflector;
//   * the scalars of reflector k + 1 (rsqrt, reciprocal, Newton steps) need only |a_{k+1}[rows >= k+1]|^2 and the pivot
//     element, and both are known as soon as the coefficient w of reflector k is: H_k is orthogonal on rows >= k, so the
//     squared norm of every remaining column shrinks by exactly R[k][c]^2 (the partial-norm downdate of xGEQP3), and the
//     new pivot element is one FMA.  Every lane evaluates them redundantly (no broadcast on the chain), so the scalar
//     chain of reflector k + 1 runs beside window k + 1 instead of behind it.  Downdated norms lose digits when a column
//     is nearly in the span of the earlier ones: a column whose remaining norm^2 fell below 2^-6 of its explicit value
//     at the start of the panel takes the explicit norm from the window (bound: 8 downdates x 64 x a few ulps, < 1e-13
//     relative on the norm); exact zero columns -- dlarfg's tau = 0 -- are detected on the explicit partial sums, and
//     both tests are votes whose results are only selected at the join (no branch on the chain).
// Measured (tools/refl_bench.cu, one warp per scheduler): 528 -> 408 cycles per reflector at six row blocks; what is left
// is the warp's own issue time (~60 FP64 instructions at two cycles each, 14 shared-memory and 10 shuffle instructions).
// Scratch: xbuf = 2 x 8 NQ doubles (double buffer of the travelling column).
template <int NQ>
__device__ __forceinline__ void panel_reflectors_la(double2 (&pa)[NQ], double2& Rd, double& tauv, double* tq,
                                                    double* xbuf, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    tauv = 0.0;
    double scalv = 1.0;
    double n2;  // |a_g[rows >= k]|^2: explicit at k = 0, then downdated by R[k][g]^2
    {
        double s0 = pa[0].x * pa[0].x, s1 = pa[0].y * pa[0].y;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            s0 = fma(pa[q].x, pa[q].x, s0);
            s1 = fma(pa[q].y, pa[q].y, s1);
        }
        n2 = red4_mma(s0 + s1);
    }
    const double n2ref = n2 * (1.0 / 64.0);
    double2* xb2 = reinterpret_cast<double2*>(xbuf);
    constexpr int XS = 4 * NQ;  // double2 per buffer
    // columns 0 and 1 as they are
    if (g < 2) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) xb2[g * XS + q * 4 + t] = pa[q];
    }
    __syncwarp();
    double2 x[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) x[q] = xb2[q * 4 + t];
    double alpha = xbuf[0];
    Refl R = reflector(__shfl_sync(FULL, n2, 0), alpha);
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
        const int kn = (k + 1) & 7;  // the next pivot (k = 7: results unused)
        const double2* yb = xb2 + (kn & 1) * XS;
        double2 y[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) y[q] = yb[q * 4 + t];
        const double akn = reinterpret_cast<const double*>(yb)[k];   // [row k][col k+1]
        const double ann = reinterpret_cast<const double*>(yb)[kn];  // [row k+1][col k+1]
        const double akc = __shfl_sync(FULL, (k & 1) ? pa[0].y : pa[0].x, 4 * g + (k >> 1));  // [row k][col g]
        const double xkn = __shfl_sync(FULL, (kn & 1) ? x[0].y : x[0].x, 4 * g + (kn >> 1));  // [row k+1][col k]
        const double n2n = __shfl_sync(FULL, n2, 4 * kn);
        double2 xm = x[0];
        if (2 * t <= k) xm.x = 0.0;
        if (2 * t + 1 <= k) xm.y = 0.0;
        double d0 = xm.x * pa[0].x, d1 = xm.y * pa[0].y;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            d0 = fma(x[q].x, pa[q].x, d0);
            d1 = fma(x[q].y, pa[q].y, d1);
        }
        const double dsum = d0 + d1;
        const bool nz = __any_sync(FULL, g == k && dsum != 0.0);     // |x|^2 != 0 (a sum of squares in group k)
        const bool low = __any_sync(FULL, g == k && !(n2 >= n2ref));  // the downdated norm lost too many digits
        const double dot = red4_mma(dsum);
        const double dn = __shfl_sync(FULL, dot, 4 * kn);
        if (low) {
            const double ss = __shfl_sync(FULL, dot, 4 * k);
            R = reflector(fma(alpha, alpha, ss), alpha);
        }
        // dlarfg: tau = 0, H = I when the column below the pivot is exactly zero
        const double ts = nz ? R.ts : 0.0, tau = nz ? R.tau : 0.0, beta = nz ? R.beta : alpha, scal = nz ? R.scal : 1.0;
        const double amb = alpha - beta;
        const double zz = fma(scal, dot, akc);  // v_k . a_c (c > k);  scal_i zz = v_i . v_k (i < k)
        if (g < k && t == 0) tq[g * 8 + k] = nz ? scalv * zz : 0.0;
        const bool piv = g == k;
        const double w = g > k ? ts * zz : 0.0;
        // the next reflector, in every lane: w, R[k][k+1], downdated norm and pivot element of column k + 1
        const double wn = ts * fma(scal, dn, akn);
        const double rkn = fma(-wn, amb, akn);
        const double x2n = fma(-rkn, rkn, n2n);
        const double alphan = fma(-wn, xkn, ann);
        {
            const double rkc = fma(-w, amb, akc);  // R[k][g]
            n2 = fma(-rkc, rkc, n2);
        }
        double2 ud = xm;  // u on my rows of the diagonal block
        if (2 * t == k) ud.x = amb;
        if (2 * t + 1 == k) ud.y = amb;
        {
            const double nx = fma(-w, ud.x, pa[0].x), ny = fma(-w, ud.y, pa[0].y);
            pa[0].x = (piv && 2 * t == k) ? beta : nx;
            pa[0].y = (piv && 2 * t + 1 == k) ? beta : ny;
        }
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
            pa[q].x = fma(-w, x[q].x, pa[q].x);
            pa[q].y = fma(-w, x[q].y, pa[q].y);
        }
        // the column after the next one starts its trip (its owner is up to date through reflector k)
        if (g == ((k + 2) & 7)) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) xb2[(k & 1) * XS + q * 4 + t] = pa[q];
        }
        // x <- column k + 1 after reflector k
        {
            const double2 x0 = make_double2(fma(-wn, ud.x, y[0].x), fma(-wn, ud.y, y[0].y));
#pragma unroll
            for (int q = 1; q < NQ; ++q) {
                x[q].x = fma(-wn, x[q].x, y[q].x);
                x[q].y = fma(-wn, x[q].y, y[q].y);
            }
            x[0] = x0;
        }
        if (piv) {
            tauv = tau;
            scalv = scal;
        }
        alpha = alphan;
        R = reflector(x2n, alphan);
        __syncwarp();
    }
    const double2 r = pa[0];
    Rd = make_double2(2 * t <= g ? r.x : 0.0, 2 * t + 1 <= g ? r.y : 0.0);
    pa[0].x = 2 * t > g ? r.x * scalv : (2 * t == g ? 1.0 : 0.0);
    pa[0].y = 2 * t + 1 > g ? r.y * scalv : (2 * t + 1 == g ? 1.0 : 0.0);
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
        pa[q].x *= scalv;
        pa[q].y *= scalv;
    }
}

// Compact-WY factor (dlarft) from the Gram matrix of the reflectors:  T^{-1} = striu(V^T V) + diag(1 / tau), so column j
// of T solves  x_i = tau_i (delta_ij - sum_{l > i} G_il x_l)  by back substitution -- no division, a reflector with
// tau = 0 gives a zero row and column.  Group j solves column j with broadcast reads of G; the sums are kept per row
// and advanced as soon as an x_l is known (two dependent operations per level).  Returns T^T in C layout: lane
// (g = j, t) holds T[2t][j], T[2t+1][j].  tq: G (64, written by panel_reflectors) + tau (8).
__device__ __forceinline__ double2 t_factor(double* __restrict__ tq, const double tauv, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    if (t == 0) tq[64 + g] = tauv;
    __syncwarp();
    double x[8], sacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc[i] = 0.0;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        const double ti = tq[64 + i];
        x[i] = i == g ? ti : (i < g ? -ti * sacc[i] : 0.0);
#pragma unroll
        for (int l = 0; l < i; ++l) sacc[l] = fma(tq[l * 8 + i], x[i], sacc[l]);
    }
    double2 tt;
    tt.x = t == 0 ? x[0] : (t == 1 ? x[2] : (t == 2 ? x[4] : x[6]));
    tt.y = t == 0 ? x[1] : (t == 1 ? x[3] : (t == 2 ? x[5] : x[7]));
    __syncwarp();
    return tt;
}

// (A^T V) block of one column block: independent accumulator chains (even / odd q, first / second k-half)
template <int NQ>
__device__ __forceinline__ double2 block_w(const double2 (&c)[NQ], const double2 (&pa)[NQ]) {
    double2 wa = zero2(), wb = zero2(), wc = zero2(), wd = zero2();
#pragma unroll
    for (int q = 0; q < NQ; q += 2) {
        dmma(wa.x, wa.y, c[q].x, pa[q].x);
        if (q + 1 < NQ) dmma(wc.x, wc.y, c[q + 1 < NQ ? q + 1 : q].x, pa[q + 1 < NQ ? q + 1 : q].x);
        dmma(wb.x, wb.y, c[q].y, pa[q].y);
        if (q + 1 < NQ) dmma(wd.x, wd.y, c[q + 1 < NQ ? q + 1 : q].y, pa[q + 1 < NQ ? q + 1 : q].y);
    }
    return make_double2((wa.x + wb.x) + (wc.x + wd.x), (wa.y + wb.y) + (wc.y + wd.y));
}
// A^T -= ((A^T V) T) V^T
template <int NQ>
__device__ __forceinline__ void block_apply(double2 (&c)[NQ], const double2 wt, const double2 tt, const double2 (&vv)[NQ]) {
    double2 u0 = zero2(), u1 = zero2();  // (A^T V) T
    dmma(u0.x, u0.y, wt.x, tt.x);
    dmma(u1.x, u1.y, wt.y, tt.y);
    const double2 ut = make_double2(-(u0.x + u1.x), -(u0.y + u1.y));
    // first k-halves of every tile, then the second halves
#pragma unroll
    for (int q = 0; q < NQ; ++q) dmma(c[q].x, c[q].y, ut.x, vv[q].x);
#pragma unroll
    for (int q = 0; q < NQ; ++q) dmma(c[q].x, c[q].y, ut.y, vv[q].y);
}

// One panel (8 columns, NQ active row blocks, the diagonal block first) of the Householder sweep and its trailing update.
// tile(jt, rt) -> tile index.  The reflector loop is rolled (runtime k): see qr_tiles.
template <int NQ, int WT, class ColF>
__device__ __forceinline__ void qr_panel(double* __restrict__ S, double* __restrict__ tq, const int jt, const int np,
                                         const int row0, ColF colf, const int lane) {
    // tile (column block j, active row block q) = (row0 + j) * WT + col[q]: the column part is fixed for the panel
    int col[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) col[q] = colf(jt + q);
    auto tile = [&](int j, int rt) { return (row0 + j) * WT + col[rt - jt]; };
    double2 pa[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) pa[q] = ldt(S, tile(jt, jt + q), lane);
    double tauv;
    double2 Rd;
    {
        BKF_PTICK();
        BKF_PANEL<NQ>(pa, Rd, tauv, tq, tq + 72, lane);
        BKF_PTOCK(0);
    }
    BKF_PTICK();
    stt(S, tile(jt, jt), lane, Rd);
#pragma unroll
    for (int q = 1; q < NQ; ++q) stt(S, tile(jt, jt + q), lane, zero2());
    if (jt + 1 < np) {
        // the first column block's A^T V does not need T: its DMMA chains run beside the back substitution
        double2 c[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) c[q] = ldt(S, tile(jt + 1, jt + q), lane);
        double2 wt = block_w<NQ>(c, pa);
        const double2 tt = t_factor(tq, tauv, lane);
        double2 vv[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) vv[q] = ttrans(pa[q], lane);
        BKF_PTOCK(1);
#pragma unroll 1
        for (int j2 = jt + 1;;) {
            block_apply<NQ>(c, wt, tt, vv);
#pragma unroll
            for (int q = 0; q < NQ; ++q) stt(S, tile(j2, jt + q), lane, c[q]);
            if (++j2 >= np) break;
#pragma unroll
            for (int q = 0; q < NQ; ++q) c[q] = ldt(S, tile(j2, jt + q), lane);
            wt = block_w<NQ>(c, pa);
        }
    }
    __syncwarp();
    BKF_PTOCK(2);
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
    static_assert(NRMAX <= 6, "the travelling-column buffer of panel_reflectors_la holds six row blocks (Dims::FWD_ELEMS)");
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
                         V_in = V_sv + p, V_w = V_in + p, V_ym = V_w + p, V_tq = V_ym + p;  // V_tq: Gram tile + taus (72), travelling columns (2 x 48)
    static constexpr int FWD_ELEMS = V_tq + 72 + 96;
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

// TZS: T and Z are kept in shared memory as tiles (12 KB more per series at the C4 shape).  Chosen when every series has a
// scheduler to itself anyway (B <= 4 series per SM): the per-step fetch of T from L2 then leaves the critical path;
// larger batches keep the small arena (8 series per SM) and read T / Z through L1.
template <int MT, int PT, int RT, bool TZS>
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
    double* TTs = smem + D::FWD_ELEMS;          // TZS only: T tiles (I, K) at I MT + K, then Z tiles (I, J) at I MT + J
    double* ZZs = TTs + MT * MT * 64;
    if (TZS) {
        for (int idx = 0; idx < MT * MT; ++idx) stt(TTs, idx, lane, ldg_tile(Tg, m, idx / MT, idx % MT, lane));
        for (int idx = 0; idx < PT * MT; ++idx) stt(ZZs, idx, lane, ldg_tile(Zg, m, idx / MT, idx % MT, lane));
        __syncwarp();
    }
    auto t_tile = [&](int I, int K) { return TZS ? ldt(TTs, I * MT + K, lane) : ldg_tile(Tg, m, I, K, lane); };
    auto z_tile = [&](int I, int J) { return TZS ? ldt(ZZs, I * MT + J, lane) : ldg_tile(Zg, m, I, J, lane); };
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
    // record for the adjoint (bkf_sqrt_bwd4.cuh): [a | a_f | v | s = Fc^{-1} v | inner = Fc^{-1} s | tiles of Rf^T | tiles of
    // Rp^T], the tiles exactly as they lie in the arena (upper tile (I, K) at tri(I, K), 512-byte coalesced stores); Pc is
    // not stored (it is the previous step's Rp^T, or P0)
    constexpr int NTRI1 = D::NT1 * (D::NT1 + 1) / 2, NTRIP = MT * (MT + 1) / 2;
    constexpr size_t off_rf = (size_t)(2 * m + 3 * p), off_rp = off_rf + (size_t)NTRI1 * 64, tstride = off_rp + (size_t)NTRIP * 64;
    double logp = 0.0;
    for (int ts = 0; ts < n; ++ts) {
        BKF_PTICK();
#ifdef BKF_SQW_PROFILE
        long long _pm = clock64();
#endif
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
                double2 z = z_tile(I, J);
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
            BKF_PMARK(4);  // mask, innovation, Z tiles
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
            // Zm Pc.  From the second step on Pc is a predict factor -- lower triangular, its tiles K < J exact zeros -- and
            // the products with them are left out at compile time (one copy of the code per case, chosen per step).
            auto zpc = [&](auto TRI) {
                constexpr bool tri = decltype(TRI)::value;
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    double2 acc[PT];
#pragma unroll
                    for (int I = 0; I < PT; ++I) acc[I] = zero2();
#pragma unroll
                    for (int K = tri ? J : 0; K < MT; ++K) {
                        const double2 pt = ttrans(ldt(SA, D::pc_tile(K, J), lane), lane);  // (Pc[K][J])^T
#pragma unroll
                        for (int I = 0; I < PT; ++I) mma_nt(acc[I], zf[I][K], pt);
                    }
#pragma unroll
                    for (int I = 0; I < PT; ++I) stt(SA, I * WT + PT + J, lane, acc[I]);
                }
            };
            if (ts > 0) zpc(TrueT{});
            else zpc(FalseT{});
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int J = 0; J < PT; ++J) stt(SA, (PT + I) * WT + J, lane, zero2());
            __syncwarp();
            BKF_PMARK(5);  // update array set-up (Zm Pc)
            qr_tiles<D::NT1, MT, PT, WT>(SA, smem + D::V_tq, D::NT1, D::NT1, false, lane);
            BKF_PMARK(6);  // update sweep
            if (rec) {
                int idx = 0;
#pragma unroll
                for (int I = 0; I < D::NT1; ++I)
#pragma unroll
                    for (int K = I; K < D::NT1; ++K, ++idx)
                        reinterpret_cast<double2*>(rec + off_rf)[idx * 32 + lane] = ldt(SA, K * WT + I, lane);
            }
            if (G.obs_cov)  // F = Fc Fc^T, Fc = B[:p,:p]
                store_square<PT>(G.obs_cov + bt * p * p, [&](int I, int K) { return ldt(SA, I * WT + K, lane); }, lane);
            BKF_PMARK(7);  // record of Rf, obs_cov
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
            BKF_PMARK(8);  // triangular solves, ll
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
        if (rec) {
            for (int i = lane; i < m; i += 32) rec[m + i] = vaf[i];
            if (lane < p) {
                rec[2 * m + lane] = vv[lane];
                rec[2 * m + p + lane] = vsv[lane];
                rec[2 * m + 2 * p + lane] = vin[lane];
            }
        }
        if (G.filt_a)
            for (int i = lane; i < m; i += 32) G.filt_a[bt * m + i] = vaf[i];
        if (G.filt_P)
            store_square<MT>(G.filt_P + bt * m * m, [&](int I, int K) { return ldt(SA, D::pc_tile(I, K), lane); }, lane);
        if (G.ll && lane == 0) G.ll[bt] = ll;
        logp += ll;
        // ---- predict array (transposed) [T Pcf, R Qc]; a+ = T a_f + c.  Pcf is a QR factor (tiles K < J exact zeros) unless
        //      it still is the caller's P0 (first step with every observation missing) ----
        BKF_PMARK(9);  // a_f, jitter, vector record, outputs
        auto tpcf = [&](auto TRI) {
            constexpr bool tri = decltype(TRI)::value;
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
                for (int J = 0; J < (tri ? K + 1 : MT); ++J) pt[J] = ttrans(ldt(SA, D::pc_tile(K, J), lane), lane);
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    const double2 tf = t_tile(I, K);
                    an[I] = fma(tf.x, vaf[8 * K + 2 * t], an[I]);
                    an[I] = fma(tf.y, vaf[8 * K + 2 * t + 1], an[I]);
#pragma unroll
                    for (int J = 0; J < (tri ? K + 1 : MT); ++J) mma_nt(tp[I][J], tf, pt[J]);
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
        };
        if (ts > 0 || !flag) tpcf(TrueT{});
        else tpcf(FalseT{});
        BKF_PMARK(10);  // predict array (T Pcf), a+
        qr_tiles<D::NT1, MT, PT, WT>(SA, smem + D::V_tq, MT, MT + RT, true, lane);
        BKF_PMARK(11);  // predict sweep
        if (rec) {
            int idx = 0;
#pragma unroll
            for (int I = 0; I < MT; ++I)
#pragma unroll
                for (int K = I; K < MT; ++K, ++idx)
                    reinterpret_cast<double2*>(rec + off_rp)[idx * 32 + lane] = ldt(SA, D::pc_tile(K, I), lane);
        }
        BKF_PTOCK(3);
    }
#ifdef BKF_SQW_PROFILE
    if (b == 0 && lane == 0)
        printf("sqw profile (cycles per step): reflectors %llu  T-factor+store %llu  (with trailing %llu)  step %llu\n",
               g_prof[0] / n, g_prof[1] / n, g_prof[2] / n, g_prof[3] / n);
    if (b == 0 && lane == 0) {
        printf("   phases: innovation %llu  update set-up %llu  update sweep %llu  Rf record %llu  solves %llu  a_f/record %llu  predict set-up %llu  "
               "predict sweep %llu\n", g_prof[4] / n, g_prof[5] / n, g_prof[6] / n, g_prof[7] / n, g_prof[8] / n, g_prof[9] / n,
               g_prof[10] / n, g_prof[11] / n);
        for (int i = 0; i < 16; ++i) g_prof[i] = 0;
    }
#endif
    if (lane == 0) {
        if (!isfinite(logp)) st |= BKF_STATUS_NONFINITE;
        if (G.logp) G.logp[b] = logp;
        if (G.status) G.status[b] = st;
    }
}

// ===============================================================================================================
// Helpers of the adjoint sweep (bkf_sqrt_bwd4.cuh).

template <int NT>
__host__ __device__ constexpr int tri(int I, int K) { return I * NT - (I * (I - 1)) / 2 + (K - I); }  // I <= K

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

}  // namespace sqw
}  // namespace bkf
