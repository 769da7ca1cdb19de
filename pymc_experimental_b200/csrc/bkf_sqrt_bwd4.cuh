// bkf_sqrt_bwd4.cuh -- adjoint sweep of the SquareRootFilter (kalman_filter.py:620-750), FOUR WARPS PER SERIES.
//
// Same mathematics as the one-warp adjoint it replaces (Q-free QR pull-back  Abar = A Psi,
// Psi = R^{-1} copyltu(R Rbar^T) R^{-T}, the R factors read back from the forward sweep's record), reorganised so that
// the critical path of a step is short -- a few hundred series per GPU means the sweep runs at the latency of ONE
// series' dependency chain, not at the throughput of the FP64 pipe:
//
//   * every tile phase is split over the four warps of the CTA by output tiles with statically balanced costs; a phase
//     ends with a CTA barrier (12 per step);
//   * the two right-solves of Psi run RIGHT-LOOKING on whole tile rows held in registers (a row of X R^{-T} depends on
//     that row of X only): per tile column one product with the inverted diagonal tile and one update, everything else
//     off the chain; six rows over four warps = two interleaved rows on two of them;
//   * Psi is symmetric: the second pass computes its lower tiles only and mirrors them;
//   * the inverses of the diagonal tiles come from a column-oriented back substitution (8 dependent steps instead of
//     28) with hardware-seeded reciprocals, the four triangular vector solves of the step are products with the
//     explicit inverse of the p x p factor block (its tiles are those inverses); the vector chain of the step runs on
//     one warp beside tile work of the other three that does not depend on it;
//   * the factors of the step arrive by cp.async (no registers held across the L2 latency) one phase before they are
//     needed; T^T stays in shared memory for the whole sweep;
//   * the record holds the factors as 8 x 8 tiles exactly as the forward sweep has them in shared memory (R^T tiles in
//     DMMA accumulator order, 512-byte coalesced accesses): [a | tiles of R_update^T | tiles of R_predict^T].  Pc is not
//     stored (it is the previous step's R_predict^T, or P0).  The adjoint needs both orientations of a tile: the natural
//     one goes to shared memory through one in-register transposition at load time, the transposed one is the record
//     itself (L2);
//   * sum_t Hc^T Psi11 = Hc^T sum_t Psi11: the per-step product is gone, sum Psi11 joins the add-only accumulators in L2
//     (T-bar, Z-bar, sum Psi_p), all updated with fire-and-forget RED.ADDs.
//
// Shapes: as bkf_sqrt_warp.cuh (float64; m, p, r multiples of 8; r <= p), p <= 16.
#pragma once
#include "bkf_sqrt_warp.cuh"

namespace bkf {
namespace sqb {

using namespace sqw;

constexpr int NWARP = 4;
constexpr int NTHR = 32 * NWARP;

__device__ __forceinline__ double2 ldg_rec(const double* __restrict__ rec, int tile, int lane) {
    return reinterpret_cast<const double2*>(rec)[tile * 32 + lane];
}
// owner of lower tile (I, J), J <= I, of an NT x NT grid whose tile costs fall with I (NT - I products): the tiles in
// row-major order dealt out boustrophedon -- exactly balanced for NT = 6 (14 products each), 6 / 5 / 5 / 4 for NT = 4
__host__ __device__ constexpr int snake_owner(int I, int J) {
    const int s = I * (I + 1) / 2 + J;
    return ((s >> 2) & 1) ? 3 - (s & 3) : (s & 3);
}
__device__ __forceinline__ double2 add_diag(double2 v, double x, int lane) {
    const int g = lane >> 2, t = lane & 3;
    if (2 * t == g) v.x += x;
    if (2 * t + 1 == g) v.y += x;
    return v;
}
__device__ __forceinline__ double2 upper_only(double2 v, int lane) {  // zero the strictly lower part of a diagonal tile
    const int g = lane >> 2, t = lane & 3;
    if (2 * t < g) v.x = 0.0;
    if (2 * t + 1 < g) v.y = 0.0;
    return v;
}

template <int MT, int PT, int RT>
struct BD {
    static constexpr int m = 8 * MT, p = 8 * PT, r = 8 * RT, NT1 = MT + PT;
    static_assert(PT <= 2, "the explicit inverse of the p x p factor block is written for p <= 16");
    static constexpr int NTRI1 = NT1 * (NT1 + 1) / 2, NTRIP = MT * (MT + 1) / 2;
    // record (doubles)
    static constexpr int REC_AF = m, REC_V = 2 * m, REC_S = 2 * m + p, REC_IN = 2 * m + 2 * p, REC_RF = 2 * m + 3 * p,
                         REC_RP = REC_RF + NTRI1 * 64, REC = REC_RP + NTRIP * 64;
    // per-series accumulators in global memory (doubles)
    static constexpr int SCR_GT = 0, SCR_GZ = m * m, SCR_PS = m * m + p * m, SCR_HB = 2 * m * m + p * m,
                         SCR = 2 * m * m + p * m + p * p;
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    // shared memory (doubles)
    static constexpr int WTILES = cmax(NT1 * NT1, 2 * MT * MT + MT);  // Psi | Psi_p, TP^T / U1^T, inverses (predict)
    static constexpr int RFT = cmax(NTRI1, MT * PT);                  // update factor (later: Abar21)
    static constexpr int W = 0, RF = W + WTILES * 64, RP = RF + RFT * 64, PBT = RP + NTRIP * 64,
                         TT = PBT + MT * MT * 64, RINV = TT + MT * MT * 64, U01 = RINV + NT1 * 64, VA = U01 + 64;
    static constexpr int V_a = VA, V_af = V_a + m, V_ab = V_af + m, V_afb = V_ab + m, V_gc = V_afb + m, V_v = V_gc + m,
                         V_s = V_v + p, V_in = V_s + p, V_w = V_in + p, V_ym = V_w + p, V_vb = V_ym + p, V_sb = V_vb + p,
                         V_t1 = V_sb + p, V_t2 = V_t1 + p, V_gd = V_t2 + p, V_d = V_gd + p, V_x = V_d + p, SI = V_x + p;
    static constexpr int ELEMS = SI + 2;  // 4 ints: flag, partial, any partial
    static_assert(ELEMS >= cmax(5 * p * p, 6 * r * r + 2 * m * r) + p * p, "epilogue scratch");
};

__device__ __forceinline__ double fast_rcp(double x) {  // 1 / x from the hardware seed and two Newton steps
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    r0 = fma(r0, fma(-x, r0, 1.0), r0);
    return fma(r0, fma(-x, r0, 1.0), r0);
}
// inverse of the upper-triangular 8 x 8 tile D (row-major = C layout) into out, one column per lane c = 0..7
__device__ __forceinline__ void inv_upper_tile(const double* __restrict__ Dg, double* __restrict__ out, const int c) {
    double s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.0;
#pragma unroll
    for (int q = 7; q >= 0; --q) {
        const double x = ((q == c ? 1.0 : 0.0) - s[q]) * fast_rcp(Dg[q * 9]);
        out[q * 8 + c] = x;
#pragma unroll
        for (int i = 0; i < q; ++i) s[i] = fma(Dg[i * 8 + q], x, s[i]);
    }
}
// 16 bytes global -> shared without a register in between
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// tiles idx = warp, warp + 4, ... of a record block: fetch as they are (transposed content) / turn in place afterwards
__device__ __forceinline__ void fetch_tiles(double* __restrict__ dst, const double* __restrict__ src, const int ntiles,
                                            const int warp, const int lane) {
    for (int idx = warp; idx < ntiles; idx += NWARP)
        cp_async16(reinterpret_cast<double2*>(dst) + idx * 32 + lane, reinterpret_cast<const double2*>(src) + idx * 32 + lane);
}
__device__ __forceinline__ void turn_tiles(double* __restrict__ dst, const int ntiles, const int warp, const int lane) {
    for (int idx = warp; idx < ntiles; idx += NWARP) stt(dst, idx, lane, ttrans(ldt(dst, idx, lane), lane));
}

// One right-solve pass  X <- X R^{-T}  on NR tile rows of X held in registers (a row of the result depends on that row
// of X only); RF: upper tiles of R (tri<NT> order), RINV: inverses of its diagonal tiles.
//   COL = false: X = rows r0, r1 of W, all tile columns, result in place: G = S R^{-T}.
//   COL = true:  X = the transposes of COLUMNS r0, r1 of W from the diagonal down (tile rows I >= r); the result tile
//                Y[I] = Psi[r][I] goes to (r, I) as it is and, transposed, to (I, r): applied to G this leaves the
//                symmetric Psi = R^{-1} G = (G^T R^{-T})^T in W, computed for its lower tiles only.
// Either way a warp reads and writes tiles nobody else touches in that pass (column r below the diagonal and row r
// right of it): no barrier inside a pass.  NR = 2 with r1 < 0: one row.
template <int NT, int NR, bool COL>
__device__ __forceinline__ void solve_rows(double* __restrict__ W, const double* __restrict__ RF,
                                           const double* __restrict__ RINV, const int r0, const int r1, const int lane) {
    double2 X[NR][NT];
    int rw[NR];
    rw[0] = r0;
    if (NR > 1) rw[NR - 1] = r1;
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int I = 0; I < NT; ++I) {
            if (COL) X[rr][I] = I >= rw[rr] ? ttrans(ldt(W, I * NT + rw[rr], lane), lane) : zero2();
            else X[rr][I] = ldt(W, rw[rr] * NT + I, lane);
        }
#pragma unroll
    for (int I = NT - 1; I >= 0; --I) {
        if (COL && I < rw[0] && (NR == 1 || I < rw[NR - 1])) break;
        const double2 ri = ldt(RINV, I, lane);
        double2 ny[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
            if (COL && I < rw[rr]) continue;
            // two independent half-products instead of a dependent pair: this product is on the chain
            double2 ya = zero2(), yb = zero2();
            dmma(ya.x, ya.y, X[rr][I].x, ri.x);
            dmma(yb.x, yb.y, X[rr][I].y, ri.y);
            X[rr][I] = make_double2(ya.x + yb.x, ya.y + yb.y);
            ny[rr] = neg(X[rr][I]);
        }
#pragma unroll
        for (int K = I - 1; K >= 0; --K) {  // the next column first: it is the one on the chain
            if (COL && K < rw[0] && (NR == 1 || K < rw[NR - 1])) break;
            const double2 rk = ldt(RF, tri<NT>(K, I), lane);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
                if (!COL || K >= rw[rr]) mma_nt(X[rr][K], ny[rr], rk);
        }
    }
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int I = 0; I < NT; ++I) {
            if (COL) {
                if (I >= rw[rr]) {
                    stt(W, I * NT + rw[rr], lane, ttrans(X[rr][I], lane));
                    if (I > rw[rr]) stt(W, rw[rr] * NT + I, lane, X[rr][I]);
                }
            } else {
                stt(W, rw[rr] * NT + I, lane, X[rr][I]);
            }
        }
}
// rows of a pass by warp.  Full rows (first pass): w and w + 4.  Lower-part columns (second pass; column r costs
// (NT - r)(NT - r + 1) / 2 products): w and, for w >= 1, NT - w when that is a row beyond the first four.
template <int NT, bool COL>
__device__ __forceinline__ void solve_pass(double* __restrict__ W, const double* __restrict__ RF,
                                           const double* __restrict__ RINV, const int warp, const int lane) {
    if (warp >= NT) return;
    if (NT >= 5) {
        const int r1 = COL ? ((warp >= 1 && NT - warp >= 4) ? NT - warp : -1) : (warp + 4 < NT ? warp + 4 : -1);
        if (r1 >= 0) { solve_rows<NT, NT >= 5 ? 2 : 1, COL>(W, RF, RINV, warp, r1, lane); return; }
    }
    solve_rows<NT, 1, COL>(W, RF, RINV, warp, -1, lane);
}

// inverses of the diagonal tiles of the upper factor RF (NT x NT tiles): tile q by lanes 0..7 of warp q & 3 (lanes
// 8..15 take tile q + 4)
template <int NT>
__device__ __forceinline__ void inv_diag_tiles(const double* __restrict__ RF, double* __restrict__ RINV, const int warp,
                                               const int lane) {
    if (lane < 16) {
        const int q = warp + 4 * (lane >> 3);
        if (q < NT) inv_upper_tile(RF + tri<NT>(q, q) * 64, RINV + q * 64, lane & 7);
    }
}

// Lower tiles (I, J), J <= I, of S = copyltu(R Rbar^T) into W (NT x NT tiles), mirrored.  rbar(J, K), K >= J: cotangent
// tiles (upper, exact zeros below the diagonal).  The tiles are dealt out by cost (tile (I, J) costs NT - I products):
//   SET 0: all of them over the four warps;
//   SET 1: tile columns J >= J0 over warps 0 .. 2 (update adjoint: they do not depend on the vector chain of the step);
//   SET 2: tile columns J <  J0 over the four warps.
__host__ __device__ constexpr int snake(int s, int nw) { return ((s / nw) & 1) ? nw - 1 - (s % nw) : (s % nw); }
template <int NT, int SET, int J0, class RbF>
__device__ __forceinline__ void s0_tiles(double* __restrict__ W, const double* __restrict__ RF, RbF rbar, const int warp,
                                         const int lane) {
    const int g = lane >> 2, t = lane & 3;
    int seq = 0;
#pragma unroll
    for (int I = 0; I < NT; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            if (SET == 1 && J < J0) continue;
            if (SET == 2 && J >= J0) continue;
            const int owner = snake(seq++, SET == 1 ? 3 : 4);
            if (owner != warp) continue;
            double2 aa = zero2(), ab = zero2();
#pragma unroll
            for (int K = I; K < NT; ++K) {
                const double2 rk = ldt(RF, tri<NT>(I, K), lane), rb = rbar(J, K);
                if ((K - I) & 1) mma_nt(ab, rk, rb);
                else mma_nt(aa, rk, rb);
            }
            const double2 v = make_double2(aa.x + ab.x, aa.y + ab.y);
            if (J < I) {
                stt(W, I * NT + J, lane, v);
                stt(W, J * NT + I, lane, ttrans(v, lane));
            } else {
                const double2 vt = ttrans(v, lane);
                stt(W, I * NT + I, lane, make_double2(g >= 2 * t ? v.x : vt.x, g >= 2 * t + 1 ? v.y : vt.y));
            }
        }
}

template <int MT, int PT, int RT>
__global__ void __launch_bounds__(NTHR, 4) backward_kernel(KArgs<double> G) {
    using D = BD<MT, PT, RT>;
    constexpr int m = D::m, p = D::p, r = D::r, NT1 = D::NT1;
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    const int n = G.n;
    double *W = smem + D::W, *RF = smem + D::RF, *RP = smem + D::RP, *PBT = smem + D::PBT, *TTs = smem + D::TT,
           *RINV = smem + D::RINV, *UO = smem + D::U01;
    double *va = smem + D::V_a, *vaf = smem + D::V_af, *vab = smem + D::V_ab, *vafb = smem + D::V_afb,
           *vgc = smem + D::V_gc, *vv = smem + D::V_v, *vs = smem + D::V_s, *vin = smem + D::V_in, *vw = smem + D::V_w,
           *vvb = smem + D::V_vb, *vsb = smem + D::V_sb, *vt1 = smem + D::V_t1,
           *vt2 = smem + D::V_t2, *vgd = smem + D::V_gd, *vd = smem + D::V_d, *vx = smem + D::V_x;
    volatile int* si = reinterpret_cast<volatile int*>(smem + D::SI);

    const double* Zg = param_ptr(G.Z, G.shared_mask, BKF_PARAM_Z, b, (size_t)p * m);
    const double* P0g = param_ptr(G.P0, G.shared_mask, BKF_PARAM_P0, b, (size_t)m * m);
    const double* yb = G.y_shared ? G.y : G.y + (size_t)b * n * p;
    const double* rec0 = G.traj + (size_t)b * n * D::REC;  // record of step 0 of this series
    // add-only accumulators in global memory (L2): the gradient buffers themselves when requested, else scratch
    double* scr = G.traj + (size_t)G.B * n * D::REC + (size_t)b * D::SCR;
    double* gT = G.g_T ? G.g_T + (size_t)b * m * m : scr + D::SCR_GT;
    double* gZ = G.g_Z ? G.g_Z + (size_t)b * p * m : scr + D::SCR_GZ;
    double* PS = scr + D::SCR_PS;
    double* HB = scr + D::SCR_HB;  // sum of Psi11
    const bool want_ps = G.g_R || G.g_Q;

    {
        const double* dg = param_ptr(G.d, G.shared_mask, BKF_PARAM_D, b, (size_t)p);
        const double* Tg = param_ptr(G.T, G.shared_mask, BKF_PARAM_T, b, (size_t)m * m);
        for (int i = threadIdx.x; i < p; i += NTHR) { vd[i] = dg[i]; vgd[i] = 0.0; }
        for (int i = threadIdx.x; i < m; i += NTHR) { vab[i] = 0.0; vgc[i] = 0.0; }
        for (int i = threadIdx.x; i < MT * MT * 32; i += NTHR) reinterpret_cast<double2*>(PBT)[i] = zero2();
        for (int i = threadIdx.x; i < m * m; i += NTHR) { gT[i] = 0.0; PS[i] = 0.0; }
        for (int i = threadIdx.x; i < p * m; i += NTHR) gZ[i] = 0.0;
        for (int i = threadIdx.x; i < p * p; i += NTHR) HB[i] = 0.0;
        for (int idx = warp; idx < MT * MT; idx += NWARP) stt(TTs, idx, lane, ldg_tile_t(Tg, m, idx / MT, idx % MT, lane));
        if (threadIdx.x == 0) si[2] = 0;
        if (n > 0) {  // factors of the last step
            const double* rl = rec0 + (size_t)(n - 1) * D::REC;
            for (int idx = warp; idx < D::NTRIP; idx += NWARP) stt(RP, idx, lane, ttrans(ldg_rec(rl + D::REC_RP, idx, lane), lane));
            fetch_tiles(RF, rl + D::REC_RF, D::NTRI1, warp, lane);
        }
    }
    __syncthreads();

#ifdef BKF_SQC_PROFILE
    long long tb[16], tcb = clock64();
    for (int i = 0; i < 16; ++i) tb[i] = 0;
#define BKF_TOCK(i) { const long long tc1 = clock64(); tb[i] += tc1 - tcb; tcb = tc1; }
#else
#define BKF_TOCK(i)
#endif
    for (int ts = n - 1; ts >= 0; --ts) {
        const size_t bt = (size_t)b * n + ts;
        const double* rec = rec0 + (size_t)ts * D::REC;
        const double gcot = G.ll_cot ? G.ll_cot[bt] : 1.0;
        const bool dense_pc = ts == 0;  // P0 is whatever the caller passed; later factors are lower triangular
        if (ts > 1) {  // what the next step will fetch: pull it towards L2 while this step computes
            const char* nxt = reinterpret_cast<const char*>(rec - D::REC);
            for (int off = threadIdx.x * 128; off < (int)(D::REC_RP * sizeof(double)); off += NTHR * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
            const char* nx2 = reinterpret_cast<const char*>(rec - 2 * D::REC + D::REC_RP);
            for (int off = threadIdx.x * 128; off < (int)(D::NTRIP * 64 * sizeof(double)); off += NTHR * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx2 + off));
        }
        // ================= predict adjoint =================
        // RP: R_predict of this step (natural orientation, from the previous iteration).  Psi_p -> W[0 .. MT*MT), the
        // inverses of the diagonal tiles of R_predict behind the second block of W.
        double* RINVP = W + 2 * MT * MT * 64;
        inv_diag_tiles<MT>(RP, RINVP, warp, lane);
        s0_tiles<MT, 0, 0>(W, RP, [&](int J, int K) {
            const double2 v = ldt(PBT, J * MT + K, lane);
            return J == K ? upper_only(v, lane) : v;
        }, warp, lane);
        if (warp == NWARP - 1) {  // mask; a, a_f, v, s = Fc^{-1} v, inner = Fc^{-1} s as the forward sweep left them
            bool miss = true;
            if (lane < p) {
                const double yv = yb[(size_t)ts * p + lane];
                miss = is_missing(yv, G.fill, true);
                vw[lane] = miss ? 0.0 : 1.0;
                vv[lane] = rec[D::REC_V + lane];
                vs[lane] = rec[D::REC_S + lane];
                vin[lane] = rec[D::REC_IN + lane];
            }
            const int nobs = __popc(__ballot_sync(FULL, !miss));
            if (lane == 0) {
                si[0] = nobs == 0 ? 1 : 0;
                si[1] = (nobs != 0 && nobs != p) ? 1 : 0;
                if (nobs != 0 && nobs != p) si[2] = 1;
            }
            for (int i = lane; i < m; i += 32) {
                va[i] = rec[i];
                vaf[i] = rec[D::REC_AF + i];
            }
        }
        cp_async_wait();  // this step's update factor has arrived (fetched during the previous step): natural orientation
        turn_tiles(RF, D::NTRI1, warp, lane);
        __syncthreads(); BKF_TOCK(0)
        const bool flag = si[0] != 0;
        solve_pass<MT, false>(W, RP, RINVP, warp, lane);
        if (!flag) inv_diag_tiles<NT1>(RF, RINV, warp, lane);  // inverses of the diagonal tiles of the update factor
        __syncthreads(); BKF_TOCK(1)
        solve_pass<MT, true>(W, RP, RINVP, warp, lane);
        // (Pc^T)[I][K], K >= I (all K at t = 0), natural orientation: from RP once the previous step's predict factor is
        // there (update adjoint); slow = from the record (a skipped update needs it before)
        auto pct_slow = [&](int I, int K) {
            if (dense_pc) return ldg_tile_t(P0g, m, I, K, lane);
            return K >= I ? ttrans(ldg_rec(rec - D::REC + D::REC_RP, tri<MT>(I, K >= I ? K : I), lane), lane) : zero2();
        };
        auto pct = [&](int I, int K) {
            if (dense_pc) return ldg_tile_t(P0g, m, I, K, lane);
            return K >= I ? ldt(RP, tri<MT>(I, K >= I ? K : I), lane) : zero2();
        };
        auto pcl = [&](int J, int K) {  // Pc[J][K], K <= J
            if (dense_pc) return ldg_tile(P0g, m, J, K, lane);
            return K <= J ? ldg_rec(rec - D::REC + D::REC_RP, tri<MT>(K <= J ? K : J, J), lane) : zero2();
        };
        // (Pcf + jitter I)^T tile (I, K) and (Pcf + jitter I) tile (J, K); Pcf = Pc when the update was skipped
        auto pft = [&](int I, int K) {
            double2 v;
            if (flag) v = pct_slow(I, K);
            else v = K >= I ? ldt(RF, tri<NT1>(PT + I, PT + (K >= I ? K : I)), lane) : zero2();
            return I == K ? add_diag(v, G.jitter, lane) : v;
        };
        auto pfl = [&](int J, int K) {
            double2 v;
            if (flag) v = pcl(J, K);
            else v = K <= J ? ldg_rec(rec + D::REC_RF, tri<NT1>(PT + (K <= J ? K : J), PT + J), lane) : zero2();
            return J == K ? add_diag(v, G.jitter, lane) : v;
        };
        const bool dense_pf = flag && dense_pc;  // (a skipped step keeps Pc, which is dense only at t = 0)
        constexpr int TPT = MT * MT;             // tile offset of the second block inside W
        // TP^T = (T Pcf)^T = Pcf^T T^T :  [I][J] = sum_K pft(I,K) (T[J][K])^T, tile column J by warp J
        if (warp < MT) {
            const int J = warp;
            double2 acc[MT];
#pragma unroll
            for (int I = 0; I < MT; ++I) acc[I] = zero2();
#pragma unroll
            for (int K = 0; K < MT; ++K) {
                const double2 tk = ttrans(ldt(TTs, K * MT + J, lane), lane);  // T[J][K] = ((T^T)[K][J])^T
#pragma unroll
                for (int I = 0; I < MT; ++I)
                    if (dense_pf || K >= I) mma_nt(acc[I], pft(I, K), tk);
            }
#pragma unroll
            for (int I = 0; I < MT; ++I) stt(W, TPT + I * MT + J, lane, acc[I]);
        }
        __syncthreads(); BKF_TOCK(2)
        // R_predict is dead: fetch the previous step's (= Pc^T of this step, the next iteration's R_predict)
        if (ts > 0) fetch_tiles(RP, rec - D::REC + D::REC_RP, D::NTRIP, warp, lane);
        if (want_ps && warp < MT) {
#pragma unroll
            for (int J = 0; J < MT; ++J) red_tile(PS, m, warp, J, lane, ldt(W, warp * MT + J, lane));
        }
        if (warp == NWARP - 1 && !flag && PT == 2) {  // U = (Rf[:p,:p])^{-1} = Fc^{-T}:  U01 = -U00 R01 U11
            double2 tmp = zero2(), u = zero2();
            mma_nt(tmp, ldt(RINV, 0, lane), ttrans(ldt(RF, tri<NT1>(0, 1), lane), lane));
            mma_nt(u, neg(tmp), ttrans(ldt(RINV, 1, lane), lane));
            stt(UO, 0, lane, u);
        }
        // U1^T = TP^T Psi_p, tile row I by warp I, in place
        if (warp < MT) {
            const int I = warp;
            double2 tp[MT];
#pragma unroll
            for (int K = 0; K < MT; ++K) tp[K] = ldt(W, TPT + I * MT + K, lane);
            double2 acc[MT];
#pragma unroll
            for (int J = 0; J < MT; ++J) acc[J] = zero2();
#pragma unroll
            for (int K = 0; K < MT; ++K)
#pragma unroll
                for (int J = 0; J < MT; ++J) mma_nt(acc[J], tp[K], ldt(W, J * MT + K, lane));
#pragma unroll
            for (int J = 0; J < MT; ++J) stt(W, TPT + I * MT + J, lane, acc[J]);
        }
        __syncthreads(); BKF_TOCK(3)
        if (warp < MT) {
            const int I = warp;
            // gT += U1 Pcf^T + abar af^T (tile row I);  U1[I][K] = (U1^T[K][I])^T
            if (G.g_T) {
                double2 u1[MT];
#pragma unroll
                for (int K = 0; K < MT; ++K) u1[K] = ttrans(ldt(W, TPT + K * MT + I, lane), lane);
                const double ab = vab[8 * I + g];
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    double2 acc = make_double2(ab * vaf[8 * J + 2 * t], ab * vaf[8 * J + 2 * t + 1]);
#pragma unroll
                    for (int K = 0; K < MT; ++K)
                        if (dense_pf || K <= J) mma_nt(acc, u1[K], pfl(J, K));
                    red_tile(gT, m, I, J, lane, acc);
                }
            }
            // Pcfbar^T = U1^T T :  [I][J] = sum_K U1T[I][K] ((T^T)[J][K])^T  -> PBT (the carried cotangent is dead)
            {
                double2 ut[MT], acc[MT];
#pragma unroll
                for (int K = 0; K < MT; ++K) ut[K] = ldt(W, TPT + I * MT + K, lane);
#pragma unroll
                for (int J = 0; J < MT; ++J) acc[J] = zero2();
#pragma unroll
                for (int K = 0; K < MT; ++K)
#pragma unroll
                    for (int J = 0; J < MT; ++J) mma_nt(acc[J], ut[K], ldt(TTs, J * MT + K, lane));
#pragma unroll
                for (int J = 0; J < MT; ++J) stt(PBT, I * MT + J, lane, acc[J]);
            }
            // afbar = T^T abar, tile column J = warp
            {
                const int J = warp;
                double afp = 0.0;
#pragma unroll
                for (int K = 0; K < MT; ++K) {
                    const double2 tk = ldt(TTs, J * MT + K, lane);  // [g][2t+e] = T[8K + 2t + e][8J + g]
                    afp = fma(tk.x, vab[8 * K + 2 * t], fma(tk.y, vab[8 * K + 2 * t + 1], afp));
                }
                afp = red_t(afp);
                if (t == 0) vafb[8 * J + g] = afp;
            }
        }
        __syncthreads(); BKF_TOCK(4)
        for (int i = threadIdx.x; i < m; i += NTHR) vgc[i] += vab[i];
        if (flag) {  // degenerate branch (:700-705): (a_f, Pcf, ll) = (a, Pc, 0)
            for (int i = threadIdx.x; i < m; i += NTHR) vab[i] = vafb[i];
            cp_async_wait();
            turn_tiles(RP, D::NTRIP, warp, lane);
            if (ts > 0) fetch_tiles(RF, rec - D::REC + D::REC_RF, D::NTRI1, warp, lane);
            __syncthreads(); BKF_TOCK(5)
            continue;
        }
        // ================= update adjoint =================
        // Rbar = Bbar^T (upper):  [Fcbar^T, KFbar^T; 0, triu(Pcfbar^T)], formed tile by tile where it is multiplied
        auto rbar = [&](int J, int K) {  // K >= J
            double2 v;
            if (J < PT) {
                const int i = 8 * J + g;
                const double sv = vs[i];
                if (K < PT) {
                    const int k = 8 * K + 2 * t;
                    const double ini = vin[i], dg = vx[i];
                    v.x = k >= i ? fma(-vt1[k], ini, -vt2[k] * sv) + (k == i ? dg : 0.0) : 0.0;
                    v.y = k + 1 >= i ? fma(-vt1[k + 1], ini, -vt2[k + 1] * sv) + (k + 1 == i ? dg : 0.0) : 0.0;
                } else {
                    v = make_double2(vafb[8 * (K - PT) + 2 * t] * sv, vafb[8 * (K - PT) + 2 * t + 1] * sv);
                }
            } else {
                v = ldt(PBT, (J - PT) * MT + (K - PT), lane);
                if (J == K) v = upper_only(v, lane);
            }
            return v;
        };
        if (warp == NWARP - 1) {  // the vector chain of the update adjoint
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
            auto utile = [&](int I, int J) { return I == J ? ldt(RINV, I, lane) : ldt(UO, 0, lane); };  // I <= J
            // y = U (alpha x + z) : y[8I + g] = sum_{J >= I} sum_{t,e} U[I][J][g][2t+e] (alpha x + z)[8J + 2t + e]
            auto u_mv = [&](const double alpha, const double* x, const double* z, double* y) {
#pragma unroll
                for (int I = 0; I < PT; ++I) {
                    double acc = 0.0;
#pragma unroll
                    for (int J = I; J < PT; ++J) {
                        const double2 u = utile(I, J);
                        const int k = 8 * J + 2 * t;
                        acc = fma(u.x, fma(alpha, x[k], z ? z[k] : 0.0), acc);
                        acc = fma(u.y, fma(alpha, x[k + 1], z ? z[k + 1] : 0.0), acc);
                    }
                    acc = red_t(acc);
                    if (t == 0) y[8 * I + g] = acc;
                }
                __syncwarp();
            };
            u_mv(lossbar, vv, nullptr, vt1);  // Fc^{-T} innerbar
            u_mv(1.0, vsb, vt1, vt2);         // Fc^{-T} (sbar + t1)
            if (lane < p) {
                vvb[lane] = fma(lossbar, vin[lane], vt2[lane]);
                vx[lane] = logdetbar * 2.0 * RINV[(lane >> 3) * 64 + (lane & 7) * 9];  // diagonal term of Fcbar: 1 / Fc_ii
            }
        } else {  // beside it: the tiles of S that do not depend on it (tile columns J >= PT of the lower triangle)
            s0_tiles<NT1, 1, PT>(W, RF, rbar, warp, lane);
        }
        cp_async_wait();  // Pc^T of this step has arrived
        turn_tiles(RP, D::NTRIP, warp, lane);
        __syncthreads(); BKF_TOCK(5)
        s0_tiles<NT1, 2, PT>(W, RF, rbar, warp, lane);
        __syncthreads(); BKF_TOCK(6)
        solve_pass<NT1, false>(W, RF, RINV, warp, lane);
        __syncthreads(); BKF_TOCK(7)
        solve_pass<NT1, true>(W, RF, RINV, warp, lane);
        __syncthreads(); BKF_TOCK(8)
        // ---- Abar = A Psi, A = [[Hc^T, 0], [A21, Pc^T]],  A21 = (Zm Pc)^T = Pc^T Zm^T -> first PT tiles of the rows of
        //      PBT (Pcfbar^T is dead; row I is read and overwritten by warp I alone below); sum of Psi11 for Hcbar ----
        double* AB21 = RF;  // the update factor is dead
#pragma unroll
        for (int I = 0; I < MT; ++I)
#pragma unroll
            for (int J = 0; J < PT; ++J) {
                // tile (I, J) costs MT - I products: rows I and MT - 1 - I pair up
                constexpr int half = (MT + 1) / 2;
                const int pairidx = (I < half ? I : MT - 1 - I) * PT + J;
                if ((pairidx & (NWARP - 1)) != warp) continue;
                double2 acc = zero2();
#pragma unroll
                for (int K = 0; K < MT; ++K)
                    if (dense_pc || K >= I) {
                        double2 z = ldg_tile(Zg, m, J, K, lane);
                        const double wJ = vw[8 * J + g];
                        z.x *= wJ;
                        z.y *= wJ;
                        mma_nt(acc, pct(I, K), z);
                    }
                stt(PBT, I * MT + J, lane, acc);
            }
        if (G.g_H) {
#pragma unroll
            for (int I = 0; I < PT; ++I)
#pragma unroll
                for (int J = 0; J < PT; ++J)
                    if (((I * PT + J) & (NWARP - 1)) == warp) red_tile(HB, p, I, J, lane, ldt(W, I * NT1 + J, lane));
        }
        __syncthreads(); BKF_TOCK(9)
        // rows of [Abar21, Abar22] (tile row I by warp I); Abar21 -> AB21, Pbar^T = Abar22 + Abar21 Zm -> PBT
        if (warp < MT) {
            const int I = warp;
            double2 acc[NT1];
#pragma unroll
            for (int J = 0; J < NT1; ++J) acc[J] = zero2();
#pragma unroll
            for (int K = 0; K < NT1; ++K)
                if (K < PT || dense_pc || K - PT >= I) {
                    const double2 ak = K < PT ? ldt(PBT, I * MT + (K < PT ? K : 0), lane) : pct(I, K >= PT ? K - PT : 0);
#pragma unroll
                    for (int J = 0; J < NT1; ++J) mma_nt(acc[J], ak, ldt(W, J * NT1 + K, lane));
                }
#pragma unroll
            for (int J = 0; J < PT; ++J) stt(AB21, I * PT + J, lane, acc[J]);
#pragma unroll
            for (int J = 0; J < MT; ++J) {
                double2 pb = acc[PT + J];
#pragma unroll
                for (int K = 0; K < PT; ++K) {
                    double2 zt = ldg_tile_t(Zg, m, J, K, lane);  // (Zm^T)[J][K]: [g][2t+e] = Zm[8K+2t+e][8J+g]
                    zt.x *= vw[8 * K + 2 * t];
                    zt.y *= vw[8 * K + 2 * t + 1];
                    mma_nt(pb, acc[K], zt);
                }
                stt(PBT, I * MT + J, lane, pb);
            }
        }
        __syncthreads(); BKF_TOCK(10)
        if (warp < MT) {
            const int J = warp;
            // gZ += W (ZPbar Pc^T - vbar a^T),  ZPbar = Abar21^T, tile column J
            if (G.g_Z) {
                double2 acc[PT];
#pragma unroll
                for (int Io = 0; Io < PT; ++Io) acc[Io] = zero2();
#pragma unroll
                for (int K = 0; K < MT; ++K)
                    if (dense_pc || K <= J) {
                        const double2 pc = pcl(J, K);
#pragma unroll
                        for (int Io = 0; Io < PT; ++Io) mma_nt(acc[Io], ttrans(ldt(AB21, K * PT + Io, lane), lane), pc);
                    }
#pragma unroll
                for (int Io = 0; Io < PT; ++Io) {
                    const double wo = vw[8 * Io + g], vb = vvb[8 * Io + g];
                    red_tile(gZ, m, Io, J, lane, make_double2(wo * fma(-vb, va[8 * J + 2 * t], acc[Io].x),
                                                              wo * fma(-vb, va[8 * J + 2 * t + 1], acc[Io].y)));
                }
            }
            // abar = afbar - Zm^T vbar (tile column J)
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
        if (warp == NWARP - 1 && lane < p) vgd[lane] -= vvb[lane];
        __syncthreads(); BKF_TOCK(11)
        // Abar21 is dead: fetch the next iteration's update factor
        if (ts > 0) fetch_tiles(RF, rec - D::REC + D::REC_RF, D::NTRI1, warp, lane);
    }
    cp_async_wait();
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
    const bool any_partial = si[2] != 0;
    __threadfence();  // the reductions into PS / gT / gZ / HB were issued by all warps; the epilogue below reads PS, HB
    __syncthreads();
#ifdef BKF_SQC_PROFILE
    if (b == 0 && lane == 0 && G.g_T)
        for (int i = 0; i < 16; ++i) G.g_T[(size_t)b * m * m + 64 + warp * 16 + i] = (double)tb[i];
    __syncthreads();
#endif
    if (warp != 0) return;
    double* E = smem;  // every array of the sweep is dead: one row-major scratch block
    if (G.g_H) {
        // Hcbar^T = sum_t Hc^T Psi11 = Hc^T (sum_t Psi11)   (a partially missing observation: the reference's
        // cholesky(W H) fails there -> NaN; H all zero: Hc = H = 0 -> zero)
        double* dst = G.g_H + (size_t)b * p * p;
        double *Hs = E, *L = E + p * p, *Lb = L + p * p, *out = Lb + p * p, *tmp = out + p * p;  // Hs: p*p .. tmp: 2 p*p
        bool Hzero;
        setup_consts<sqw::Dims<MT, PT, RT>>(G, b, tmp, nullptr, tmp + p * p, L, &Hzero, lane);
        // (setup_consts leaves chol(H) row-major in L[0 .. p*p); tmp receives the tile copy, tmp + p*p chol(Q))
        if (Hzero) {
            for (int idx = lane; idx < p * p; idx += 32) dst[idx] = 0.0;
        } else {
            for (int idx = lane; idx < p * p; idx += 32) Hs[idx] = HB[idx];
            __syncwarp();
            // Lb = tril(Hcbar),  Hcbar = (Hc^T HB)^T = HB^T Hc  (HB symmetric up to rounding: use it as accumulated)
            for (int idx = lane; idx < p * p; idx += 32) {
                const int i = idx / p, j = idx - i * p;
                double acc = 0.0;
                for (int k = j; k < p; ++k) acc = fma(Hs[k * p + i], L[k * p + j], acc);  // Hc lower: k >= j
                Lb[idx] = j <= i ? (any_partial ? NAN : acc) : 0.0;
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

}  // namespace sqb
}  // namespace bkf
