// bkf_common.cuh -- shared declarations for the libbkf kernels (sm_100a).
//
// Math restated from the reference (paths relative to pymc-devs/pymc-experimental):
//   pymc_extras/statespace/filters/kalman_filter.py    (kalman_step :529-539, StandardFilter.update :594-617,
//                                                       SquareRootFilter :629-716, UnivariateFilter :769-820)
//   pymc_extras/statespace/filters/kalman_smoother.py  (:66-126)
//   pymc_extras/statespace/filters/utilities.py:50-59  (stabilize, quad_form_sym)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/bkf.h"

namespace bkf {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;  // kalman_filter.py:21 (MVN_CONST)

// Device-side argument block shared by every kernel family.  All pointers are device pointers.
template <typename R>
struct KArgs {
    int kind, B, n, m, p, r, y_shared;
    unsigned shared_mask;
    unsigned tv_mask;  // bit BKF_PARAM_x: parameter x carries a time axis ([B, n, ...] or [n, ...] when shared)
    R jitter, fill;
    // inputs
    const R *y, *a0, *P0, *c, *d, *T, *Z, *Rm, *H, *Q;
    // forward outputs (nullable)
    R *filt_a, *pred_a, *obs_mu, *filt_P, *pred_P, *obs_cov, *ll;
    // forward -> backward trajectory of predictions: [B, n, m + m*m] (nullable)
    R* traj;
    int rec_tiles;  // square-root family: the record holds the factors as 8 x 8 tiles (bkf_sqrt_bwd4.cuh)
    // logp + gradients (nullable)
    const R* ll_cot;
    R *logp, *g_a0, *g_P0, *g_c, *g_d, *g_T, *g_Z, *g_R, *g_H, *g_Q;
    // smoother
    const R *sm_filt_a, *sm_filt_P;
    R *sm_a, *sm_P;
    int* status;
};

template <typename R>
__device__ __forceinline__ const R* param_ptr(const R* base, unsigned shared_mask, int which, int b, size_t sz) {
    return (shared_mask >> which) & 1u ? base : base + (size_t)b * sz;
}

// Pointer to parameter `which` of series b at time step t (t is ignored for time-invariant parameters).
template <typename R>
__device__ __forceinline__ const R* param_ptr_t(const R* base, unsigned shared_mask, unsigned tv_mask, int which, int b,
                                                int t, int n, size_t sz) {
    const bool tv = (tv_mask >> which) & 1u, sh = (shared_mask >> which) & 1u;
    const size_t rec = (sh ? 0 : (size_t)b) * (tv ? (size_t)n : 1) + (tv ? (size_t)t : 0);
    return base + rec * sz;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename R>
__device__ __forceinline__ bool is_missing(R y, R fill, bool use_fill) {
    return isnan(y) || (use_fill && y == fill);
}

}  // namespace bkf
