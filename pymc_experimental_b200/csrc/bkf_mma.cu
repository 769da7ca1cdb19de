// bkf_mma.cu -- launcher of the FP64 tensor-core (DMMA) warp-per-series kernels (see bkf_mma.cuh).
#include "bkf_mma.cuh"
#include "bkf_mma_smooth.cuh"

#include "bkf_launch.h"

namespace bkf {
namespace tc {

bool supported(int kind, int m, int p) {
    if (p != 1) return false;
    if (kind != BKF_STANDARD && kind != BKF_UNIVARIATE) return false;
    return m >= 9 && m <= MP;
}

size_t traj_record_elems() { return REC; }

int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    const int warps = 2;
    const int grid = (g.B + warps - 1) / warps;
    const size_t smem = (size_t)warps * SCRATCH * sizeof(double);
    // OUT: the per-step filter outputs are requested (the logp+grad path only needs logp and the trajectory)
    const bool out = g.filt_a || g.filt_P || g.pred_a || g.pred_P || g.obs_mu || g.obs_cov;
    const bool tv = g.tv_mask != 0;  // c, d, Z, H with a time axis (tv_supported): re-read per step
#define BKF_TC_LAUNCH(KIND)                                                                                        \
    if (backward) {                                                                                                \
        if (tv) backward_kernel<KIND, true><<<grid, 32 * warps, smem, stream>>>(g);                                \
        else backward_kernel<KIND, false><<<grid, 32 * warps, smem, stream>>>(g);                                  \
    } else if (out) {                                                                                              \
        if (tv) forward_kernel<KIND, true, 1, true><<<grid, 32 * warps, smem, stream>>>(g);                        \
        else forward_kernel<KIND, true, 1, false><<<grid, 32 * warps, smem, stream>>>(g);                          \
    } else { /* 128 registers: 8 CTAs/SM */                                                                        \
        if (tv) forward_kernel<KIND, false, 8, true><<<grid, 32 * warps, smem, stream>>>(g);                       \
        else forward_kernel<KIND, false, 8, false><<<grid, 32 * warps, smem, stream>>>(g);                         \
    }
    if (g.kind == BKF_UNIVARIATE) { BKF_TC_LAUNCH(K_UNI) } else { BKF_TC_LAUNCH(K_STD) }
#undef BKF_TC_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

bool smoother_supported(int m) { return m >= 9 && m <= MP; }

// time axes the kernels take: on the vectors and scalars of the observation / state equations only
bool tv_supported(unsigned tv_mask) {
    const unsigned ok = (1u << BKF_PARAM_C) | (1u << BKF_PARAM_D) | (1u << BKF_PARAM_Z) | (1u << BKF_PARAM_H);
    return (tv_mask & ~ok) == 0;
}

int launch_smoother_f64(const KArgs<double>& g, double jitter0, int* ill_list, cudaStream_t stream, const char** err) {
    const int warps = 2;
    const int grid = (g.B + warps - 1) / warps;
    const size_t smem = (size_t)warps * SCRATCH * sizeof(double);
    smoother_kernel<<<grid, 32 * warps, smem, stream>>>(g, jitter0, ill_list);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace tc
}  // namespace bkf
