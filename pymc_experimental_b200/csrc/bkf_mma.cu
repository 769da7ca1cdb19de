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
    if (g.kind == BKF_UNIVARIATE) {
        if (backward) backward_kernel<K_UNI><<<grid, 32 * warps, smem, stream>>>(g);
        else if (out) forward_kernel<K_UNI, true, 1><<<grid, 32 * warps, smem, stream>>>(g);
        else forward_kernel<K_UNI, false, 8><<<grid, 32 * warps, smem, stream>>>(g);  // 128 registers: 8 CTAs/SM
    } else {
        if (backward) backward_kernel<K_STD><<<grid, 32 * warps, smem, stream>>>(g);
        else if (out) forward_kernel<K_STD, true, 1><<<grid, 32 * warps, smem, stream>>>(g);
        else forward_kernel<K_STD, false, 8><<<grid, 32 * warps, smem, stream>>>(g);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

bool smoother_supported(int m) { return m >= 9 && m <= MP; }

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
