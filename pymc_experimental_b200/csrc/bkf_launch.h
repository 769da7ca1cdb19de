// bkf_launch.h -- host-side launcher declarations of the kernel families.
#pragma once
#include <cuda_runtime.h>

#include "bkf_common.cuh"

namespace bkf {
namespace gen {  // generic warp-per-series kernels (bkf_generic.cu)
template <typename R> int launch_forward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_backward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_smoother(const KArgs<R>& g, cudaStream_t stream, const char** err);
}  // namespace gen
namespace sq {  // square-root filter, CTA per series (bkf_sqrt.cu)
size_t traj_record_elems_host(int m, int p);  // forward-sweep record kept for the adjoint (a, Pc, R factors)
template <typename R> int launch_forward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_backward(const KArgs<R>& g, cudaStream_t stream, const char** err);
}  // namespace sq
namespace sqw {  // square-root filter, warp per series on DMMA tiles, float64, selected shapes (bkf_sqrt_warp.cu)
bool supported(const KArgs<double>& g, bool backward);
int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace sqw
namespace sqc {  // square-root filter, four warps per series with a look-ahead Householder chain (bkf_sqrt_cta.cu)
bool supported(const KArgs<double>& g, bool backward);
int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace sqc
namespace sw {  // register-tiled sub-warp-per-series kernels, float64, p = 1 (bkf_subwarp.cu)
bool supported(int kind, int m, int p);
int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
bool smoother_supported(int m);
int launch_smoother_f64(const KArgs<double>& g, double jitter0, int* any_ill, cudaStream_t stream, const char** err);
}  // namespace sw
namespace tc {  // FP64 tensor-core (DMMA) warp-per-series kernels, 9 <= m <= 16, p = 1 (bkf_mma.cu)
bool supported(int kind, int m, int p);
size_t traj_record_elems();
int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
bool smoother_supported(int m);
int launch_smoother_f64(const KArgs<double>& g, double jitter0, int* any_ill, cudaStream_t stream, const char** err);
}  // namespace tc
namespace stat {  // stationary initialisation: Lyapunov / (I - T)^{-1} c by doubling (bkf_stationary.cu)
template <typename R> int launch(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace stat
namespace smp {  // multivariate-normal draws from (mu, cov) sequences (bkf_sample.cu)
template <typename R>
int launch(const R* mu, const R* cov, const R* z, R* x, int* status, long long N, int n, int m, cudaStream_t stream,
           const char** err);
template <typename R>
int launch_simulate(const KArgs<R>& g, const R* z_x0, const R* z_y0, const R* z_state, const R* z_obs, R* xs, R* ys,
                    cudaStream_t stream, const char** err);
}  // namespace smp
namespace tps {  // thread-per-series kernels, m <= 4, p = 1 (bkf_thread.cu)
bool supported(int kind, int m, int p);
template <typename R> int launch_any(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace tps
}  // namespace bkf
