// bkf_launch.h -- host-side launcher declarations of the kernel families.
#pragma once
#include <cuda_runtime.h>

#include "bkf_common.cuh"

namespace bkf {
namespace gen {  // generic warp-per-series kernels (bkf_generic.cu)
template <typename R> int launch_forward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_backward(const KArgs<R>& g, cudaStream_t stream, const char** err);
// list / count (nullable): only the series list[0 .. *count), both on the device
template <typename R> int launch_smoother(const KArgs<R>& g, const int* list, const int* count, cudaStream_t stream, const char** err);
}  // namespace gen
namespace gent {  // the same kernels with DMMA tile products, float64 (bkf_generic_tc.cu); the dispatcher uses them for m >= 16
template <typename R> int launch_forward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_backward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_smoother(const KArgs<R>& g, const int* list, const int* count, cudaStream_t stream, const char** err);
int width(const KArgs<double>& g, bool backward);  // warps per series the launcher picks for this shape (1: warp per series)
}  // namespace gent
namespace sq {  // square-root filter, CTA per series (bkf_sqrt.cu)
size_t traj_record_elems_host(int m, int p);  // forward-sweep record kept for the adjoint (a, Pc, R factors)
template <typename R> int launch_forward(const KArgs<R>& g, cudaStream_t stream, const char** err);
template <typename R> int launch_backward(const KArgs<R>& g, cudaStream_t stream, const char** err);
}  // namespace sq
namespace sqw {  // square-root filter, warp per series on DMMA tiles, float64, selected shapes (bkf_sqrt_warp.cu)
bool supported(const KArgs<double>& g);
int launch_f64(const KArgs<double>& g, cudaStream_t stream, const char** err);  // forward sweep; its adjoint is sqb
}  // namespace sqw
namespace sqb {  // square-root filter adjoint, four warps per series, tile-format record (bkf_sqrt_bwd4.cu)
bool tile_record(int kind, int m, int p, int r, unsigned tv_mask);  // the sqw forward / sqb adjoint pair handles it
void sizes(int m, int p, int r, size_t* record, size_t* scratch);   // doubles per (series, step) / per series
int launch_f64(const KArgs<double>& g, cudaStream_t stream, const char** err);
}  // namespace sqb
namespace sw {  // register-tiled sub-warp-per-series kernels, float64, p = 1 (bkf_subwarp.cu)
bool supported(int kind, int m, int p);
int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
bool smoother_supported(int m);
int launch_smoother_f64(const KArgs<double>& g, double jitter0, int* ill, cudaStream_t stream, const char** err);
}  // namespace sw
namespace swf {  // the same kernels in float32, 5 <= m <= 16, p = 1 (bkf_subwarp_f32.cu)
bool supported(int kind, int m, int p);
int launch_f32(const KArgs<float>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace swf
namespace tc {  // FP64 tensor-core (DMMA) warp-per-series kernels, 9 <= m <= 16, p = 1 (bkf_mma.cu)
bool supported(int kind, int m, int p);
bool tv_supported(unsigned tv_mask);  // time axes on c, d, Z, H only
size_t traj_record_elems();
int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
bool smoother_supported(int m);
int launch_smoother_f64(const KArgs<double>& g, double jitter0, int* ill, cudaStream_t stream, const char** err);
}  // namespace tc
namespace stat {  // stationary initialisation: Lyapunov / (I - T)^{-1} c by doubling (bkf_stationary.cu)
template <typename R> int launch(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace stat
namespace smp {  // multivariate-normal draws from (mu, cov) sequences (bkf_sample.cu)
struct Rng { unsigned long long seed, offset; };  // counter-based generator for NULL variate pointers (bkf_problem.rng_*)
template <typename R>
int launch(const R* mu, const R* cov, const R* z, Rng rng, R* x, int* status, long long N, int n, int m,
           cudaStream_t stream, const char** err);
template <typename R>
int launch_simulate(const KArgs<R>& g, const R* z_x0, const R* z_y0, const R* z_state, const R* z_obs, Rng rng, R* xs,
                    R* ys, cudaStream_t stream, const char** err);
}  // namespace smp
namespace tps {  // thread-per-series kernels, m <= 4, p = 1 (bkf_thread.cu)
bool supported(int kind, int m, int p);
template <typename R> int launch_any(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err);
}  // namespace tps
}  // namespace bkf
