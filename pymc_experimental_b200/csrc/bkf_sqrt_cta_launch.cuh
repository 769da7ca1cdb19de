// bkf_sqrt_cta_launch.cuh -- per-shape launcher of the four-warps-per-series SquareRootFilter kernels.
#pragma once
#include <cuda_runtime.h>

#include "bkf_common.cuh"

namespace bkf {
namespace sqc {

template <int MT, int PT, int RT>
int launch_t(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err);
template <int MT, int PT, int RT>
size_t record_elems_t();

}  // namespace sqc
}  // namespace bkf

#ifdef BKF_SQC_DEFINE_LAUNCH
#include "bkf_sqrt_cta.cuh"
namespace bkf {
namespace sqc {

template <int MT, int PT, int RT>
size_t record_elems_t() {
    return Dims<MT, PT, RT>::RSTRIDE;
}

template <int MT, int PT, int RT>
int launch_t(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    using D = Dims<MT, PT, RT>;
    using BD = BDimsC<MT, PT, RT>;
    const size_t bytes = (size_t)(backward ? BD::ELEMS : D::FWD_ELEMS) * sizeof(double);
    if (bytes > 227 * 1024) { *err = "shape too large for the four-warp square-root kernels"; return BKF_ERR_UNSUPPORTED; }
    auto kern = backward ? backward_kernel<MT, PT, RT> : forward_kernel<MT, PT, RT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<g.B, NTHR, bytes, stream>>>(g);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace sqc
}  // namespace bkf
#endif
