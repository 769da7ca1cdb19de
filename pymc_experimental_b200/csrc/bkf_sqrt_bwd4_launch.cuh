// bkf_sqrt_bwd4_launch.cuh -- per-shape launcher of the four-warps-per-series SquareRootFilter adjoint.
#pragma once
#include <cuda_runtime.h>

#include "bkf_common.cuh"

namespace bkf {
namespace sqb {

template <int MT, int PT, int RT>
int launch_t(const KArgs<double>& g, cudaStream_t stream, const char** err);
template <int MT, int PT, int RT>
void sizes_t(size_t* record, size_t* scratch);

}  // namespace sqb
}  // namespace bkf

#ifdef BKF_SQB_DEFINE_LAUNCH
#include "bkf_sqrt_bwd4.cuh"
namespace bkf {
namespace sqb {

template <int MT, int PT, int RT>
void sizes_t(size_t* record, size_t* scratch) {
    *record = BD<MT, PT, RT>::REC;
    *scratch = BD<MT, PT, RT>::SCR;
}

template <int MT, int PT, int RT>
int launch_t(const KArgs<double>& g, cudaStream_t stream, const char** err) {
    using D = BD<MT, PT, RT>;
    const size_t bytes = (size_t)D::ELEMS * sizeof(double);
    if (bytes > 227 * 1024) { *err = "shape too large for the four-warp square-root adjoint"; return BKF_ERR_UNSUPPORTED; }
    auto kern = backward_kernel<MT, PT, RT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<g.B, NTHR, bytes, stream>>>(g);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace sqb
}  // namespace bkf
#endif
