// bkf_sqrt_warp_launch.cuh -- per-shape launcher of the warp-per-series SquareRootFilter kernels.
#pragma once
#include <cuda_runtime.h>

#include "bkf_common.cuh"

namespace bkf {
namespace sqw {

template <int MT, int PT, int RT>
int launch_t(const KArgs<double>& g, cudaStream_t stream, const char** err);

}  // namespace sqw
}  // namespace bkf

#ifdef BKF_SQW_DEFINE_LAUNCH
#include "bkf_sqrt_warp.cuh"
namespace bkf {
namespace sqw {

template <int MT, int PT, int RT>
int launch_t(const KArgs<double>& g, cudaStream_t stream, const char** err) {
    using D = Dims<MT, PT, RT>;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool tzs = g.B <= 4 * sms;  // one series per scheduler at most: shared memory is free, latency is not
    const size_t bytes = ((size_t)D::FWD_ELEMS + (tzs ? (size_t)(MT * MT + PT * MT) * 64 : 0)) * sizeof(double);
    if (bytes > 227 * 1024) { *err = "shape too large for the warp-per-series square-root kernels"; return BKF_ERR_UNSUPPORTED; }
    auto kern = tzs ? forward_kernel<MT, PT, RT, true> : forward_kernel<MT, PT, RT, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<g.B, 32, bytes, stream>>>(g);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

}  // namespace sqw
}  // namespace bkf
#endif
