// bkf_sqrt_warp_inst_c.cu -- instantiations of the warp-per-series SquareRootFilter kernels (m/8, p/8, r/8).
#define BKF_SQW_DEFINE_LAUNCH
#include "bkf_sqrt_warp_launch.cuh"

namespace bkf {
namespace sqw {
template int launch_t<3, 1, 1>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_t<4, 1, 1>(const KArgs<double>&, cudaStream_t, const char**);
template int launch_t<2, 2, 1>(const KArgs<double>&, cudaStream_t, const char**);
}  // namespace sqw
}  // namespace bkf
