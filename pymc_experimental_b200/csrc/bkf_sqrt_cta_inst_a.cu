// bkf_sqrt_cta_inst_a.cu -- instantiations of the four-warps-per-series SquareRootFilter kernels (m/8, p/8, r/8).
#define BKF_SQC_DEFINE_LAUNCH
#include "bkf_sqrt_cta_launch.cuh"

namespace bkf {
namespace sqc {
template int launch_t<4, 2, 2>(const KArgs<double>&, bool, cudaStream_t, const char**);
template size_t record_elems_t<4, 2, 2>();
}  // namespace sqc
}  // namespace bkf
