// bkf_sqrt_cta_inst_b.cu -- instantiations of the four-warps-per-series SquareRootFilter kernels (m/8, p/8, r/8).
#define BKF_SQC_DEFINE_LAUNCH
#include "bkf_sqrt_cta_launch.cuh"

namespace bkf {
namespace sqc {
template int launch_t<2, 1, 1>(const KArgs<double>&, bool, cudaStream_t, const char**);
template size_t record_elems_t<2, 1, 1>();
template int launch_t<2, 2, 2>(const KArgs<double>&, bool, cudaStream_t, const char**);
template size_t record_elems_t<2, 2, 2>();
template int launch_t<3, 1, 1>(const KArgs<double>&, bool, cudaStream_t, const char**);
template size_t record_elems_t<3, 1, 1>();
}  // namespace sqc
}  // namespace bkf
