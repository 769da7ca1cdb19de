// bkf_sqrt_bwd4_inst_a.cu -- instantiations of the four-warps-per-series SquareRootFilter adjoint (m/8, p/8, r/8).
#define BKF_SQB_DEFINE_LAUNCH
#include "bkf_sqrt_bwd4_launch.cuh"

namespace bkf {
namespace sqb {
template int launch_t<4, 2, 2>(const KArgs<double>&, cudaStream_t, const char**);
template void sizes_t<4, 2, 2>(size_t*, size_t*);
}  // namespace sqb
}  // namespace bkf
