// explicit instantiations of the sub-warp kernels for m in {15 16} (see bkf_subwarp.cuh)
#include "bkf_subwarp.cuh"
#include "bkf_subwarp_smooth.cuh"
namespace bkf { namespace sw {
template int launch<15>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<16>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch_smoother<15>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<16>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
} }
