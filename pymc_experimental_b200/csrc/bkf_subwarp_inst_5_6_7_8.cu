// explicit instantiations of the sub-warp kernels for m in {5 6 7 8} (see bkf_subwarp.cuh)
#include "bkf_subwarp.cuh"
#include "bkf_subwarp_smooth.cuh"
namespace bkf { namespace sw {
template int launch<5>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<6>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<7>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<8>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch_smoother<5>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<6>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<7>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<8>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
} }
