// explicit instantiations of the sub-warp kernels for m in {12 13 14} (see bkf_subwarp.cuh)
#include "bkf_subwarp.cuh"
#include "bkf_subwarp_smooth.cuh"
namespace bkf { namespace sw {
template int launch<12>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<13>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<14>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch_smoother<12>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<13>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<14>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
} }
