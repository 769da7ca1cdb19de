// explicit instantiations of the sub-warp kernels for m in {9 10 11} (see bkf_subwarp.cuh)
#include "bkf_subwarp.cuh"
#include "bkf_subwarp_smooth.cuh"
namespace bkf { namespace sw {
template int launch<9>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<10>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch<11>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch_smoother<9>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<10>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
template int launch_smoother<11>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
} }
