// explicit instantiations of the float32 sub-warp kernels for m in (13, 14, 15, 16) (see bkf_subwarp_f32.cu)
#define BKF_SW_F32
#include "bkf_subwarp.cuh"
namespace bkf { namespace swf {
template int launch<13>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<14>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<15>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<16>(const KArgs<float>&, bool, cudaStream_t, const char**);
} }
