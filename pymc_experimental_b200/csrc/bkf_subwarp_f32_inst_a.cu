// explicit instantiations of the float32 sub-warp kernels for m in (5, 6, 7, 8) (see bkf_subwarp_f32.cu)
#define BKF_SW_F32
#include "bkf_subwarp.cuh"
namespace bkf { namespace swf {
template int launch<5>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<6>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<7>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<8>(const KArgs<float>&, bool, cudaStream_t, const char**);
} }
