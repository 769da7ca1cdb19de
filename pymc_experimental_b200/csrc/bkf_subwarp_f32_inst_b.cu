// explicit instantiations of the float32 sub-warp kernels for m in (9, 10, 11, 12) (see bkf_subwarp_f32.cu)
#define BKF_SW_F32
#include "bkf_subwarp.cuh"
namespace bkf { namespace swf {
template int launch<9>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<10>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<11>(const KArgs<float>&, bool, cudaStream_t, const char**);
template int launch<12>(const KArgs<float>&, bool, cudaStream_t, const char**);
} }
