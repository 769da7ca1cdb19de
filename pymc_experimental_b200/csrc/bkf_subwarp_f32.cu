// bkf_subwarp_f32.cu -- dispatcher of the float32 sub-warp kernels (bkf_subwarp.cuh compiled with real = float, namespace
// swf): StandardFilter / UnivariateFilter forward and adjoint, p = 1, 5 <= m <= 16.  Instantiations in
// bkf_subwarp_f32_inst_*.cu.
#define BKF_SW_F32
#include "bkf_subwarp.cuh"

#include "bkf_launch.h"

namespace bkf {
namespace swf {

#define BKF_SWF_DECL(M) extern template int launch<M>(const KArgs<float>&, bool, cudaStream_t, const char**);
BKF_SWF_DECL(5) BKF_SWF_DECL(6) BKF_SWF_DECL(7) BKF_SWF_DECL(8) BKF_SWF_DECL(9) BKF_SWF_DECL(10) BKF_SWF_DECL(11)
BKF_SWF_DECL(12) BKF_SWF_DECL(13) BKF_SWF_DECL(14) BKF_SWF_DECL(15) BKF_SWF_DECL(16)
#undef BKF_SWF_DECL

bool supported(int kind, int m, int p) {
    if (p != 1) return false;
    if (kind != BKF_STANDARD && kind != BKF_UNIVARIATE) return false;
    return m >= 5 && m <= 16;
}

int launch_f32(const KArgs<float>& g, bool backward, cudaStream_t stream, const char** err) {
    switch (g.m) {
#define BKF_SWF_CASE(M) case M: return launch<M>(g, backward, stream, err);
        BKF_SWF_CASE(5) BKF_SWF_CASE(6) BKF_SWF_CASE(7) BKF_SWF_CASE(8) BKF_SWF_CASE(9) BKF_SWF_CASE(10) BKF_SWF_CASE(11)
        BKF_SWF_CASE(12) BKF_SWF_CASE(13) BKF_SWF_CASE(14) BKF_SWF_CASE(15) BKF_SWF_CASE(16)
#undef BKF_SWF_CASE
    }
    *err = "no float32 sub-warp kernel for this m";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace swf
}  // namespace bkf
