// bkf_subwarp.cu -- instantiations + dispatcher of the sub-warp kernels (see bkf_subwarp.cuh).
#include "bkf_subwarp.cuh"

#include "bkf_launch.h"

namespace bkf {
namespace sw {

bool supported(int kind, int m, int p) {
    if (p != 1) return false;
    if (kind != BKF_STANDARD && kind != BKF_UNIVARIATE) return false;
    return m == 15;
}

int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    switch (g.m) {
        case 15: return launch<15>(g, backward, stream, err);
    }
    *err = "no sub-warp kernel for this m";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace sw
}  // namespace bkf
