// bkf_sqrt_warp.cu -- dispatch of the warp-per-series SquareRootFilter kernels (bkf_sqrt_warp.cuh).  The kernels are
// instantiated per shape (m/8, p/8, r/8) in bkf_sqrt_warp_inst_*.cu so that the shapes compile in parallel.
#include "bkf_launch.h"
#include "bkf_sqrt_warp_launch.cuh"

namespace bkf {
namespace sqw {


#define BKF_SQW_SHAPES(X) X(1, 1, 1) X(2, 1, 1) X(3, 1, 1) X(4, 1, 1) X(2, 2, 1) X(2, 2, 2) X(4, 2, 2)

static bool shape_ok(int m, int p, int r) {
#define X(MT, PT, RT) if (m == 8 * MT && p == 8 * PT && r == 8 * RT) return true;
    BKF_SQW_SHAPES(X)
#undef X
    return false;
}

bool supported(const KArgs<double>& g) { return g.kind == BKF_CHOLESKY && !g.tv_mask && shape_ok(g.m, g.p, g.r); }

int launch_f64(const KArgs<double>& g, cudaStream_t stream, const char** err) {
#define X(MT, PT, RT) if (g.m == 8 * MT && g.p == 8 * PT && g.r == 8 * RT) return launch_t<MT, PT, RT>(g, stream, err);
    BKF_SQW_SHAPES(X)
#undef X
    *err = "shape not instantiated for the warp-per-series square-root kernels";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace sqw
}  // namespace bkf
