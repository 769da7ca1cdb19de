// bkf_sqrt_bwd4.cu -- dispatch of the four-warps-per-series SquareRootFilter adjoint (bkf_sqrt_bwd4.cuh).  It goes with
// the warp-per-series forward sweep (bkf_sqrt_warp.cuh) writing the tile-format record.
#include "bkf_launch.h"
#include "bkf_sqrt_bwd4_launch.cuh"

namespace bkf {
namespace sqb {

#define BKF_SQB_SHAPES(X) X(1, 1, 1) X(2, 1, 1) X(3, 1, 1) X(4, 1, 1) X(2, 2, 1) X(2, 2, 2) X(4, 2, 2)

bool tile_record(int kind, int m, int p, int r, unsigned tv_mask) {
    if (kind != BKF_CHOLESKY || tv_mask) return false;
#define X(MT, PT, RT) if (m == 8 * MT && p == 8 * PT && r == 8 * RT) return true;
    BKF_SQB_SHAPES(X)
#undef X
    return false;
}

void sizes(int m, int p, int r, size_t* record, size_t* scratch) {
    *record = *scratch = 0;
#define X(MT, PT, RT) if (m == 8 * MT && p == 8 * PT && r == 8 * RT) return sizes_t<MT, PT, RT>(record, scratch);
    BKF_SQB_SHAPES(X)
#undef X
}

int launch_f64(const KArgs<double>& g, cudaStream_t stream, const char** err) {
#define X(MT, PT, RT) if (g.m == 8 * MT && g.p == 8 * PT && g.r == 8 * RT) return launch_t<MT, PT, RT>(g, stream, err);
    BKF_SQB_SHAPES(X)
#undef X
    *err = "shape not instantiated for the four-warp square-root adjoint";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace sqb
}  // namespace bkf
