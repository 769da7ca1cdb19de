// bkf_sqrt_cta.cu -- dispatch of the four-warps-per-series SquareRootFilter kernels (bkf_sqrt_cta.cuh).
#include <cstdlib>

#include "bkf_launch.h"
#include "bkf_sqrt_cta_launch.cuh"

namespace bkf {
namespace sqc {

// BKF_SQRT_CTA (debugging aid): bit 0 = forward sweep of filter-only calls (its record is not read by any adjoint yet),
// bit 1 = adjoint sweep (reads the record of either warp-per-series / CTA-per-series forward family)
static int enabled_mask() {
    static const int mask = [] {
        const char* e = std::getenv("BKF_SQRT_CTA");
        return e ? std::atoi(e) : 0;
    }();
    return mask;
}

#define BKF_SQC_SHAPES(X) X(2, 1, 1) X(3, 1, 1) X(2, 2, 2) X(4, 2, 2)

static bool shape_ok(int m, int p, int r) {
#define X(MT, PT, RT) if (m == 8 * MT && p == 8 * PT && r == 8 * RT) return true;
    BKF_SQC_SHAPES(X)
#undef X
    return false;
}

bool supported(const KArgs<double>& g, bool backward) {
    if (g.kind != BKF_CHOLESKY || g.tv_mask || !shape_ok(g.m, g.p, g.r)) return false;
    if (backward) return g.traj != nullptr && ((enabled_mask() >> 1) & 1);
    return g.traj == nullptr && (enabled_mask() & 1);
}

int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
#define X(MT, PT, RT) if (g.m == 8 * MT && g.p == 8 * PT && g.r == 8 * RT) return launch_t<MT, PT, RT>(g, backward, stream, err);
    BKF_SQC_SHAPES(X)
#undef X
    *err = "shape not instantiated for the four-warp square-root kernels";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace sqc
}  // namespace bkf
