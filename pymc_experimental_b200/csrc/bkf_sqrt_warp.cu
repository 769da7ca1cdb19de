// bkf_sqrt_warp.cu -- launchers of the warp-per-series SquareRootFilter kernels (bkf_sqrt_warp.cuh).
#include <cstdlib>

#include "bkf_launch.h"
#include "bkf_sqrt_warp.cuh"

namespace bkf {
namespace sqw {

// BKF_SQRT_WARP (debugging aid): bit 0 = forward sweep, bit 1 = adjoint sweep on the warp-per-series kernels (default 3)
static int enabled_mask() {
    static const int mask = [] {
        const char* e = std::getenv("BKF_SQRT_WARP");
        return e ? std::atoi(e) : 3;
    }();
    return mask;
}

static bool shape_ok(int m, int p, int r) {
    return (m == 32 && p == 16 && r == 16) || (m == 16 && p == 8 && r == 8) || (m == 16 && p == 16 && r == 8);
}

bool supported(const KArgs<double>& g, bool backward) {
    if (g.kind != BKF_CHOLESKY || g.tv_mask || !shape_ok(g.m, g.p, g.r)) return false;
    if (!((enabled_mask() >> (backward ? 1 : 0)) & 1)) return false;
    if (backward) return g.traj != nullptr;
    // the forward sweep of this family only produces the log-likelihood and the adjoint's record
    return !(g.filt_a || g.pred_a || g.obs_mu || g.filt_P || g.pred_P || g.obs_cov);
}

template <int MT, int PT, int RT>
static int launch_t(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    using D = Dims<MT, PT, RT>;
    using BD = BDims<MT, PT, RT>;
    const size_t bytes = (size_t)(backward ? BD::ELEMS : D::FWD_ELEMS) * sizeof(double);
    auto kern = backward ? backward_kernel<MT, PT, RT> : forward_kernel<MT, PT, RT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    kern<<<g.B, 32, bytes, stream>>>(g);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BKF_ERR_CUDA; }
    return BKF_OK;
}

int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    if (g.m == 32 && g.p == 16 && g.r == 16) return launch_t<4, 2, 2>(g, backward, stream, err);
    if (g.m == 16 && g.p == 8 && g.r == 8) return launch_t<2, 1, 1>(g, backward, stream, err);
    if (g.m == 16 && g.p == 16 && g.r == 8) return launch_t<2, 2, 1>(g, backward, stream, err);
    *err = "shape not instantiated for the warp-per-series square-root kernels";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace sqw
}  // namespace bkf
