// bkf_thread.cu -- instantiations + dispatcher of the thread-per-series kernels (see bkf_thread.cuh).
#include "bkf_thread.cuh"

#include "bkf_launch.h"

namespace bkf {
namespace tps {

bool supported(int kind, int m, int p) {
    return p == 1 && m >= 1 && m <= 4 && (kind == BKF_STANDARD || kind == BKF_UNIVARIATE);
}

template <typename R>
int launch_any(const KArgs<R>& g, bool backward, cudaStream_t stream, const char** err) {
    switch (g.m) {
        case 1: return launch<R, 1>(g, backward, stream, err);
        case 2: return launch<R, 2>(g, backward, stream, err);
        case 3: return launch<R, 3>(g, backward, stream, err);
        case 4: return launch<R, 4>(g, backward, stream, err);
    }
    *err = "no thread-per-series kernel for this m";
    return BKF_ERR_UNSUPPORTED;
}
template int launch_any<double>(const KArgs<double>&, bool, cudaStream_t, const char**);
template int launch_any<float>(const KArgs<float>&, bool, cudaStream_t, const char**);

}  // namespace tps
}  // namespace bkf
