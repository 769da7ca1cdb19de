// bkf_subwarp.cu -- dispatcher of the sub-warp kernels (see bkf_subwarp.cuh); instantiations live in
// bkf_subwarp_inst_*.cu so that nvcc compiles them in parallel.
#include "bkf_subwarp.cuh"
#include "bkf_subwarp_smooth.cuh"

#include "bkf_launch.h"

namespace bkf {
namespace sw {

#define BKF_SW_DECL(M)                                                                          \
    extern template int launch<M>(const KArgs<double>&, bool, cudaStream_t, const char**);      \
    extern template int launch_smoother<M>(const KArgs<double>&, double, int*, cudaStream_t, const char**);
BKF_SW_DECL(5) BKF_SW_DECL(6) BKF_SW_DECL(7) BKF_SW_DECL(8)
#undef BKF_SW_DECL

bool supported(int kind, int m, int p) {
    if (p != 1) return false;
    if (kind != BKF_STANDARD && kind != BKF_UNIVARIATE) return false;
    return m >= 5 && m <= 8;  // 9 <= m <= 16 is served by the FP64 tensor-core family (bkf_mma.cu)
}

int launch_f64(const KArgs<double>& g, bool backward, cudaStream_t stream, const char** err) {
    switch (g.m) {
#define BKF_SW_CASE(M) case M: return launch<M>(g, backward, stream, err);
        BKF_SW_CASE(5) BKF_SW_CASE(6) BKF_SW_CASE(7) BKF_SW_CASE(8)
#undef BKF_SW_CASE
    }
    *err = "no sub-warp kernel for this m";
    return BKF_ERR_UNSUPPORTED;
}

bool smoother_supported(int m) { return m >= 5 && m <= 8; }

int launch_smoother_f64(const KArgs<double>& g, double jitter0, int* ill_list, cudaStream_t stream, const char** err) {
    switch (g.m) {
#define BKF_SW_CASE(M) case M: return launch_smoother<M>(g, jitter0, ill_list, stream, err);
        BKF_SW_CASE(5) BKF_SW_CASE(6) BKF_SW_CASE(7) BKF_SW_CASE(8)
#undef BKF_SW_CASE
    }
    *err = "no sub-warp smoother kernel for this m";
    return BKF_ERR_UNSUPPORTED;
}

}  // namespace sw
}  // namespace bkf
