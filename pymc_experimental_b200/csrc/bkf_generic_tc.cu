// bkf_generic_tc.cu -- the generic warp-per-series kernels once more (bkf_generic.cu, namespace gent, float64 only) with
// their dense products on the FP64 tensor pipe (mm_dmma); used for m >= 16.
#define BKF_GEN_NS gent
#define BKF_GEN_DMMA
#include "bkf_generic.cu"
