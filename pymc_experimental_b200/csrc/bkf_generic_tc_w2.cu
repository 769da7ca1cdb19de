// bkf_generic_tc_w2.cu -- the DMMA build of the generic kernels (namespace gent, float64) with one series per CTA of 2
// warps: forward and adjoint sweeps for shapes whose arena leaves an SM only a few series (see pick_w in bkf_generic.cu).
#define BKF_GEN_NS gent
#define BKF_GEN_DMMA
#define BKF_GEN_W 2
#include "bkf_generic.cu"
