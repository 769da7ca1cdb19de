// bkf_api.cu -- the C ABI of libbkf.so (include/bkf.h): handle, staging of host buffers, kernel dispatch.
//
// No torch types, no exceptions across the boundary.  Host pointers (NumPy) are staged through a grow-only
// device workspace owned by the handle; device pointers are used in place.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bkf_common.cuh"
#include "bkf_launch.h"

using namespace bkf;

enum Slot {
    S_Y = 0, S_A0, S_P0, S_C, S_D, S_T, S_Z, S_R, S_H, S_Q,            // inputs (order = BKF_PARAM_* + 1)
    S_FA, S_PA, S_OM, S_FP, S_PP, S_OC, S_LL,                          // forward outputs
    S_TRAJ, S_COT, S_LOGP, S_GA0, S_GP0, S_GC, S_GD, S_GT, S_GZ, S_GR, S_GH, S_GQ, S_STATUS,
    S_SMA, S_SMP, S_WFA, S_WFP, S_FLAG, S_U, S_GU, S_MAP, S_BASE, S_THETA, S_GTHETA, S_PROG, S_NSLOT
};

struct bkf_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::string path = "none";
    long long launches = 0;
    struct Buf { void* p = nullptr; size_t cap = 0; };
    Buf slots[S_NSLOT];
    // optional per-kernel timing (bench.py roofline)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[BKF_NKERNEL];
    size_t ev_used[BKF_NKERNEL] = {0, 0, 0};
    double ms_total[BKF_NKERNEL] = {0, 0, 0};
    long long ms_launches[BKF_NKERNEL] = {0, 0, 0};
    // bkf_filter_forward_saved / bkf_filter_backward_saved: `gen` counts the times a workspace slot was handed out
    unsigned long long gen = 0, tickets = 0;
    struct Saved {
        bool valid = false;
        unsigned long long ticket = 0, gen = 0;
        bkf_problem P;
        KArgs<double> gd;
        KArgs<float> gf;
    } saved;
    // bkf_comm_*: the NCCL communicator of this rank (ncclComm_t), NULL until bkf_comm_init
    void* comm = nullptr;
    int comm_world = 0, comm_rank = 0;
};

// RAII bracket: records start/stop events around one kernel launch when timing is on.
struct KernelTimer {
    bkf_handle* h;
    int cls;
    cudaEvent_t stop = nullptr;
    KernelTimer(bkf_handle* h_, int cls_) : h(h_), cls(cls_) {
        if (!h->timing) return;
        auto& v = h->ev[cls];
        if (h->ev_used[cls] == v.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            v.emplace_back(a, b);
        }
        auto& pr = v[h->ev_used[cls]++];
        cudaEventRecord(pr.first, h->stream);
        stop = pr.second;
    }
    ~KernelTimer() {
        if (stop) cudaEventRecord(stop, h->stream);
    }
};

static void collect_timing(bkf_handle* h) {
    for (int c = 0; c < BKF_NKERNEL; ++c) {
        for (size_t i = 0; i < h->ev_used[c]; ++i) {
            float ms = 0;
            cudaEventSynchronize(h->ev[c][i].second);
            if (cudaEventElapsedTime(&ms, h->ev[c][i].first, h->ev[c][i].second) == cudaSuccess) {
                h->ms_total[c] += ms;
                h->ms_launches[c] += 1;
            }
        }
        h->ev_used[c] = 0;
    }
}

static int fail(bkf_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

static int cuda_fail(bkf_handle* h, cudaError_t e, const char* what) {
    return fail(h, BKF_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static void* ensure(bkf_handle* h, int slot, size_t bytes, int* rc) {
    auto& b = h->slots[slot];
    h->gen += 1;
    if (bytes == 0) bytes = 8;
    if (b.cap < bytes) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&b.p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            *rc = fail(h, BKF_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
            cudaGetLastError();
            return nullptr;
        }
        b.cap = want;
    }
    return b.p;
}

extern "C" int bkf_version(void) { return BKF_VERSION; }

extern "C" int bkf_create(int device, bkf_handle** out) {
    if (!out) return BKF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return BKF_ERR_CUDA;  // no CUDA device: there is no CPU fallback by design
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return BKF_ERR_CUDA;
    bkf_handle* h = new (std::nothrow) bkf_handle();
    if (!h) return BKF_ERR_NOMEM;
    h->device = device;
    *out = h;
    return BKF_OK;
}

extern "C" int bkf_destroy(bkf_handle* h) {
    if (!h) return BKF_OK;
    cudaSetDevice(h->device);
    bkf_comm_destroy(h);
    for (auto& b : h->slots)
        if (b.p) cudaFree(b.p);
    for (auto& v : h->ev)
        for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    delete h;
    return BKF_OK;
}

extern "C" const char* bkf_last_error(const bkf_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" const char* bkf_last_path(const bkf_handle* h) { return h ? h->path.c_str() : "none"; }
extern "C" long long bkf_launch_count(const bkf_handle* h) { return h ? h->launches : 0; }
extern "C" int bkf_enable_timing(bkf_handle* h, int on) {
    if (!h) return BKF_ERR_ARG;
    h->timing = on != 0;
    return BKF_OK;
}
extern "C" int bkf_reset_timing(bkf_handle* h) {
    if (!h) return BKF_ERR_ARG;
    collect_timing(h);
    for (int c = 0; c < BKF_NKERNEL; ++c) { h->ms_total[c] = 0; h->ms_launches[c] = 0; }
    return BKF_OK;
}
extern "C" int bkf_kernel_time_ms(bkf_handle* h, int cls, double* total_ms, long long* launches) {
    if (!h || cls < 0 || cls >= BKF_NKERNEL) return BKF_ERR_ARG;
    cudaSetDevice(h->device);
    collect_timing(h);
    if (total_ms) *total_ms = h->ms_total[cls];
    if (launches) *launches = h->ms_launches[cls];
    return BKF_OK;
}

// FMA-pipe peak: 8 independent FMA chains per thread, 1024 threads x 4 CTAs per SM, no memory traffic.
template <typename R>
__global__ void fma_peak_kernel(R* out, int iters) {
    R a0 = R(threadIdx.x) * R(1e-6), a1 = a0 + R(1), a2 = a0 + R(2), a3 = a0 + R(3), a4 = a0 + R(4), a5 = a0 + R(5),
      a6 = a0 + R(6), a7 = a0 + R(7);
    const R x = R(0.999999), y = R(1e-9);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
            a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
        }
    }
    R s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == R(-1)) out[0] = s;
}

extern "C" int bkf_measure_fma_peak(bkf_handle* h, int dtype, double* tflops) {
    if (!h || !tflops) return BKF_ERR_ARG;
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    int rc = BKF_OK;
    void* out = ensure(h, S_STATUS, 64, &rc);
    if (!out) return rc;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, h->device);
    const int grid = prop.multiProcessorCount * 4, block = 512;
    const int iters = dtype == BKF_F64 ? 4096 : 16384;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a, h->stream);
        if (dtype == BKF_F64) fma_peak_kernel<double><<<grid, block, 0, h->stream>>>((double*)out, iters);
        else fma_peak_kernel<float><<<grid, block, 0, h->stream>>>((float*)out, iters);
        cudaEventRecord(b, h->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double flops = 2.0 * 64.0 * iters * (double)grid * block;
        if (rep > 0 && ms > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(h, e, "fma peak kernel");
    *tflops = best;
    return BKF_OK;
}

extern "C" int bkf_host_alloc(size_t bytes, void** out) {
    if (!out) return BKF_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return BKF_ERR_NOMEM;
    }
    return BKF_OK;
}
extern "C" int bkf_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return BKF_ERR_CUDA;
    }
    return BKF_OK;
}

extern "C" int bkf_set_stream(bkf_handle* h, void* s) {
    if (!h) return BKF_ERR_ARG;
    h->stream = reinterpret_cast<cudaStream_t>(s);
    return BKF_OK;
}

// generic kernels, float64: the build with DMMA tile products from m = 16 on (bkf_generic_tc.cu), the scalar one below
static int gen_forward_f64(const KArgs<double>& g, cudaStream_t st, const char** err) {
    return g.m >= 16 ? gent::launch_forward<double>(g, st, err) : gen::launch_forward<double>(g, st, err);
}
static int gen_backward_f64(const KArgs<double>& g, cudaStream_t st, const char** err) {
    return g.m >= 16 ? gent::launch_backward<double>(g, st, err) : gen::launch_backward<double>(g, st, err);
}
static int gen_smoother_f64(const KArgs<double>& g, const int* list, const int* count, cudaStream_t st, const char** err) {
    return g.m >= 16 ? gent::launch_smoother<double>(g, list, count, st, err) : gen::launch_smoother<double>(g, list, count, st, err);
}

// "generic_cta" when the launcher gives a series more than one warp (large arenas, bkf_generic.cu pick_w)
static const char* gen_path_f64(const KArgs<double>& g) {
    return g.m >= 16 && (gent::width(g, false) > 1 || gent::width(g, true) > 1) ? "generic_cta" : "generic_warp";
}

// ----------------------------------------------------------------------------------------------------
// kernel-family dispatch: specialised register-tiled kernels where they exist, the generic path otherwise
static int run_forward(bkf_handle* h, const KArgs<double>& g, const char** err) {
    KernelTimer kt(h, BKF_KERNEL_FORWARD);
    h->launches += 1;
    if (g.kind == BKF_CHOLESKY) {
        if (sqw::supported(g)) {  // (with a trajectory it writes the tile record the four-warp adjoint reads)
            h->path = "sqrt_warp_dmma";
            return sqw::launch_f64(g, h->stream, err);
        }
        h->path = "sqrt_cta";
        return sq::launch_forward<double>(g, h->stream, err);
    }
    // time-varying matrices: the generic kernels stream them per step -- except a time axis on c, d, Z, H only (the
    // regression-component pattern) for p = 1 and m <= 16, which the specialised float64 kernels re-read per step themselves
    const bool tc_tv = g.tv_mask && tc::tv_supported(g.tv_mask) &&
                       (tps::supported(g.kind, g.m, g.p) || tc::supported(g.kind, g.m, g.p) ||
                        sw::supported(g.kind, g.m, g.p));
    if (g.tv_mask && !tc_tv) {
        h->path = gen_path_f64(g);
        return gen_forward_f64(g, h->stream, err);
    }
    if (tps::supported(g.kind, g.m, g.p)) {
        h->path = "thread_f64";
        return tps::launch_any<double>(g, false, h->stream, err);
    }
    if (tc::supported(g.kind, g.m, g.p)) {
        h->path = "dmma_f64";
        return tc::launch_f64(g, false, h->stream, err);
    }
    if (sw::supported(g.kind, g.m, g.p)) {
        h->path = "subwarp_f64";
        return sw::launch_f64(g, false, h->stream, err);
    }
    h->path = gen_path_f64(g);
    return gen_forward_f64(g, h->stream, err);
}
static int run_forward(bkf_handle* h, const KArgs<float>& g, const char** err) {
    KernelTimer kt(h, BKF_KERNEL_FORWARD);
    h->launches += 1;
    if (g.kind == BKF_CHOLESKY) {
        h->path = "sqrt_cta";
        return sq::launch_forward<float>(g, h->stream, err);
    }
    if ((!g.tv_mask || tc::tv_supported(g.tv_mask)) && tps::supported(g.kind, g.m, g.p)) {
        h->path = "thread_f32";
        return tps::launch_any<float>(g, false, h->stream, err);
    }
    if (!g.tv_mask && swf::supported(g.kind, g.m, g.p)) {
        h->path = "subwarp_f32";
        return swf::launch_f32(g, false, h->stream, err);
    }
    h->path = "generic_warp";
    return gen::launch_forward<float>(g, h->stream, err);
}
static int run_backward(bkf_handle* h, const KArgs<double>& g, const char** err) {
    KernelTimer kt(h, BKF_KERNEL_BACKWARD);
    h->launches += 1;
    if (g.kind == BKF_CHOLESKY && g.rec_tiles) {
        h->path += "+adjoint_4warp";
        return sqb::launch_f64(g, h->stream, err);
    }
    if (g.kind == BKF_CHOLESKY) return sq::launch_backward<double>(g, h->stream, err);
    const bool tc_tv = g.tv_mask && tc::tv_supported(g.tv_mask) &&
                       (tps::supported(g.kind, g.m, g.p) || tc::supported(g.kind, g.m, g.p) ||
                        sw::supported(g.kind, g.m, g.p));
    if (g.tv_mask && !tc_tv) return gen_backward_f64(g, h->stream, err);
    if (tps::supported(g.kind, g.m, g.p)) return tps::launch_any<double>(g, true, h->stream, err);
    if (tc::supported(g.kind, g.m, g.p)) return tc::launch_f64(g, true, h->stream, err);
    if (sw::supported(g.kind, g.m, g.p)) return sw::launch_f64(g, true, h->stream, err);
    return gen_backward_f64(g, h->stream, err);
}
static int run_backward(bkf_handle* h, const KArgs<float>& g, const char** err) {
    KernelTimer kt(h, BKF_KERNEL_BACKWARD);
    h->launches += 1;
    if (g.kind == BKF_CHOLESKY) return sq::launch_backward<float>(g, h->stream, err);
    if ((!g.tv_mask || tc::tv_supported(g.tv_mask)) && tps::supported(g.kind, g.m, g.p))
        return tps::launch_any<float>(g, true, h->stream, err);
    if (!g.tv_mask && swf::supported(g.kind, g.m, g.p)) return swf::launch_f32(g, true, h->stream, err);
    return gen::launch_backward<float>(g, h->stream, err);
}

// Smoother dispatch: the DMMA / sub-warp kernels when available, with a per-series fallback to the eigen-decomposition
// pinv of the generic kernel for the series they flag as ill-conditioned.
static int run_smoother(bkf_handle* h, const KArgs<double>& g, const char** err) {
    const double jitter0 = 1e-8;  // JITTER_DEFAULT (float64), kalman_smoother.py:117
    const bool use_tc = tc::smoother_supported(g.m);
    const unsigned sm_tv = g.tv_mask & ((1u << BKF_PARAM_T) | (1u << BKF_PARAM_R) | (1u << BKF_PARAM_Q));  // what it reads
    if (!sm_tv && (use_tc || sw::smoother_supported(g.m))) {
        // The fast kernels invert P_hat directly and list the series where that is not safe (pivot <= 0 or spanning
        // > 1e13); the generic kernel then redoes exactly those with the eigen-decomposition pinv.  The list lives on
        // the device and the second launch strides over it (it returns at once when the list is empty): no host round
        // trip, the flagged series are reported through status (BKF_STATUS_SMOOTHER_ILL).
        int rc = BKF_OK;
        int* ill = (int*)ensure(h, S_FLAG, ((size_t)g.B + 1) * sizeof(int), &rc);
        if (!ill) return rc;
        cudaMemsetAsync(ill, 0, sizeof(int), h->stream);
        {
            KernelTimer kt(h, BKF_KERNEL_SMOOTHER);
            h->launches += 1;
            rc = use_tc ? tc::launch_smoother_f64(g, jitter0, ill, h->stream, err)
                        : sw::launch_smoother_f64(g, jitter0, ill, h->stream, err);
        }
        if (rc != BKF_OK) return rc;
        h->path = use_tc ? "dmma_f64" : "subwarp_f64";
        h->launches += 1;
        return gen_smoother_f64(g, ill + 1, ill, h->stream, err);
    }
    KernelTimer kt(h, BKF_KERNEL_SMOOTHER);
    h->launches += 1;
    return gen_smoother_f64(g, nullptr, nullptr, h->stream, err);
}
static int run_smoother(bkf_handle* h, const KArgs<float>& g, const char** err) {
    KernelTimer kt(h, BKF_KERNEL_SMOOTHER);
    h->launches += 1;
    return gen::launch_smoother<float>(g, nullptr, nullptr, h->stream, err);
}

// elements of parameter `which` for ONE series: the matrix itself, times n when it carries a time axis
static size_t param_elems(const bkf_problem* P, int which) {
    size_t m = P->m, p = P->p, r = P->r, e = 0;
    switch (which) {
        case BKF_PARAM_A0: case BKF_PARAM_C: e = m; break;
        case BKF_PARAM_P0: case BKF_PARAM_T: e = m * m; break;
        case BKF_PARAM_D: e = p; break;
        case BKF_PARAM_Z: e = p * m; break;
        case BKF_PARAM_R: e = m * r; break;
        case BKF_PARAM_H: e = p * p; break;
        case BKF_PARAM_Q: e = r * r; break;
    }
    return ((P->tv_mask >> which) & 1u) ? e * (size_t)P->n : e;
}

static int validate(bkf_handle* h, const bkf_problem* P, bool need_obs) {
    if (!h) return BKF_ERR_ARG;
    if (!P) return fail(h, BKF_ERR_ARG, "null problem");
    if (P->kind < BKF_STANDARD || P->kind > BKF_CHOLESKY) return fail(h, BKF_ERR_ARG, "unknown filter kind");
    if (P->dtype != BKF_F64 && P->dtype != BKF_F32) return fail(h, BKF_ERR_ARG, "unknown dtype");
    if (P->mem != BKF_MEM_HOST && P->mem != BKF_MEM_DEVICE) return fail(h, BKF_ERR_ARG, "unknown memory space");
    if (P->B < 0 || P->n < 1 || P->m < 1 || P->r < 1 || (need_obs && P->p < 1))
        return fail(h, BKF_ERR_ARG, "non-positive dimension");
    if (P->tv_mask & ((1u << BKF_PARAM_A0) | (1u << BKF_PARAM_P0)))
        return fail(h, BKF_ERR_ARG, "a0 and P0 are never time-varying (utils/constants.py NEVER_TIME_VARYING)");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    return BKF_OK;
}

// Stage an input: returns the device pointer to use (the caller's own pointer in device mode).
template <typename R>
static const R* stage_in(bkf_handle* h, const bkf_problem* P, int slot, const void* src, size_t elems, int* rc) {
    if (*rc != BKF_OK || !src) return reinterpret_cast<const R*>(src);
    if (P->mem == BKF_MEM_DEVICE) return reinterpret_cast<const R*>(src);
    void* d = ensure(h, slot, elems * sizeof(R), rc);
    if (!d) return nullptr;
    cudaError_t e = cudaMemcpyAsync(d, src, elems * sizeof(R), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { *rc = cuda_fail(h, e, "H2D copy"); return nullptr; }
    return reinterpret_cast<const R*>(d);
}

// Prepare an output: device pointer the kernels write to (NULL if the caller passed NULL).
template <typename R>
static R* stage_out(bkf_handle* h, const bkf_problem* P, int slot, void* dst, size_t elems, int* rc) {
    if (*rc != BKF_OK || !dst) return nullptr;
    if (P->mem == BKF_MEM_DEVICE) return reinterpret_cast<R*>(dst);
    return reinterpret_cast<R*>(ensure(h, slot, elems * sizeof(R), rc));
}

template <typename R>
static void copy_back(bkf_handle* h, const bkf_problem* P, void* dst, const R* dev, size_t elems, int* rc) {
    if (*rc != BKF_OK || !dst || P->mem == BKF_MEM_DEVICE) return;
    cudaError_t e = cudaMemcpyAsync(dst, dev, elems * sizeof(R), cudaMemcpyDeviceToHost, h->stream);
    if (e != cudaSuccess) *rc = cuda_fail(h, e, "D2H copy");
}

static int finish(bkf_handle* h, const bkf_problem* P, int rc) {
    if (rc != BKF_OK) return rc;
    if (P->mem == BKF_MEM_HOST) {
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return cuda_fail(h, e, "kernel execution");
    } else {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(h, e, "kernel launch");
    }
    return BKF_OK;
}

template <typename R>
static void fill_inputs(bkf_handle* h, const bkf_problem* P, KArgs<R>& g, const void* y, const void* const* prm,
                        int* rc) {
    std::memset(&g, 0, sizeof(g));
    g.kind = P->kind; g.B = P->B; g.n = P->n; g.m = P->m; g.p = P->p; g.r = P->r; g.y_shared = P->y_shared;
    g.shared_mask = P->shared_mask;
    g.tv_mask = P->tv_mask;
    g.jitter = (R)P->cov_jitter;
    g.fill = (R)P->missing_fill;
    g.rec_tiles = P->dtype == BKF_F64 && sqb::tile_record(P->kind, P->m, P->p, P->r, P->tv_mask);
    size_t B = P->B;
    if (y) g.y = stage_in<R>(h, P, S_Y, y, (P->y_shared ? 1 : B) * (size_t)P->n * P->p, rc);
    const R** dst[BKF_NPARAM] = {&g.a0, &g.P0, &g.c, &g.d, &g.T, &g.Z, &g.Rm, &g.H, &g.Q};
    for (int i = 0; i < BKF_NPARAM; ++i) {
        if (!prm[i]) continue;
        size_t nb = ((P->shared_mask >> i) & 1u) ? 1 : B;
        *dst[i] = stage_in<R>(h, P, S_A0 + i, prm[i], nb * param_elems(P, i), rc);
    }
}

static bool any_null(const void* const* p, int n) {
    for (int i = 0; i < n; ++i)
        if (!p[i]) return true;
    return false;
}

// ----------------------------------------------------------------------------------------------------
template <typename R>
static int forward_impl(bkf_handle* h, const bkf_problem* P, const void* y, const void* const* prm, void* fa,
                        void* pa, void* om, void* fP, void* pP, void* oc, void* ll, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    fill_inputs<R>(h, P, g, y, prm, &rc);
    const size_t B = P->B, n = P->n, m = P->m, p = P->p;
    const size_t po = P->kind == BKF_UNIVARIATE ? 1 : p;
    g.filt_a = stage_out<R>(h, P, S_FA, fa, B * n * m, &rc);
    g.pred_a = stage_out<R>(h, P, S_PA, pa, B * n * m, &rc);
    g.obs_mu = stage_out<R>(h, P, S_OM, om, B * n * po, &rc);
    g.filt_P = stage_out<R>(h, P, S_FP, fP, B * n * m * m, &rc);
    g.pred_P = stage_out<R>(h, P, S_PP, pP, B * n * m * m, &rc);
    g.obs_cov = stage_out<R>(h, P, S_OC, oc, B * n * po * po, &rc);
    g.ll = stage_out<R>(h, P, S_LL, ll, B * n, &rc);
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    rc = run_forward(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, fa, g.filt_a, B * n * m, &rc);
    copy_back<R>(h, P, pa, g.pred_a, B * n * m, &rc);
    copy_back<R>(h, P, om, g.obs_mu, B * n * po, &rc);
    copy_back<R>(h, P, fP, g.filt_P, B * n * m * m, &rc);
    copy_back<R>(h, P, pP, g.pred_P, B * n * m * m, &rc);
    copy_back<R>(h, P, oc, g.obs_cov, B * n * po * po, &rc);
    copy_back<R>(h, P, ll, g.ll, B * n, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_filter_forward(bkf_handle* h, const bkf_problem* P, const void* y, const void* a0, const void* P0,
                                  const void* c, const void* d, const void* T, const void* Z, const void* R,
                                  const void* H, const void* Q, void* fa, void* pa, void* om, void* fP, void* pP,
                                  void* oc, void* ll, int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    const void* prm[BKF_NPARAM] = {a0, P0, c, d, T, Z, R, H, Q};
    if (!y || any_null(prm, BKF_NPARAM)) return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64 ? forward_impl<double>(h, P, y, prm, fa, pa, om, fP, pP, oc, ll, status)
                               : forward_impl<float>(h, P, y, prm, fa, pa, om, fP, pP, oc, ll, status);
}

static size_t traj_record_elems(const bkf_problem* P) {
    // the trajectory layout is private to a kernel family: dense [a | P] records, DMMA fragment records, or the
    // square-root family's [a | Pc | triu(R_update) | triu(R_predict)]
    const bool f64 = P->dtype == BKF_F64;
    if (P->kind == BKF_CHOLESKY) return sq::traj_record_elems_host(P->m, P->p);
    const bool frag = f64 && tc::tv_supported(P->tv_mask) && !tps::supported(P->kind, P->m, P->p) &&
                      tc::supported(P->kind, P->m, P->p);
    return frag ? tc::traj_record_elems() : (size_t)P->m + (size_t)P->m * P->m;
}

// Elements of the whole trajectory workspace.  The square-root family appends one record per series: scratch for the
// adjoint's add-only accumulators (bkf_sqrt_warp.cuh).
static size_t traj_elems(const bkf_problem* P) {
    const size_t B = P->B, n = P->n;
    if (P->dtype == BKF_F64 && sqb::tile_record(P->kind, P->m, P->p, P->r, P->tv_mask)) {
        size_t rec, scr;
        sqb::sizes(P->m, P->p, P->r, &rec, &scr);
        return B * n * rec + B * scr;
    }
    return B * (n + (P->kind == BKF_CHOLESKY ? 1 : 0)) * traj_record_elems(P);
}

extern "C" int bkf_logp_grad_workspace_bytes(const bkf_problem* P, size_t* bytes) {
    if (!P || !bytes) return BKF_ERR_ARG;
    const size_t es = P->dtype == BKF_F64 ? 8 : 4, B = P->B, n = P->n;
    size_t tot = traj_elems(P) * es;
    if (P->mem == BKF_MEM_HOST) {
        tot += (P->y_shared ? 1 : B) * n * (size_t)P->p * es + B * n * es /* cotangent */ + B * (es + sizeof(int));
        for (int i = 0; i < BKF_NPARAM; ++i)
            tot += ((((P->shared_mask >> i) & 1u) ? 1 : B) + B) * param_elems(P, i) * es;  // input + gradient
    }
    *bytes = tot;
    return BKF_OK;
}

extern "C" int bkf_device_memory(bkf_handle* h, size_t* free_bytes, size_t* total_bytes) {
    if (!h) return BKF_ERR_ARG;
    cudaError_t e = cudaSetDevice(h->device);
    size_t f = 0, t = 0;
    if (e == cudaSuccess) e = cudaMemGetInfo(&f, &t);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMemGetInfo");
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return BKF_OK;
}

// ----------------------------------------------------------------------------------------------------
template <typename R>
static int logp_grad_impl(bkf_handle* h, const bkf_problem* P, const void* y, const void* const* prm,
                          const void* cot, void* logp, void* const* grads, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    fill_inputs<R>(h, P, g, y, prm, &rc);
    const size_t B = P->B, n = P->n;
    bool want_grad = false;
    for (int i = 0; i < BKF_NPARAM; ++i) want_grad = want_grad || grads[i];
    if (cot) g.ll_cot = stage_in<R>(h, P, S_COT, cot, B * n, &rc);
    g.logp = stage_out<R>(h, P, S_LOGP, logp, B, &rc);
    R** gdst[BKF_NPARAM] = {&g.g_a0, &g.g_P0, &g.g_c, &g.g_d, &g.g_T, &g.g_Z, &g.g_R, &g.g_H, &g.g_Q};
    for (int i = 0; i < BKF_NPARAM; ++i)
        *gdst[i] = stage_out<R>(h, P, S_GA0 + i, grads[i], B * param_elems(P, i), &rc);
    if (want_grad) {
        g.traj = (R*)ensure(h, S_TRAJ, traj_elems(P) * sizeof(R), &rc);
    }
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    rc = run_forward(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    if (want_grad) {
        rc = run_backward(h, g, &err);
        if (rc != BKF_OK) return fail(h, rc, err);
    }
    copy_back<R>(h, P, logp, g.logp, B, &rc);
    for (int i = 0; i < BKF_NPARAM; ++i) copy_back<R>(h, P, grads[i], *gdst[i], B * param_elems(P, i), &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_filter_logp_grad(bkf_handle* h, const bkf_problem* P, const void* y, const void* a0,
                                    const void* P0, const void* c, const void* d, const void* T, const void* Z,
                                    const void* R, const void* H, const void* Q, const void* ll_cotangent,
                                    void* logp, void* g_a0, void* g_P0, void* g_c, void* g_d, void* g_T, void* g_Z,
                                    void* g_R, void* g_H, void* g_Q, int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    const void* prm[BKF_NPARAM] = {a0, P0, c, d, T, Z, R, H, Q};
    if (!y || any_null(prm, BKF_NPARAM)) return fail(h, BKF_ERR_ARG, "null input pointer");
    void* grads[BKF_NPARAM] = {g_a0, g_P0, g_c, g_d, g_T, g_Z, g_R, g_H, g_Q};
    return P->dtype == BKF_F64 ? logp_grad_impl<double>(h, P, y, prm, ll_cotangent, logp, grads, status)
                               : logp_grad_impl<float>(h, P, y, prm, ll_cotangent, logp, grads, status);
}

// ----------------------------------------------------------------------------------------------------
// forward sweep now, adjoint sweep later (bkf.h: bkf_filter_forward_saved / bkf_filter_backward_saved)
static KArgs<double>& saved_args(bkf_handle* h, double) { return h->saved.gd; }
static KArgs<float>& saved_args(bkf_handle* h, float) { return h->saved.gf; }

template <typename R>
static int forward_saved_impl(bkf_handle* h, const bkf_problem* P, const void* y, const void* const* prm, void* ll,
                              void* logp, int* status, unsigned long long* ticket) {
    int rc = BKF_OK;
    h->saved.valid = false;
    KArgs<R> g;
    fill_inputs<R>(h, P, g, y, prm, &rc);
    const size_t B = P->B, n = P->n;
    g.ll = stage_out<R>(h, P, S_LL, ll, B * n, &rc);
    g.logp = stage_out<R>(h, P, S_LOGP, logp, B, &rc);
    g.traj = (R*)ensure(h, S_TRAJ, traj_elems(P) * sizeof(R), &rc);
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    if (rc != BKF_OK) return rc;
    if (B > 0) {
        const char* err = "";
        rc = run_forward(h, g, &err);
        if (rc != BKF_OK) return fail(h, rc, err);
        copy_back<R>(h, P, ll, g.ll, B * n, &rc);
        copy_back<R>(h, P, logp, g.logp, B, &rc);
        if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
            cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
            if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
        }
    }
    rc = finish(h, P, rc);
    if (rc != BKF_OK) return rc;
    g.ll = nullptr;  // (the adjoint sweep reads the record, not the outputs)
    g.logp = nullptr;
    g.status = nullptr;
    h->saved.P = *P;
    saved_args(h, R(0)) = g;
    h->saved.ticket = ++h->tickets;
    h->saved.gen = h->gen;
    h->saved.valid = true;
    *ticket = h->saved.ticket;
    return BKF_OK;
}

template <typename R>
static int backward_saved_impl(bkf_handle* h, const void* cot, void* const* grads) {
    const bkf_problem* P = &h->saved.P;
    KArgs<R> g = saved_args(h, R(0));
    h->saved.valid = false;  // one ticket, one adjoint
    int rc = BKF_OK;
    const size_t B = P->B, n = P->n;
    if (cot) g.ll_cot = stage_in<R>(h, P, S_COT, cot, B * n, &rc);
    R** gdst[BKF_NPARAM] = {&g.g_a0, &g.g_P0, &g.g_c, &g.g_d, &g.g_T, &g.g_Z, &g.g_R, &g.g_H, &g.g_Q};
    for (int i = 0; i < BKF_NPARAM; ++i)
        *gdst[i] = stage_out<R>(h, P, S_GA0 + i, grads[i], B * param_elems(P, i), &rc);
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    rc = run_backward(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    for (int i = 0; i < BKF_NPARAM; ++i) copy_back<R>(h, P, grads[i], *gdst[i], B * param_elems(P, i), &rc);
    return finish(h, P, rc);
}

extern "C" int bkf_filter_forward_saved(bkf_handle* h, const bkf_problem* P, const void* y, const void* a0,
                                        const void* P0, const void* c, const void* d, const void* T, const void* Z,
                                        const void* R, const void* H, const void* Q, void* ll, void* logp, int* status,
                                        unsigned long long* ticket) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    const void* prm[BKF_NPARAM] = {a0, P0, c, d, T, Z, R, H, Q};
    if (!y || !ticket || any_null(prm, BKF_NPARAM)) return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64 ? forward_saved_impl<double>(h, P, y, prm, ll, logp, status, ticket)
                               : forward_saved_impl<float>(h, P, y, prm, ll, logp, status, ticket);
}

extern "C" int bkf_filter_backward_saved(bkf_handle* h, unsigned long long ticket, const void* cot, void* g_a0,
                                         void* g_P0, void* g_c, void* g_d, void* g_T, void* g_Z, void* g_R, void* g_H,
                                         void* g_Q) {
    if (!h) return BKF_ERR_ARG;
    if (!h->saved.valid || h->saved.ticket != ticket || h->saved.gen != h->gen)
        return fail(h, BKF_ERR_STALE, "the saved forward sweep is gone (another call used the workspace)");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    void* grads[BKF_NPARAM] = {g_a0, g_P0, g_c, g_d, g_T, g_Z, g_R, g_H, g_Q};
    return h->saved.P.dtype == BKF_F64 ? backward_saved_impl<double>(h, cot, grads) : backward_saved_impl<float>(h, cot, grads);
}

// ----------------------------------------------------------------------------------------------------
// compact parameterisation (bkf.h: bkf_filter_logp_grad_compact)
template <typename R>
__global__ void expand_kernel(R* __restrict__ out, const R* __restrict__ base, size_t elems, size_t B) {
    const size_t total = elems * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = base[i % elems];
}
struct MapEntry { int which, index; };
template <typename R>
struct ScatterArgs { R* mat[BKF_NPARAM]; size_t elems[BKF_NPARAM]; };
template <typename R>
__global__ void scatter_kernel(ScatterArgs<R> a, const MapEntry* __restrict__ map, int nnz, const R* __restrict__ u,
                               size_t B) {
    const size_t total = (size_t)nnz * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / nnz;
        const MapEntry e = map[i - b * nnz];
        a.mat[e.which][b * a.elems[e.which] + e.index] = u[i];
    }
}
template <typename R>
__global__ void gather_kernel(R* __restrict__ gu, ScatterArgs<R> a, const MapEntry* __restrict__ map, int nnz, size_t B) {
    const size_t total = (size_t)nnz * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / nnz;
        const MapEntry e = map[i - b * nnz];
        gu[i] = a.mat[e.which][b * a.elems[e.which] + e.index];
    }
}

// theta -> entries program (bkf.h: bkf_filter_logp_grad_theta).  Entry k of the compact vector is
//     u_k(theta) = sum over its terms  coef * prod_{f < nfac} fn_f(theta[idx_f]),   nfac <= 3.
// Device layout of the table (ints, then doubles): start[nnz + 1] | nfac[nterms] | idx[3 nterms] | fn[3 nterms] ; coef[nterms].
struct ThetaProg {
    int nterms, ntheta;
    const int *start, *nfac, *idx, *fn;  // device
    const double* coef;                  // device
    const void* theta;                   // caller's theta [B, ntheta] (prob->mem)
    void* g_theta;                       // caller's output [B, ntheta] (nullable)
};
__device__ __forceinline__ double theta_fn(int fn, double x, double* dx) {
    switch (fn) {
        case BKF_FN_COS: *dx = -sin(x); return cos(x);
        case BKF_FN_SIN: *dx = cos(x); return sin(x);
        case BKF_FN_EXP: { const double e = exp(x); *dx = e; return e; }
        case BKF_FN_ONE_MINUS: *dx = -1.0; return 1.0 - x;
        case BKF_FN_SQUARE: *dx = 2.0 * x; return x * x;
        default: *dx = 1.0; return x;
    }
}
template <typename R>
__global__ void theta_eval_kernel(R* __restrict__ u, const R* __restrict__ theta, ThetaProg pg, int nnz, size_t B) {
    const size_t total = (size_t)nnz * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / nnz;
        const int k = (int)(i - b * nnz);
        const R* th = theta + b * pg.ntheta;
        double acc = 0.0;
        for (int q = pg.start[k]; q < pg.start[k + 1]; ++q) {
            double v = pg.coef[q], dx;
            for (int f = 0; f < pg.nfac[q]; ++f) v *= theta_fn(pg.fn[3 * q + f], (double)th[pg.idx[3 * q + f]], &dx);
            acc += v;
        }
        u[i] = (R)acc;
    }
}
// chain rule: g_theta[b, i] = sum_k g_u[b, k] du_k / dtheta_i; one thread per draw walks the table in order (deterministic)
template <typename R>
__global__ void theta_chain_kernel(R* __restrict__ gth, const R* __restrict__ gu, const R* __restrict__ theta, ThetaProg pg,
                                   int nnz, size_t B) {
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < B; b += (size_t)gridDim.x * blockDim.x) {
        const R* th = theta + b * pg.ntheta;
        R* g = gth + b * pg.ntheta;
        for (int i = 0; i < pg.ntheta; ++i) g[i] = R(0);
        for (int k = 0; k < nnz; ++k) {
            const double gk = (double)gu[b * nnz + k];
            for (int q = pg.start[k]; q < pg.start[k + 1]; ++q) {
                double val[3], der[3];
                const int nf = pg.nfac[q];
                for (int f = 0; f < nf; ++f) val[f] = theta_fn(pg.fn[3 * q + f], (double)th[pg.idx[3 * q + f]], &der[f]);
                for (int f = 0; f < nf; ++f) {
                    double v = gk * pg.coef[q] * der[f];
                    for (int f2 = 0; f2 < nf; ++f2)
                        if (f2 != f) v *= val[f2];
                    g[pg.idx[3 * q + f]] += (R)v;
                }
            }
        }
    }
}

template <typename R>
static int compact_impl(bkf_handle* h, const bkf_problem* P, const void* y, const void* const* base, int nnz,
                        const int* which, const int* index, const void* u, const void* cot, void* logp, void* g_u,
                        int* status, const ThetaProg* prog = nullptr) {
    int rc = BKF_OK;
    const size_t B = P->B, n = P->n;
    const bool host = P->mem == BKF_MEM_HOST;
    bool mapped[BKF_NPARAM] = {false};
    std::vector<MapEntry> hmap(nnz);
    for (int k = 0; k < nnz; ++k) {
        if (which[k] < 0 || which[k] >= BKF_NPARAM || index[k] < 0 || (size_t)index[k] >= param_elems(P, which[k]))
            return fail(h, BKF_ERR_ARG, "compact map entry out of range");
        for (int q = 0; q < k; ++q)
            if (which[q] == which[k] && index[q] == index[k]) return fail(h, BKF_ERR_ARG, "compact map entries must be unique");
        mapped[which[k]] = true;
        hmap[k] = {which[k], index[k]};
    }
    // ---- inputs: bases are shared; mapped parameters get a per-draw expansion in the staging slots ----
    KArgs<R> g;
    std::memset(&g, 0, sizeof(g));
    g.kind = P->kind; g.B = P->B; g.n = P->n; g.m = P->m; g.p = P->p; g.r = P->r; g.y_shared = P->y_shared;
    g.jitter = (R)P->cov_jitter; g.fill = (R)P->missing_fill;
    bkf_problem Pd = *P;  // what the kernels see: device memory, unmapped parameters shared
    Pd.mem = BKF_MEM_DEVICE;
    Pd.shared_mask = 0;
    Pd.tv_mask = 0;
    g.y = stage_in<R>(h, P, S_Y, y, (P->y_shared ? 1 : B) * n * (size_t)P->p, &rc);
    const R* du = nullptr;
    const R* dtheta = nullptr;
    if (prog) {  // u is evaluated on the device from theta
        dtheta = stage_in<R>(h, P, S_THETA, prog->theta, B * (size_t)prog->ntheta, &rc);
        du = (const R*)ensure(h, S_U, (B * (size_t)nnz + 1) * sizeof(R), &rc);
    } else {
        du = stage_in<R>(h, P, S_U, u, B * (size_t)nnz, &rc);
    }
    if (cot) g.ll_cot = stage_in<R>(h, P, S_COT, cot, B * n, &rc);
    size_t base_off[BKF_NPARAM + 1] = {0};
    for (int i = 0; i < BKF_NPARAM; ++i) base_off[i + 1] = base_off[i] + param_elems(P, i);
    R* dbase = nullptr;
    if (host) {  // one staging buffer for the nine small base matrices
        dbase = (R*)ensure(h, S_BASE, base_off[BKF_NPARAM] * sizeof(R), &rc);
        if (!dbase) return rc;
        for (int i = 0; i < BKF_NPARAM; ++i)
            cudaMemcpyAsync(dbase + base_off[i], base[i], param_elems(P, i) * sizeof(R), cudaMemcpyHostToDevice, h->stream);
    }
    MapEntry* dmap = (MapEntry*)ensure(h, S_MAP, (size_t)(nnz > 0 ? nnz : 1) * sizeof(MapEntry), &rc);
    if (rc != BKF_OK) return rc;
    if (nnz) cudaMemcpyAsync(dmap, hmap.data(), (size_t)nnz * sizeof(MapEntry), cudaMemcpyHostToDevice, h->stream);
    const R** in[BKF_NPARAM] = {&g.a0, &g.P0, &g.c, &g.d, &g.T, &g.Z, &g.Rm, &g.H, &g.Q};
    R** gd[BKF_NPARAM] = {&g.g_a0, &g.g_P0, &g.g_c, &g.g_d, &g.g_T, &g.g_Z, &g.g_R, &g.g_H, &g.g_Q};
    ScatterArgs<R> sa, ga;
    for (int i = 0; i < BKF_NPARAM; ++i) {
        const size_t e = param_elems(P, i);
        const R* bsrc = host ? dbase + base_off[i] : (const R*)base[i];
        sa.elems[i] = ga.elems[i] = e;
        sa.mat[i] = ga.mat[i] = nullptr;
        if (mapped[i]) {
            R* ex = (R*)ensure(h, S_A0 + i, B * e * sizeof(R), &rc);
            R* gx = (R*)ensure(h, S_GA0 + i, B * e * sizeof(R), &rc);
            if (rc != BKF_OK) return rc;
            if (B) expand_kernel<R><<<(unsigned)std::min<size_t>((B * e + 255) / 256, 4096), 256, 0, h->stream>>>(ex, bsrc, e, B);
            sa.mat[i] = ex;
            ga.mat[i] = gx;
            *in[i] = ex;
            *gd[i] = gx;
        } else {
            Pd.shared_mask |= 1u << i;
            *in[i] = bsrc;
        }
    }
    g.shared_mask = Pd.shared_mask;
    g.logp = stage_out<R>(h, P, S_LOGP, logp, B, &rc);
    R* dgu = prog ? (prog->g_theta ? (R*)ensure(h, S_GU, (B * (size_t)nnz + 1) * sizeof(R), &rc) : nullptr)
                  : stage_out<R>(h, P, S_GU, g_u, B * (size_t)nnz, &rc);
    R* dgth = prog ? stage_out<R>(h, P, S_GTHETA, prog->g_theta, B * (size_t)prog->ntheta, &rc) : nullptr;
    int* dstatus = nullptr;
    if (status) dstatus = host ? (int*)ensure(h, S_STATUS, B * sizeof(int), &rc) : status;
    g.status = dstatus;
    const bool want_grad = (prog ? prog->g_theta != nullptr : g_u != nullptr) && nnz > 0;
    if (want_grad) g.traj = (R*)ensure(h, S_TRAJ, traj_elems(&Pd) * sizeof(R), &rc);
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const unsigned sgrid = (unsigned)std::min<size_t>((B * (size_t)(nnz > 0 ? nnz : 1) + 255) / 256, 4096);
    if (prog && nnz) {
        theta_eval_kernel<R><<<sgrid, 256, 0, h->stream>>>(const_cast<R*>(du), dtheta, *prog, nnz, B);
        h->launches += 1;
    }
    if (nnz) scatter_kernel<R><<<sgrid, 256, 0, h->stream>>>(sa, dmap, nnz, du, B);
    h->launches += 1 + (nnz ? 1 : 0);
    const char* err = "";
    rc = run_forward(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    if (want_grad) {
        rc = run_backward(h, g, &err);
        if (rc != BKF_OK) return fail(h, rc, err);
        gather_kernel<R><<<sgrid, 256, 0, h->stream>>>(dgu, ga, dmap, nnz, B);
        h->launches += 1;
        if (prog) {
            theta_chain_kernel<R><<<(unsigned)std::min<size_t>((B + 127) / 128, 4096), 128, 0, h->stream>>>(dgth, dgu, dtheta, *prog, nnz, B);
            h->launches += 1;
        }
    }
    copy_back<R>(h, P, logp, g.logp, B, &rc);
    if (want_grad && !prog) copy_back<R>(h, P, g_u, dgu, B * (size_t)nnz, &rc);
    if (want_grad && prog) copy_back<R>(h, P, prog->g_theta, dgth, B * (size_t)prog->ntheta, &rc);
    if (status && host && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_filter_logp_grad_compact(bkf_handle* h, const bkf_problem* P, const void* y,
                                            const void* const base[9], int nnz, const int* which, const int* index,
                                            const void* u, const void* ll_cotangent, void* logp, void* g_u,
                                            int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    if (P->tv_mask) return fail(h, BKF_ERR_UNSUPPORTED, "the compact parameterisation does not take time-varying inputs");
    if (!y || !base || any_null(base, BKF_NPARAM) || nnz < 0 || (nnz > 0 && (!which || !index || !u)))
        return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64
               ? compact_impl<double>(h, P, y, base, nnz, which, index, u, ll_cotangent, logp, g_u, status)
               : compact_impl<float>(h, P, y, base, nnz, which, index, u, ll_cotangent, logp, g_u, status);
}

extern "C" int bkf_filter_logp_grad_theta(bkf_handle* h, const bkf_problem* P, const void* y, const void* const base[9],
                                          int nnz, const int* which, const int* index, int nterms, const int* term_entry,
                                          const double* term_coef, const int* term_nfac, const int* term_theta,
                                          const int* term_fn, int ntheta, const void* theta, const void* ll_cotangent,
                                          void* logp, void* g_theta, int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    if (P->tv_mask) return fail(h, BKF_ERR_UNSUPPORTED, "the theta parameterisation does not take time-varying inputs");
    if (!y || !base || any_null(base, BKF_NPARAM) || nnz < 0 || nterms < 0 || ntheta < 0 ||
        (nnz > 0 && (!which || !index)) || (nterms > 0 && (!term_entry || !term_coef || !term_nfac || !term_theta || !term_fn)) ||
        (ntheta > 0 && !theta))
        return fail(h, BKF_ERR_ARG, "null input pointer");
    // table: terms grouped by entry (CSR), validated on the host, shipped once per call (a few hundred bytes)
    std::vector<int> start(nnz + 1, 0), order(nterms);
    for (int q = 0; q < nterms; ++q) {
        if (term_entry[q] < 0 || term_entry[q] >= nnz || term_nfac[q] < 0 || term_nfac[q] > 3)
            return fail(h, BKF_ERR_ARG, "theta program: term entry / factor count out of range");
        for (int f = 0; f < term_nfac[q]; ++f)
            if (term_theta[3 * q + f] < 0 || term_theta[3 * q + f] >= ntheta || term_fn[3 * q + f] < 0 ||
                term_fn[3 * q + f] > BKF_FN_SQUARE)
                return fail(h, BKF_ERR_ARG, "theta program: parameter index / function code out of range");
        start[term_entry[q] + 1] += 1;
    }
    for (int k = 0; k < nnz; ++k) start[k + 1] += start[k];
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int q = 0; q < nterms; ++q) order[fill[term_entry[q]]++] = q;  // stable: the caller's order within an entry
    }
    const size_t nint = (size_t)(nnz + 1) + 7 * (size_t)nterms, nint_pad = (nint + 1) & ~(size_t)1;
    std::vector<int> ints(nint_pad, 0);
    std::vector<double> coef(nterms > 0 ? nterms : 1, 0.0);
    int* pstart = ints.data();
    int* pnfac = pstart + nnz + 1;
    int* pidx = pnfac + nterms;
    int* pfn = pidx + 3 * (size_t)nterms;
    for (int k = 0; k <= nnz; ++k) pstart[k] = start[k];
    for (int j = 0; j < nterms; ++j) {
        const int q = order[j];
        pnfac[j] = term_nfac[q];
        coef[j] = term_coef[q];
        for (int f = 0; f < 3; ++f) {
            pidx[3 * j + f] = f < term_nfac[q] ? term_theta[3 * q + f] : 0;
            pfn[3 * j + f] = f < term_nfac[q] ? term_fn[3 * q + f] : 0;
        }
    }
    char* dtab = (char*)ensure(h, S_PROG, nint_pad * sizeof(int) + coef.size() * sizeof(double), &rc);
    if (!dtab) return rc;
    cudaError_t e = cudaMemcpyAsync(dtab, ints.data(), nint_pad * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dtab + nint_pad * sizeof(int), coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // (the host vectors die with this frame)
    if (e != cudaSuccess) return cuda_fail(h, e, "theta program upload");
    ThetaProg pg;
    pg.nterms = nterms; pg.ntheta = ntheta;
    pg.start = (const int*)dtab; pg.nfac = pg.start + nnz + 1; pg.idx = pg.nfac + nterms; pg.fn = pg.idx + 3 * (size_t)nterms;
    pg.coef = (const double*)(dtab + nint_pad * sizeof(int));
    pg.theta = theta; pg.g_theta = g_theta;
    return P->dtype == BKF_F64
               ? compact_impl<double>(h, P, y, base, nnz, which, index, nullptr, ll_cotangent, logp, nullptr, status, &pg)
               : compact_impl<float>(h, P, y, base, nnz, which, index, nullptr, ll_cotangent, logp, nullptr, status, &pg);
}

// ----------------------------------------------------------------------------------------------------
// stationary initialisation (bkf.h: bkf_stationary_init, bkf_stationary_init_grad)
template <typename R>
static int stationary_impl(bkf_handle* h, const bkf_problem* P, const void* c, const void* T, const void* Rm,
                           const void* Q, void* a0, void* P0, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    const void* prm[BKF_NPARAM] = {nullptr, nullptr, c, nullptr, T, nullptr, Rm, nullptr, Q};
    bkf_problem P2 = *P;
    P2.p = P2.p > 0 ? P2.p : 1;
    P2.n = P2.n > 0 ? P2.n : 1;
    fill_inputs<R>(h, &P2, g, nullptr, prm, &rc);
    const size_t B = P->B, m = P->m;
    g.sm_a = stage_out<R>(h, P, S_SMA, a0, B * m, &rc);
    g.sm_P = stage_out<R>(h, P, S_SMP, P0, B * m * m, &rc);
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    if (rc != BKF_OK) return rc;
    if (!g.sm_a || !g.sm_P) return fail(h, BKF_ERR_ARG, "null output pointer");
    if (B == 0) return BKF_OK;
    const char* err = "";
    h->launches += 1;
    h->path = "stationary_warp";
    rc = stat::launch<R>(g, false, h->stream, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, a0, g.sm_a, B * m, &rc);
    copy_back<R>(h, P, P0, g.sm_P, B * m * m, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

template <typename R>
static int stationary_grad_impl(bkf_handle* h, const bkf_problem* P, const void* T, const void* Rm, const void* Q,
                                const void* a0, const void* P0, const void* a0b, const void* P0b, void* g_c, void* g_T,
                                void* g_R, void* g_Q) {
    int rc = BKF_OK;
    KArgs<R> g;
    const void* prm[BKF_NPARAM] = {nullptr, nullptr, nullptr, nullptr, T, nullptr, Rm, nullptr, Q};
    bkf_problem P2 = *P;
    P2.p = P2.p > 0 ? P2.p : 1;
    P2.n = P2.n > 0 ? P2.n : 1;
    fill_inputs<R>(h, &P2, g, nullptr, prm, &rc);
    const size_t B = P->B, m = P->m, r = P->r;
    g.sm_filt_a = stage_in<R>(h, P, S_WFA, a0, B * m, &rc);
    g.sm_filt_P = stage_in<R>(h, P, S_WFP, P0, B * m * m, &rc);
    g.a0 = stage_in<R>(h, P, S_A0, a0b, B * m, &rc);   // cotangents travel in the a0 / P0 input slots
    g.P0 = stage_in<R>(h, P, S_P0, P0b, B * m * m, &rc);
    g.g_c = stage_out<R>(h, P, S_GC, g_c, B * m, &rc);
    g.g_T = stage_out<R>(h, P, S_GT, g_T, B * m * m, &rc);
    g.g_R = stage_out<R>(h, P, S_GR, g_R, B * m * r, &rc);
    g.g_Q = stage_out<R>(h, P, S_GQ, g_Q, B * r * r, &rc);
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    h->launches += 1;
    h->path = "stationary_warp";
    rc = stat::launch<R>(g, true, h->stream, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, g_c, g.g_c, B * m, &rc);
    copy_back<R>(h, P, g_T, g.g_T, B * m * m, &rc);
    copy_back<R>(h, P, g_R, g.g_R, B * m * r, &rc);
    copy_back<R>(h, P, g_Q, g.g_Q, B * r * r, &rc);
    return finish(h, P, rc);
}

static int validate_stationary(bkf_handle* h, const bkf_problem* P) {
    if (!h) return BKF_ERR_ARG;
    if (!P) return fail(h, BKF_ERR_ARG, "null problem");
    if (P->dtype != BKF_F64 && P->dtype != BKF_F32) return fail(h, BKF_ERR_ARG, "unknown dtype");
    if (P->mem != BKF_MEM_HOST && P->mem != BKF_MEM_DEVICE) return fail(h, BKF_ERR_ARG, "unknown memory space");
    if (P->B < 0 || P->m < 1 || P->r < 1) return fail(h, BKF_ERR_ARG, "non-positive dimension");
    if (P->tv_mask) return fail(h, BKF_ERR_UNSUPPORTED, "a time-varying model has no stationary initialisation");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    return BKF_OK;
}

extern "C" int bkf_stationary_init(bkf_handle* h, const bkf_problem* P, const void* c, const void* T, const void* R,
                                   const void* Q, void* a0, void* P0, int* status) {
    int rc = validate_stationary(h, P);
    if (rc != BKF_OK) return rc;
    if (!c || !T || !R || !Q) return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64 ? stationary_impl<double>(h, P, c, T, R, Q, a0, P0, status)
                               : stationary_impl<float>(h, P, c, T, R, Q, a0, P0, status);
}

extern "C" int bkf_stationary_init_grad(bkf_handle* h, const bkf_problem* P, const void* T, const void* R,
                                        const void* Q, const void* a0, const void* P0, const void* a0_bar,
                                        const void* P0_bar, void* g_c, void* g_T, void* g_R, void* g_Q) {
    int rc = validate_stationary(h, P);
    if (rc != BKF_OK) return rc;
    if (!T || !R || !Q || !a0 || !P0) return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64
               ? stationary_grad_impl<double>(h, P, T, R, Q, a0, P0, a0_bar, P0_bar, g_c, g_T, g_R, g_Q)
               : stationary_grad_impl<float>(h, P, T, R, Q, a0, P0, a0_bar, P0_bar, g_c, g_T, g_R, g_Q);
}

// ----------------------------------------------------------------------------------------------------
// multivariate-normal draws (bkf.h: bkf_mvn_sample)
template <typename R>
static int sample_impl(bkf_handle* h, const bkf_problem* P, const void* mu, const void* cov, const void* z, void* x,
                       int* status) {
    int rc = BKF_OK;
    const size_t B = P->B, n = P->n, m = P->m;
    const R* dmu = stage_in<R>(h, P, S_WFA, mu, B * n * m, &rc);
    const R* dcov = stage_in<R>(h, P, S_WFP, cov, B * n * m * m, &rc);
    const R* dz = stage_in<R>(h, P, S_Y, z, B * n * m, &rc);
    R* dx = stage_out<R>(h, P, S_SMA, x, B * n * m, &rc);
    int* dstatus = nullptr;
    if (status) {
        dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
        if (dstatus && rc == BKF_OK) cudaMemsetAsync(dstatus, 0, B * sizeof(int), h->stream);
    }
    if (rc != BKF_OK) return rc;
    if (B * n == 0) return BKF_OK;
    const char* err = "";
    h->launches += 1;
    h->path = "mvn_chol_subwarp";
    rc = smp::launch<R>(dmu, dcov, dz, smp::Rng{P->rng_seed, P->rng_offset}, dx, dstatus, (long long)(B * n), (int)n, (int)m, h->stream, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, x, dx, B * n * m, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_mvn_sample(bkf_handle* h, const bkf_problem* P, const void* mu, const void* cov, const void* z,
                              void* x, int* status) {
    if (!h) return BKF_ERR_ARG;
    if (!P) return fail(h, BKF_ERR_ARG, "null problem");
    if (P->dtype != BKF_F64 && P->dtype != BKF_F32) return fail(h, BKF_ERR_ARG, "unknown dtype");
    if (P->mem != BKF_MEM_HOST && P->mem != BKF_MEM_DEVICE) return fail(h, BKF_ERR_ARG, "unknown memory space");
    if (P->B < 0 || P->n < 0 || P->m < 1) return fail(h, BKF_ERR_ARG, "non-positive dimension");
    if (!mu || !cov || !x) return fail(h, BKF_ERR_ARG, "null pointer");  // (z == NULL: variates drawn in the kernel)
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    return P->dtype == BKF_F64 ? sample_impl<double>(h, P, mu, cov, z, x, status)
                               : sample_impl<float>(h, P, mu, cov, z, x, status);
}

template <typename R>
static int simulate_impl(bkf_handle* h, const bkf_problem* P, const void* const* prm, const void* z_x0, const void* z_y0,
                         const void* z_state, const void* z_obs, void* states, void* obs, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    fill_inputs<R>(h, P, g, nullptr, prm, &rc);
    const size_t B = P->B, n = P->n, m = P->m, p = P->p, r = P->r;
    const R* dzx = stage_in<R>(h, P, S_WFA, z_x0, B * m, &rc);
    const R* dzy = stage_in<R>(h, P, S_COT, z_y0, B * p, &rc);
    const R* dzs = stage_in<R>(h, P, S_WFP, z_state, B * n * r, &rc);
    const R* dzo = stage_in<R>(h, P, S_Y, z_obs, B * n * p, &rc);
    R* dxs = stage_out<R>(h, P, S_SMA, states, B * (n + 1) * m, &rc);
    R* dys = stage_out<R>(h, P, S_SMP, obs, B * (n + 1) * p, &rc);
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    h->launches += 1;
    h->path = "simulate_warp";
    rc = smp::launch_simulate<R>(g, dzx, dzy, dzs, dzo, smp::Rng{P->rng_seed, P->rng_offset}, dxs, dys, h->stream, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, states, dxs, B * (n + 1) * m, &rc);
    copy_back<R>(h, P, obs, dys, B * (n + 1) * p, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_simulate(bkf_handle* h, const bkf_problem* P, const void* a0, const void* P0, const void* c,
                            const void* d, const void* T, const void* Z, const void* R, const void* H, const void* Q,
                            const void* z_x0, const void* z_y0, const void* z_state, const void* z_obs, void* states,
                            void* obs, int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    const void* prm[BKF_NPARAM] = {a0, P0, c, d, T, Z, R, H, Q};
    const int nz = (z_x0 != nullptr) + (z_y0 != nullptr) + (z_state != nullptr) + (z_obs != nullptr);
    if (nz != 0 && nz != 4) return fail(h, BKF_ERR_ARG, "pass all four variate arrays, or none (in-kernel generator)");
    if (any_null(prm, BKF_NPARAM) || !states || !obs)
        return fail(h, BKF_ERR_ARG, "null pointer");
    return P->dtype == BKF_F64 ? simulate_impl<double>(h, P, prm, z_x0, z_y0, z_state, z_obs, states, obs, status)
                               : simulate_impl<float>(h, P, prm, z_x0, z_y0, z_state, z_obs, states, obs, status);
}

// ----------------------------------------------------------------------------------------------------
template <typename R>
static int smooth_impl(bkf_handle* h, const bkf_problem* P, const void* T, const void* Rm, const void* Q,
                       const void* fa, const void* fP, void* sa, void* sP, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    const void* prm[BKF_NPARAM] = {nullptr, nullptr, nullptr, nullptr, T, nullptr, Rm, nullptr, Q};
    bkf_problem P2 = *P;
    P2.p = P2.p > 0 ? P2.p : 1;
    fill_inputs<R>(h, &P2, g, nullptr, prm, &rc);
    const size_t B = P->B, n = P->n, m = P->m;
    g.sm_filt_a = stage_in<R>(h, P, S_WFA, fa, B * n * m, &rc);
    g.sm_filt_P = stage_in<R>(h, P, S_WFP, fP, B * n * m * m, &rc);
    g.sm_a = stage_out<R>(h, P, S_SMA, sa, B * n * m, &rc);
    g.sm_P = stage_out<R>(h, P, S_SMP, sP, B * n * m * m, &rc);
    int* dstatus = nullptr;
    if (status) {
        dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
        if (dstatus) cudaMemsetAsync(dstatus, 0, B * sizeof(int), h->stream);
    }
    g.status = dstatus;
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    rc = run_smoother(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, sa, g.sm_a, B * n * m, &rc);
    copy_back<R>(h, P, sP, g.sm_P, B * n * m * m, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_smooth(bkf_handle* h, const bkf_problem* P, const void* T, const void* R, const void* Q,
                          const void* fa, const void* fP, void* sa, void* sP, int* status) {
    int rc = validate(h, P, false);
    if (rc != BKF_OK) return rc;
    if (!T || !R || !Q || !fa || !fP) return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64 ? smooth_impl<double>(h, P, T, R, Q, fa, fP, sa, sP, status)
                               : smooth_impl<float>(h, P, T, R, Q, fa, fP, sa, sP, status);
}

// ----------------------------------------------------------------------------------------------------
template <typename R>
static int filter_smooth_impl(bkf_handle* h, const bkf_problem* P, const void* y, const void* const* prm, void* ll,
                              void* sa, void* sP, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    fill_inputs<R>(h, P, g, y, prm, &rc);
    const size_t B = P->B, n = P->n, m = P->m;
    g.ll = stage_out<R>(h, P, S_LL, ll, B * n, &rc);
    g.filt_a = (R*)ensure(h, S_WFA, B * n * m * sizeof(R), &rc);
    g.filt_P = (R*)ensure(h, S_WFP, B * n * m * m * sizeof(R), &rc);
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    g.sm_filt_a = g.filt_a;
    g.sm_filt_P = g.filt_P;
    g.sm_a = stage_out<R>(h, P, S_SMA, sa, B * n * m, &rc);
    g.sm_P = stage_out<R>(h, P, S_SMP, sP, B * n * m * m, &rc);
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    rc = run_forward(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    rc = run_smoother(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, ll, g.ll, B * n, &rc);
    copy_back<R>(h, P, sa, g.sm_a, B * n * m, &rc);
    copy_back<R>(h, P, sP, g.sm_P, B * n * m * m, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

// filter -> smoother -> one draw per (series, step), all on the device: only the draws (and the log-likelihood) go back
template <typename R>
static int filter_smooth_sample_impl(bkf_handle* h, const bkf_problem* P, const void* y, const void* const* prm,
                                     const void* z, void* ll, void* x, int* status) {
    int rc = BKF_OK;
    KArgs<R> g;
    fill_inputs<R>(h, P, g, y, prm, &rc);
    const size_t B = P->B, n = P->n, m = P->m;
    g.ll = stage_out<R>(h, P, S_LL, ll, B * n, &rc);
    g.filt_a = (R*)ensure(h, S_WFA, B * n * m * sizeof(R), &rc);
    g.filt_P = (R*)ensure(h, S_WFP, B * n * m * m * sizeof(R), &rc);
    int* dstatus = nullptr;
    if (status) dstatus = P->mem == BKF_MEM_DEVICE ? status : (int*)ensure(h, S_STATUS, B * sizeof(int), &rc);
    g.status = dstatus;
    g.sm_filt_a = g.filt_a;
    g.sm_filt_P = g.filt_P;
    g.sm_a = (R*)ensure(h, S_SMA, B * n * m * sizeof(R), &rc);
    g.sm_P = (R*)ensure(h, S_SMP, B * n * m * m * sizeof(R), &rc);
    const R* dz = stage_in<R>(h, P, S_COT, z, B * n * m, &rc);
    R* dx = stage_out<R>(h, P, S_GU, x, B * n * m, &rc);
    if (rc != BKF_OK) return rc;
    if (B == 0) return BKF_OK;
    const char* err = "";
    rc = run_forward(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    rc = run_smoother(h, g, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    h->launches += 1;
    rc = smp::launch<R>(g.sm_a, g.sm_P, dz, smp::Rng{P->rng_seed, P->rng_offset}, dx, dstatus, (long long)(B * n), (int)n, (int)m, h->stream, &err);
    if (rc != BKF_OK) return fail(h, rc, err);
    copy_back<R>(h, P, ll, g.ll, B * n, &rc);
    copy_back<R>(h, P, x, dx, B * n * m, &rc);
    if (status && P->mem == BKF_MEM_HOST && rc == BKF_OK) {
        cudaError_t e = cudaMemcpyAsync(status, dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H copy");
    }
    return finish(h, P, rc);
}

extern "C" int bkf_filter_smooth_sample(bkf_handle* h, const bkf_problem* P, const void* y, const void* a0,
                                        const void* P0, const void* c, const void* d, const void* T, const void* Z,
                                        const void* R, const void* H, const void* Q, const void* z, void* ll, void* x,
                                        int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    if (P->m > 32) return fail(h, BKF_ERR_UNSUPPORTED, "bkf_filter_smooth_sample: m > 32 is not supported");
    const void* prm[BKF_NPARAM] = {a0, P0, c, d, T, Z, R, H, Q};
    if (!y || any_null(prm, BKF_NPARAM) || !x) return fail(h, BKF_ERR_ARG, "null pointer");  // (z == NULL: in-kernel variates)
    return P->dtype == BKF_F64 ? filter_smooth_sample_impl<double>(h, P, y, prm, z, ll, x, status)
                               : filter_smooth_sample_impl<float>(h, P, y, prm, z, ll, x, status);
}

extern "C" int bkf_filter_smooth(bkf_handle* h, const bkf_problem* P, const void* y, const void* a0, const void* P0,
                                 const void* c, const void* d, const void* T, const void* Z, const void* R,
                                 const void* H, const void* Q, void* ll, void* sa, void* sP, int* status) {
    int rc = validate(h, P, true);
    if (rc != BKF_OK) return rc;
    const void* prm[BKF_NPARAM] = {a0, P0, c, d, T, Z, R, H, Q};
    if (!y || any_null(prm, BKF_NPARAM)) return fail(h, BKF_ERR_ARG, "null input pointer");
    return P->dtype == BKF_F64 ? filter_smooth_impl<double>(h, P, y, prm, ll, sa, sP, status)
                               : filter_smooth_impl<float>(h, P, y, prm, ll, sa, sP, status);
}

// ----------------------------------------------------------------------------------------------------
// bkf_comm_*: the final all-gather of [logp | gradients] through NCCL, loaded at run time (no link-time dependency)
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        static_assert(sizeof(ncclUniqueId) == BKF_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
        a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy && a.GetErrorString;
        return a;
    }();
    return api;
}
int nccl_fail(bkf_handle* h, ncclResult_t r) {
    if (h) h->err = std::string("NCCL: ") + nccl_api().GetErrorString(r);
    return BKF_ERR_CUDA;
}
}  // namespace

extern "C" int bkf_comm_unique_id(void* id) {
    if (!id) return BKF_ERR_ARG;
    NcclApi& a = nccl_api();
    if (!a.ok) return BKF_ERR_UNSUPPORTED;
    ncclUniqueId u;
    if (a.GetUniqueId(&u) != ncclSuccess) return BKF_ERR_CUDA;
    memcpy(id, &u, sizeof(u));
    return BKF_OK;
}

extern "C" int bkf_comm_init(bkf_handle* h, int world, int rank, const void* id) {
    if (!h || !id || world < 1 || rank < 0 || rank >= world) return BKF_ERR_ARG;
    NcclApi& a = nccl_api();
    if (!a.ok) { h->err = "libnccl.so.2 could not be loaded"; return BKF_ERR_UNSUPPORTED; }
    if (h->comm) bkf_comm_destroy(h);
    cudaSetDevice(h->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    const ncclResult_t r = a.CommInitRank(&c, world, u, rank);
    if (r != ncclSuccess) return nccl_fail(h, r);
    h->comm = c;
    h->comm_world = world;
    h->comm_rank = rank;
    return BKF_OK;
}

extern "C" int bkf_comm_allgather(bkf_handle* h, const void* send, void* recv, size_t count, int dtype) {
    if (!h || !send || !recv || (dtype != BKF_F64 && dtype != BKF_F32)) return BKF_ERR_ARG;
    if (!h->comm) { h->err = "bkf_comm_init has not been called on this handle"; return BKF_ERR_ARG; }
    cudaSetDevice(h->device);
    const ncclResult_t r = nccl_api().AllGather(send, recv, count, dtype == BKF_F64 ? ncclFloat64 : ncclFloat32,
                                               (ncclComm_t)h->comm, h->stream);
    return r == ncclSuccess ? BKF_OK : nccl_fail(h, r);
}

extern "C" int bkf_comm_destroy(bkf_handle* h) {
    if (!h) return BKF_OK;
    if (h->comm) {
        nccl_api().CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
        h->comm_world = h->comm_rank = 0;
    }
    return BKF_OK;
}
