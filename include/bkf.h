/*
 * bkf.h -- C ABI of libbkf.so, the B200 (sm_100a) batched Kalman engine.
 *
 * This is the drop-in boundary for the hot path of pymc_extras.statespace (paths relative to the reference
 * repository pymc-devs/pymc-experimental):
 *
 *   bkf_filter_forward    replaces  BaseFilter.build_graph + kalman_step + _postprocess_scan_results
 *                                   pymc_extras/statespace/filters/kalman_filter.py:144-308 (Standard :549-617,
 *                                   SquareRoot :629-750, Univariate :769-820)
 *   bkf_filter_logp_grad  replaces  sum(loglike_obs) and pytensor.grad(...) through Scan.L_op of the same graph
 *                                   (the logp + dlogp function NUTS evaluates)
 *   bkf_smooth            replaces  KalmanSmoother.build_graph, filters/kalman_smoother.py:66-126
 *   bkf_filter_smooth     fuses     filter forward + smoother (PyMCStateSpace._build_smoother_graph,
 *                                   core/statespace.py:903-951) without materialising the filter outputs
 *
 * The reference has no FFI of its own (it is pure Python over PyTensor); the reference-side binding is a
 * PyTensor Op whose perform() calls these entry points through ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - All arrays are C-contiguous, row-major, dtype = problem.dtype (float64 or float32), caller-owned.
 *     The library never frees or retains caller pointers past the call.
 *   - problem.mem says whether the pointers are host (NumPy) or device (CUDA, same device as the handle).
 *   - Batch axis first: a0 [B,m], P0 [B,m,m], c [B,m], d [B,p], T [B,m,m], Z [B,p,m], R [B,m,r], H [B,p,p],
 *     Q [B,r,r].  Bit i of problem.shared_mask (BKF_PARAM_* order) set means "parameter i has no batch axis
 *     and is shared by every series".  y is [B,n,p], or [n,p] when problem.y_shared != 0.
 *   - Outputs: filtered_states [B,n,m], predicted_states [B,n,m], observed_states [B,n,po],
 *     filtered_covariances [B,n,m,m], predicted_covariances [B,n,m,m], observed_covariances [B,n,po,po],
 *     loglike_obs [B,n]; po = p, except po = 1 for the univariate filter (kalman_filter.py:808-813).
 *     Any output pointer may be NULL (not computed / not copied back).
 *   - Gradients have the shape of the corresponding input (always with the batch axis, also for shared
 *     parameters: the caller reduces over B if it shares them).  Every matrix entry is an independent
 *     variable (no symmetrisation of dP0, dH, dQ), as pytensor.grad of the reference formulas yields.
 *   - Return value: 0 on success; < 0 is a call failure (bad argument, unsupported configuration, CUDA
 *     error) described by bkf_last_error().  Numerical trouble of a single series (non-PD F, non-finite
 *     logp) never fails the call: it is reported in status[b] (BKF_STATUS_*).
 *   - A handle owns one device, one workspace and uses one stream; it is not thread-safe.  Create one handle
 *     per host thread / process (PyMC forks one process per chain: create lazily after the fork).
 */
#ifndef BKF_H
#define BKF_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BKF_VERSION 1

/* filter kinds == FILTER_FACTORY keys of pymc_extras/statespace/core/statespace.py:51-55 */
enum { BKF_STANDARD = 0, BKF_UNIVARIATE = 1, BKF_CHOLESKY = 2 };
enum { BKF_F64 = 0, BKF_F32 = 1 };
enum { BKF_MEM_HOST = 0, BKF_MEM_DEVICE = 1 };
/* parameter order == MATRIX_NAMES of pymc_extras/statespace/utils/constants.py:24 */
enum { BKF_PARAM_A0 = 0, BKF_PARAM_P0, BKF_PARAM_C, BKF_PARAM_D, BKF_PARAM_T, BKF_PARAM_Z, BKF_PARAM_R,
       BKF_PARAM_H, BKF_PARAM_Q, BKF_NPARAM };

/* call-level error codes */
enum { BKF_OK = 0, BKF_ERR_ARG = -1, BKF_ERR_UNSUPPORTED = -2, BKF_ERR_CUDA = -3, BKF_ERR_NOMEM = -4,
       BKF_ERR_STALE = -5 /* bkf_filter_backward_saved: the saved forward sweep is gone, call bkf_filter_logp_grad */ };
/* per-series status bits */
enum { BKF_STATUS_OK = 0, BKF_STATUS_NOT_PD = 1, BKF_STATUS_NONFINITE = 2,
       BKF_STATUS_SMOOTHER_ILL = 4 /* P_hat too ill-conditioned for the Cholesky smoother: series was re-run with the
                                      eigen-decomposition pinv (informational) */,
       BKF_STATUS_NOT_STATIONARY = 8 /* bkf_stationary_init: the powers of T do not decay (spectral radius >= 1) */ };

typedef struct bkf_handle bkf_handle;

typedef struct bkf_problem {
    int kind;              /* BKF_STANDARD / BKF_UNIVARIATE / BKF_CHOLESKY */
    int dtype;             /* BKF_F64 / BKF_F32 */
    int mem;               /* BKF_MEM_HOST / BKF_MEM_DEVICE: where every pointer of the call lives */
    int B;                 /* number of independent series / parameter draws */
    int n;                 /* time steps */
    int m, p, r;           /* k_states, k_endog, k_posdef */
    int y_shared;          /* 1: y is [n,p] and shared by all series */
    unsigned shared_mask;  /* bit BKF_PARAM_x: parameter x has no batch axis */
    unsigned tv_mask;      /* bit BKF_PARAM_x: parameter x carries a time axis of length n right after the (optional)
                              batch axis: c [B,n,m], d [B,n,p], T [B,n,m,m], Z [B,n,p,m], R [B,n,m,r], H [B,n,p,p],
                              Q [B,n,r,r] (filters/utilities.py:8-47; never a0, P0).  Step t uses x[t]
                              (kalman_filter.py:110-142); the smoother pairs filtered[t] with x[t+1]
                              (kalman_smoother.py:84-92).  Gradients of time-varying
                              parameters have the same time axis.  All three filters and the smoother. */
    double cov_jitter;     /* JITTER_DEFAULT: 1e-8 (f64) / 1e-6 (f32), utils/constants.py:20 */
    double missing_fill;   /* MISSING_FILL = -9999.0, utils/constants.py:19 */
    unsigned long long rng_seed;   /* sampling calls with a NULL variate pointer draw their standard normals in the kernel: */
    unsigned long long rng_offset; /* Philox4x32-10 keyed by rng_seed, Box-Muller; variate i of variate array s of the call is
                                      a pure function of (rng_seed, rng_offset, s, i) -- independent of the launch geometry
                                      and of B (see bkf_mvn_sample).  Advance rng_offset between calls (any stride >= the
                                      largest variate array of a call), as a counter-based generator is used. */
} bkf_problem;

int bkf_version(void);
int bkf_create(int device, bkf_handle** out);
int bkf_destroy(bkf_handle* h);
const char* bkf_last_error(const bkf_handle* h);
/* Launch on this CUDA stream (a cudaStream_t passed as void*); NULL = the legacy default stream. */
int bkf_set_stream(bkf_handle* h, void* cuda_stream);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
long long bkf_launch_count(const bkf_handle* h);
/* Name of the kernel path picked by the last call ("generic_warp", "subwarp_m15", ...). */
const char* bkf_last_path(const bkf_handle* h);

/* Kernel timing for bench.py's roofline: when enabled, every kernel launch is bracketed by CUDA events on the
 * handle's stream.  bkf_kernel_time_ms synchronises those events and returns the accumulated device time and
 * launch count of one kernel class since the last bkf_reset_timing. */
enum { BKF_KERNEL_FORWARD = 0, BKF_KERNEL_BACKWARD = 1, BKF_KERNEL_SMOOTHER = 2, BKF_NKERNEL };
int bkf_enable_timing(bkf_handle* h, int on);
int bkf_reset_timing(bkf_handle* h);
int bkf_kernel_time_ms(bkf_handle* h, int kernel_class, double* total_ms, long long* launches);
/* DFMA / FFMA issue-rate microbenchmark (register-resident FMA chains on every SM): the measured FP64 / FP32
 * pipe peak in TFLOP/s that roofline fractions are quoted against (MEASURED_PEAKS.json has no such entry). */
int bkf_measure_fma_peak(bkf_handle* h, int dtype, double* tflops);

/* Page-locked host memory for call outputs / inputs (cudaHostAlloc / cudaFreeHost).  Host-pointer calls work with
 * any memory; with pinned buffers the H2D / D2H copies inside a call run at full PCIe / C2C rate instead of being
 * staged through the driver's bounce buffer.  The host side of the PyTensor Op keeps a pool of these. */
int bkf_host_alloc(size_t bytes, void** out);
int bkf_host_free(void* p);

/* Device workspace (bytes) a bkf_filter_logp_grad call on `prob` needs beyond the caller's own buffers: the forward
 * sweep's trajectory record (layout private to the kernel family), plus device staging of every input and output
 * when prob->mem is BKF_MEM_HOST.  With bkf_device_memory it lets the host side split a batch that would not fit
 * (KalmanEngine.logp_grad does: series are independent, so a batch can be processed in slices). */
int bkf_logp_grad_workspace_bytes(const bkf_problem* prob, size_t* bytes);
int bkf_device_memory(bkf_handle* h, size_t* free_bytes, size_t* total_bytes);

int bkf_filter_forward(bkf_handle* h, const bkf_problem* prob, const void* y, const void* a0, const void* P0,
                       const void* c, const void* d, const void* T, const void* Z, const void* R, const void* H,
                       const void* Q, void* filtered_states, void* predicted_states, void* observed_states,
                       void* filtered_covariances, void* predicted_covariances, void* observed_covariances,
                       void* loglike_obs, int* status);

/* The two halves of bkf_filter_logp_grad as separate calls, for a host framework whose forward and gradient are
 * separate graph nodes (PyTensor: Op.perform of the log-likelihood node, then Op.perform of its L_op node; reference:
 * pytensor.grad through Scan.L_op of kalman_filter.py:144-308).  bkf_filter_forward_saved runs the forward sweep ONCE,
 * returns loglike_obs [B, n] and / or logp [B] (either may be NULL) and keeps the sweep's record in the handle;
 * *ticket names it.  bkf_filter_backward_saved runs only the adjoint sweep on that record (same semantics of
 * ll_cotangent and of the g_* pointers as bkf_filter_logp_grad).  The record lives until the next call on the handle
 * that uses its workspace: after that, or with another ticket, bkf_filter_backward_saved returns BKF_ERR_STALE and has
 * done nothing (the caller falls back to bkf_filter_logp_grad).  With BKF_MEM_DEVICE the caller's input buffers must
 * stay unchanged between the two calls (with BKF_MEM_HOST the staged copies are used).  One ticket, one adjoint. */
int bkf_filter_forward_saved(bkf_handle* h, const bkf_problem* prob, const void* y, const void* a0, const void* P0,
                             const void* c, const void* d, const void* T, const void* Z, const void* R, const void* H,
                             const void* Q, void* loglike_obs, void* logp, int* status, unsigned long long* ticket);
int bkf_filter_backward_saved(bkf_handle* h, unsigned long long ticket, const void* ll_cotangent, void* g_a0,
                              void* g_P0, void* g_c, void* g_d, void* g_T, void* g_Z, void* g_R, void* g_H, void* g_Q);

/* logp[b] = sum_t loglike_obs[b,t];  g_x = d(sum_t ll_cotangent[b,t] * loglike_obs[b,t]) / dx
 * (ll_cotangent == NULL means all ones, the NUTS case).  Any g_* pointer may be NULL. */
int bkf_filter_logp_grad(bkf_handle* h, const bkf_problem* prob, const void* y, const void* a0, const void* P0,
                         const void* c, const void* d, const void* T, const void* Z, const void* R,
                         const void* H, const void* Q, const void* ll_cotangent, void* logp, void* g_a0,
                         void* g_P0, void* g_c, void* g_d, void* g_T, void* g_Z, void* g_R, void* g_H,
                         void* g_Q, int* status);

/* Compact parameterisation: every state-space model of the reference is a constant
 * pattern with a few parametrised entries (models/SARIMAX.py:383-536, ETS.py:468-670, structural.py components,
 * VARMAX.py:297-393), so a caller that evaluates many draws ships, per draw, only those entries:
 *   base[9]     the nine matrices WITHOUT batch axis (shared constants; prob->shared_mask is ignored)
 *   nnz, which[nnz], index[nnz]   entry k of the compact vector lands in matrix which[k] (BKF_PARAM_x) at flat
 *               row-major index index[k]; (which, index) pairs must be unique; index arrays are HOST memory
 *   u   [B, nnz]  the per-draw values;  g_u [B, nnz]  d(sum_t cot * ll)/du  (the matrix gradients gathered at the
 *               same positions: the caller chains them through its own theta -> u map)
 * The per-draw matrices are expanded on the device (base broadcast + scatter), the ordinary kernels run, and only
 * logp, g_u and status travel back: for C2 that is 7 instead of 513 doubles per draw in each direction.  y, u, ll_cot,
 * logp, g_u, status live where prob->mem says; time-varying inputs are not supported here. */
int bkf_filter_logp_grad_compact(bkf_handle* h, const bkf_problem* prob, const void* y, const void* const base[9],
                                 int nnz, const int* which, const int* index, const void* u, const void* ll_cotangent,
                                 void* logp, void* g_u, int* status);

/* theta -> matrices on the device (the rest of row f-1): the entries of the compact vector are themselves simple
 * functions of a model's free parameters -- SARIMA: phi_i, Phi_j, -phi_i Phi_j, theta, Theta, theta Theta, sigma^2
 * (models/SARIMAX.py:404-458); ETS: alpha, alpha beta, 1 - alpha, phi ... (models/ETS.py:499-555); VARMAX: the AR / MA
 * coefficients themselves and Q = L L^T (models/VARMAX.py:297-393); structural components: rho cos(lambda),
 * rho sin(lambda), sigma^2 (models/structural.py:858-875, 1220-1237, 1489-1514) -- all of the form
 *     u_k(theta) = sum over the terms of entry k of   coef * prod_{f < nfac} fn_f(theta[idx_f]),   nfac <= 3,
 * with fn one of BKF_FN_*.  The caller ships theta [B, ntheta] only; the library evaluates u on the device, expands the
 * matrices, runs the filter and its adjoint, and applies the chain rule: g_theta [B, ntheta] = d(sum_t cot ll)/dtheta.
 * Term q: entry term_entry[q] in [0, nnz), coefficient term_coef[q], term_nfac[q] factors with parameter indices
 * term_theta[3q + f] and function codes term_fn[3q + f] (HOST arrays, like which / index; any order).  An entry without
 * terms is 0; a constant is a term with nfac = 0.  base, which, index as in bkf_filter_logp_grad_compact; y, theta, ll_cot,
 * logp, g_theta (nullable), status live where prob->mem says. */
enum { BKF_FN_ID = 0, BKF_FN_COS = 1, BKF_FN_SIN = 2, BKF_FN_EXP = 3, BKF_FN_ONE_MINUS = 4, BKF_FN_SQUARE = 5 };
int bkf_filter_logp_grad_theta(bkf_handle* h, const bkf_problem* prob, const void* y, const void* const base[9], int nnz,
                               const int* which, const int* index, int nterms, const int* term_entry,
                               const double* term_coef, const int* term_nfac, const int* term_theta, const int* term_fn,
                               int ntheta, const void* theta, const void* ll_cotangent, void* logp, void* g_theta,
                               int* status);

/* Stationary initialisation, the step right before the filter in the reference's SARIMA /
 * VARMA / ETS models (models/SARIMAX.py:369-381, VARMAX.py:379-393, ETS.py:453-466):
 *     a0 = (I - T)^{-1} c          P0 = solve_discrete_lyapunov(T, R Q R^T)      (P0 = T P0 T^T + R Q R^T)
 * Uses prob->{dtype,mem,B,m,r,shared_mask(c,T,R,Q)}.  a0 [B,m], P0 [B,m,m].  status[b] gets
 * BKF_STATUS_NOT_STATIONARY when the series in T does not converge. */
int bkf_stationary_init(bkf_handle* h, const bkf_problem* prob, const void* c, const void* T, const void* R,
                        const void* Q, void* a0, void* P0, int* status);
/* Pull-back of bkf_stationary_init: given its results a0, P0 and their cotangents a0_bar [B,m] (may be NULL),
 * P0_bar [B,m,m] (may be NULL), writes g_c [B,m], g_T [B,m,m], g_R [B,m,r], g_Q [B,r,r] (any may be NULL) -- what
 * PyTensor's SolveDiscreteLyapunov / Solve L_ops yield. */
int bkf_stationary_init_grad(bkf_handle* h, const bkf_problem* prob, const void* T, const void* R, const void* Q,
                             const void* a0, const void* P0, const void* a0_bar, const void* P0_bar, void* g_c,
                             void* g_T, void* g_R, void* g_Q);

/* Uses prob->{dtype,mem,B,n,m,r,shared_mask(T,R,Q),cov_jitter}.  filtered_* are [B,n,m] / [B,n,m,m]. */
int bkf_smooth(bkf_handle* h, const bkf_problem* prob, const void* T, const void* R, const void* Q,
               const void* filtered_states, const void* filtered_covariances, void* smoothed_states,
               void* smoothed_covariances, int* status);

/* Filter forward followed by the smoother; loglike_obs / smoothed_* may be NULL. */
int bkf_filter_smooth(bkf_handle* h, const bkf_problem* prob, const void* y, const void* a0, const void* P0,
                      const void* c, const void* d, const void* T, const void* Z, const void* R, const void* H,
                      const void* Q, void* loglike_obs, void* smoothed_states, void* smoothed_covariances,
                      int* status);

/* The whole of sample_conditional_posterior (core/statespace.py:1061-1171) for a batch of posterior draws in one call:
 * filter forward, smoother, then one draw x[b,t] ~ N(smoothed_state[b,t], smoothed_cov[b,t]) (bkf_mvn_sample below) --
 * the smoothed moments never leave the device; only x [B,n,m] (and loglike_obs [B,n], nullable) come back.
 * z: standard normal variates [B,n,m] supplied by the caller, or NULL: drawn in the kernel (prob->rng_seed / rng_offset,
 * variate array 0) -- then nothing but y and the parameters goes in and only x comes out.  m <= 32. */
int bkf_filter_smooth_sample(bkf_handle* h, const bkf_problem* prob, const void* y, const void* a0, const void* P0,
                             const void* c, const void* d, const void* T, const void* Z, const void* R, const void* H,
                             const void* Q, const void* z, void* loglike_obs, void* x, int* status);

/* Row f-3 of SURVEY.md section 8: one multivariate-normal draw per (series, step),
 *     x[b,t] = mu[b,t] + L[b,t] z[b,t],   L L^T = cov[b,t],
 * the batched form of SequenceMvNormal.rv_op (pymc_extras/statespace/filters/distributions.py:411-443), which scans
 * MvNormalSVD (:52-87) over time to sample the conditional posterior from the smoother's output
 * (core/statespace.py:1119-1150).  L is the lower Cholesky factor with semidefinite pivots dropped (a pivot below
 * 64 m eps max(diag) zeroes its column): like the reference's SVD square root it accepts rank-deficient covariances, and
 * the distribution of x is the same for every square root.  z: standard normal variates drawn by the caller, or NULL:
 * generated in the kernel by the counter-based generator of prob->rng_seed / rng_offset (variate array 0, variate
 * i = flat index into [B, n, m]: counter rng_offset + i / 2 -> 128 Philox bits -> two uniforms in (0, 1) -> one
 * Box-Muller pair, variate i takes element i & 1) -- either way the call is deterministic.
 * Uses prob->B, n, m, dtype, mem; mu, z, x: [B, n, m]; cov: [B, n, m, m] (lower triangle
 * read); m <= 32.  status[B] (nullable): BKF_STATUS_NOT_PD if some cov[b,t] has a pivot below -1e-8 max(diag)
 * (-1e-4 in float32), numpy's check_valid tolerance. */
int bkf_mvn_sample(bkf_handle* h, const bkf_problem* prob, const void* mu, const void* cov, const void* z, void* x,
                   int* status);

/* Row f-3, second part: unconditional simulation of B state-space models for prob->n steps -- the batched form of
 * LinearGaussianStateSpace.rv_op (filters/distributions.py:181-293; prior / forecast / IRF draws):
 *     x_0 ~ N(a0, P0),  y_0 ~ N(Z x_0, H)                      (:262-263: no d in the first observation mean)
 *     x_t = c + T x_{t-1} + R eps_t,  eps_t ~ N(0, Q) ;  y_t = d + Z x_t + eta_t,  eta_t ~ N(0, H)       (:247-254)
 * Every normal draw is (Cholesky factor with dropped semidefinite pivots, as bkf_mvn_sample) x (standard normal variates
 * supplied by the caller): z_x0 [B,m], z_y0 [B,p], z_state [B,n,r], z_obs [B,n,p] -- or all four NULL: generated in the
 * kernel (prob->rng_seed / rng_offset; variate arrays 1, 2, 3, 4 in that order).  Outputs: states [B,n+1,m],
 * obs [B,n+1,p] (row 0 = the initial draw; the reference concatenates the two along the last axis, append_x0=True).
 * Parameters follow prob->shared_mask and prob->tv_mask (sequences of length n: step t uses x[t], the initial
 * observation Z[0] and H[0], :259-260).  status[B] (nullable):
 * BKF_STATUS_NOT_PD if P0, H or Q is not positive semidefinite. */
int bkf_simulate(bkf_handle* h, const bkf_problem* prob, const void* a0, const void* P0, const void* c, const void* d,
                 const void* T, const void* Z, const void* R, const void* H, const void* Q, const void* z_x0,
                 const void* z_y0, const void* z_state, const void* z_obs, void* states, void* obs, int* status);

/* ---- The one collective of the path (SURVEY.md section 8 rows b / e): draws and series are sharded over the GPUs of a box
 * with no traffic during filtering; at the end every rank contributes its packed [logp | gradients] buffer to ONE NCCL
 * all-gather over NVLink.  These entry points let a C caller do that without PyTorch (the Python host side issues the same
 * collective through torch.distributed on the buffer libbkf wrote in place).  NCCL (libnccl.so.2) is loaded on first use;
 * without it the calls return BKF_ERR_UNSUPPORTED and the rest of the library is unaffected.
 *   rank 0:      bkf_comm_unique_id(id)           -> ship the 128 bytes to the other ranks by any means
 *   every rank:  bkf_comm_init(h, world, rank, id) -> a communicator on h's device (one process per GPU)
 *                bkf_comm_allgather(h, send, recv, count, dtype): recv[world * count] = concatenation over ranks of
 *                send[count]; DEVICE pointers, enqueued on the handle's stream behind the kernels that produced send
 *                bkf_comm_destroy(h)                (bkf_destroy does it too) */
#define BKF_COMM_ID_BYTES 128
int bkf_comm_unique_id(void* id);
int bkf_comm_init(bkf_handle* h, int world, int rank, const void* id);
int bkf_comm_allgather(bkf_handle* h, const void* send, void* recv, size_t count, int dtype);
int bkf_comm_destroy(bkf_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* BKF_H */
