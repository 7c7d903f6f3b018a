/*
 * svgd_b200.h — C ABI of the B200-native SVGD inner loop (libsvgd_b200.so).
 *
 * This is the drop-in boundary for ONE path of khaiyichin/SVGDCpp: everything reachable
 * from SVGD::Step() (include/SVGDCpp/SVGD.hpp:373-454 in the reference).  The reference has
 * no FFI of its own: its boundary is C++ virtual dispatch on shared_ptr<Kernel|Model|
 * Optimizer> plus the in-place shared Eigen::MatrixXd of particles (SVGD.hpp:151-162,176,393).
 * Each entry point below names the reference interface it replaces; the C++ facade in
 * include/SVGDCpp/ re-creates the reference's class API on top of these calls, and
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - All host matrices are the reference's layout: m x n column-major doubles, i.e.
 *     particle i occupies X[i*d .. i*d+d-1] ("particle-contiguous").
 *   - Every call returns SVGDB_OK (0) or a negative svgdb_status; nothing throws across the
 *     ABI.  svgdb_last_error() gives the message of the last failure on that context.
 *   - A context is driven by one host thread (like the non-copyable reference SVGD object,
 *     SVGD.hpp:256).  One context drives one GPU; multi-GPU = one context per GPU -- in one process
 *     each (torch.distributed.run style) or in one process with one host thread per context (what the
 *     facade's SVGDOptions::Devices does) -- joined by svgdb_comm_init (particle rows are sharded, see
 *     DESIGN.md "Multi-GPU").  Every entry point selects its context's device itself.
 *   - There is no CPU fallback: without a CUDA device svgdb_create fails with SVGDB_ERR_CUDA.
 */
#ifndef SVGD_B200_H
#define SVGD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svgdb_ctx svgdb_ctx;

typedef enum {
    SVGDB_OK = 0,
    SVGDB_ERR_INVALID = -1,   /* -> std::invalid_argument   (SVGD.hpp:223-236, Adam.hpp:45-48) */
    SVGDB_ERR_DIMENSION = -2, /* -> DimensionMismatchException (SVGD.hpp:170-173, MultivariateNormal.hpp:41-46) */
    SVGDB_ERR_UNSET = -3,     /* -> UnsetException            (Model.hpp:438-441, GaussianRBFKernel.hpp:53-56) */
    SVGDB_ERR_CUDA = -4,      /* -> std::runtime_error */
    SVGDB_ERR_NCCL = -5,      /* -> std::runtime_error */
    SVGDB_ERR_NOMEM = -6,     /* -> std::bad_alloc / runtime_error */
    SVGDB_ERR_NUMERIC = -7    /* singular covariance, non-finite bandwidth */
} svgdb_status;

/* Arithmetic of the pair interaction.  F64: IEEE double end to end (DMMA tensor cores),
 * oracle-grade.  TC32: tcgen05 kind::f16 tensor cores on split-fp16 operands (pair kernel; split-bf16
 * in the median's distance pass), fp32 accumulation in TMEM, fp32 ex2, kernel values rounded to
 * fp16, FP64 optimizer state; error bound stated in DESIGN.md "Precision modes". */
typedef enum { SVGDB_PRECISION_F64 = 0, SVGDB_PRECISION_TC32 = 1 } svgdb_precision;

/* Arithmetic variant of the tensor-core pair kernel under SVGDB_PRECISION_TC32 (no reference counterpart; ignored in F64 mode).
 * FAST: column particle and kernel values carry one fp16 term each (phi within 2e-4 of max|phi|; the error terms are zero-mean
 * and average over a row's neighbours; for d >= 48 and at least 16,384 particles the row particle's second fp16 term enters as an
 * e5m2 product and, from 32,768 particles, v = grad - 2a x~ carries one fp16 term as well: same bound).  PRECISE: both particles and the kernel values carry two fp16 terms (1.5x the MMAs; phi
 * within 1e-5 of max|phi|).  AUTO (default): FAST for one Gaussian target with at least 16,384 particles in d >= 8, PRECISE otherwise
 * (mixtures, gradient hooks, small particle sets). */
typedef enum { SVGDB_TC32_AUTO = 0, SVGDB_TC32_FAST = 1, SVGDB_TC32_PRECISE = 2 } svgdb_tc32_variant;

/* GaussianRBFKernel::ScaleMethod (Kernel/GaussianRBFKernel.hpp:25-30) plus a constant scale
 * (the reference's "TODO: constant scale", used by tests/test_svgd.cpp:97-106 via a bare Kernel). */
typedef enum { SVGDB_SCALE_MEDIAN = 0, SVGDB_SCALE_HESSIAN = 1, SVGDB_SCALE_FIXED = 2 } svgdb_scale_method;

/* Optimizer/{AdaGrad,Adam,RMSProp}.hpp */
typedef enum { SVGDB_OPT_ADAGRAD = 0, SVGDB_OPT_ADAM = 1, SVGDB_OPT_RMSPROP = 2 } svgdb_opt_kind;

/* Device-gradient hook: replaces Model::EvaluateLogModelGrad (Model/Model.hpp:335-338) for user
 * models.  Must enqueue, on `stream`, work that writes G[r*d + k] = d/dx_k log p(x_{row0+r}) for
 * r in [0, n_rows), reading X_dev (all N particles, particle-contiguous).  Device pointers. */
typedef int (*svgdb_grad_fn)(const double *X_dev, double *G_dev, int64_t n_total, int32_t d,
                             int64_t row0, int64_t n_rows, void *cuda_stream, void *user);

typedef struct {
    uint64_t iterations;          /* SVGD::Step calls executed */
    uint64_t kernel_launches;     /* CUDA kernels this context launched (all of them ours) */
    uint64_t median_passes;       /* full pairwise-distance passes spent on the median */
    uint64_t median_bracket_hits; /* iterations whose predicted bracket held (1 pass) */
    double last_scale;            /* a of the last iteration (A = a I, GaussianRBFKernel.hpp:187) */
    double ms_median, ms_grad, ms_phi, ms_comm; /* accumulated CUDA-event time per phase, if profiling */
    uint64_t phi_launches;        /* launches of the pair-interaction kernel inside ms_phi */
    double ms_phi_kernel;         /* ... of which the pair-interaction kernel alone (events around its launch) */
    double ms_grad_kernel;        /* grad log p alone (it runs on a side stream next to the median phase, so ms_grad only shows what sticks out) */
} svgdb_stats;

/* ---- lifecycle -------------------------------------------------------------------------- */

/* Replaces the SVGD constructor's allocation of X/G/K/grad-K scratch (SVGD.hpp:176-181): the ctx
 * owns all device memory; K and grad K are never materialised.  n_total particles, dimension d. */
int svgdb_create(svgdb_ctx **out, int device, int64_t n_total, int32_t d, int precision_mode);
void svgdb_destroy(svgdb_ctx *ctx);
const char *svgdb_last_error(const svgdb_ctx *ctx);
const char *svgdb_version(void);

/* Number of visible CUDA devices (what the facade's `Parallel` flag spreads a run over). */
int svgdb_device_count(int *count);

/* Use the caller's CUDA stream (a cudaStream_t) for all work; NULL restores the ctx's own. */
int svgdb_set_stream(svgdb_ctx *ctx, void *cuda_stream);

/* Row sharding over `world` GPUs (no reference counterpart: the reference's only parallelism is
 * `#pragma omp parallel for` over particles, SVGD.hpp:418).  nccl_unique_id: the 128-byte
 * ncclUniqueId obtained by rank 0 from svgdb_nccl_unique_id and distributed by the caller. */
int svgdb_nccl_unique_id(void *out_id, size_t bytes);
int svgdb_comm_init(svgdb_ctx *ctx, int world, int rank, const void *nccl_unique_id, size_t bytes);

/* ---- particles (the shared coordinate matrix, SVGD.hpp:176,393) ----------------------------
 * The set_* calls return after the host buffer has been read (pinned or pageable): the caller may
 * reuse or free it at once. */
int svgdb_set_particles(svgdb_ctx *ctx, const double *X_dxN);
int svgdb_get_particles(svgdb_ctx *ctx, double *X_dxN);
/* Row-sharded I/O for multi-rank runs (after svgdb_comm_init): this rank owns particles [row0, row0 + n_rows).
 * set_particles_rows uploads only those (n_rows x d doubles, particle-contiguous) and all-gathers the rest over NVLink;
 * get_particles_rows downloads only those.  With one rank they equal svgdb_set_particles / svgdb_get_particles. */
int svgdb_local_rows(svgdb_ctx *ctx, int64_t *row0, int64_t *n_rows);
int svgdb_set_particles_rows(svgdb_ctx *ctx, const double *rows_local);
int svgdb_get_particles_rows(svgdb_ctx *ctx, double *rows_local);

/* ---- model: replaces MultivariateNormal (Model/MultivariateNormal.hpp:39-64) and the
 * `mvn1 + mvn2 (+ ...)` sum built with Model::operator+ (Model/Model.hpp:55-92): p = sum_c
 * exp(-1/2 (x-mu_c)^T Sigma_c^-1 (x-mu_c)), unweighted and unnormalised.  means: d x C,
 * covs: d x d x C (column-major blocks, symmetric). */
int svgdb_set_model_mvn(svgdb_ctx *ctx, const double *mean_d, const double *cov_dxd);
int svgdb_set_model_mvn_sum(svgdb_ctx *ctx, int32_t n_components, const double *means_dxC,
                            const double *covs_dxdxC);
int svgdb_set_model_device_hook(svgdb_ctx *ctx, svgdb_grad_fn fn, void *user);

/* ---- kernel: GaussianRBFKernel(x0, ScaleMethod, model) (Kernel/GaussianRBFKernel.hpp:47-88);
 * the scale is recomputed from the current X every Step (:141-156). */
int svgdb_set_kernel_rbf(svgdb_ctx *ctx, int scale_method, double fixed_a);

/* Selects the arithmetic variant of the tensor-core pair kernel (svgdb_tc32_variant above). */
int svgdb_set_tc32_variant(svgdb_ctx *ctx, int variant);

/* ---- optimizer: Adam(dim,n,lr,b1,b2,eps) / AdaGrad(dim,n,lr,eps) / RMSProp(dim,n,lr,b,eps)
 * (Optimizer/Adam.hpp:33-49, AdaGrad.hpp:31-37, RMSProp.hpp:33-46); beta1 is RMSProp's decay. */
int svgdb_set_optimizer(svgdb_ctx *ctx, int kind, double lr, double beta1, double beta2, double eps);

/* ---- bounds (SVGD.hpp:184-216, 396-399); NULL, NULL disables the clamp.  n_bound = 1
 * (replicated to every coordinate) or d. */
int svgdb_set_bounds(svgdb_ctx *ctx, const double *lb, const double *ub, int32_t n_bound);

/* ---- SVGD::Initialize (SVGD.hpp:268-296): zero optimizer state, counter = 0. */
int svgdb_initialize(svgdb_ctx *ctx);

/* ---- SVGD::Step x iters == SVGD::Run (SVGD.hpp:338-400).  Asynchronous: X stays on the device. */
int svgdb_step(svgdb_ctx *ctx, int64_t iters);
/* The same on host memory, the way the reference's Step works on its coordinate matrix (SVGD.hpp:373-400): uploads this
 * rank's rows (svgdb_set_particles_rows), runs `iters` >= 1 steps and returns the updated rows in rows_out
 * (svgdb_get_particles_rows), which may be the same buffer.  Synchronous.  On the tensor-core path the last iteration's
 * pair kernel runs in four row chunks and each chunk's rows start their device-to-host copy as soon as they are updated, so
 * most of the PCIe transfer overlaps the remaining pair interactions; pinned host memory (svgdb_host_alloc) is needed for
 * that overlap, pageable memory works but serialises. */
int svgdb_step_host(svgdb_ctx *ctx, const double *rows_in, double *rows_out, int64_t iters);

/* SVGD::ComputePhi (SVGD.hpp:407-454) on the current X, no update: writes phi (d x N) and the
 * kernel scale used.  Either output may be NULL.  Synchronous. */
int svgdb_compute_phi(svgdb_ctx *ctx, double *phi_dxN, double *scale_out);

/* GaussianRBFKernel::ComputeScale (Kernel/GaussianRBFKernel.hpp:164-212) on the current X. */
int svgdb_compute_scale(svgdb_ctx *ctx, double *scale_out);
/* The kernel's inverse scale matrix of the last Step / compute call, d x d (GaussianRBFKernel::GetParameters()[0]):
 * a * I for the median and fixed methods; for SVGDB_SCALE_HESSIAN (MultivariateNormal models and sums of them)
 * A = 1/(2 d n) sum_i -Hessian(log p)(x_i) (GaussianRBFKernel.hpp:189-210), which must be positive definite. */
int svgdb_get_scale_matrix(svgdb_ctx *ctx, double *A_dxd);

/* Model::EvaluateLogModel for every particle (Model/Model.hpp:305-308): logp has N entries.  Built-in Gaussian models only (a model
 * behind the gradient hook has no value function here); evaluated through log-sum-exp, so it stays finite where the reference's
 * literal log(sum exp) underflows. */
int svgdb_compute_log_model(svgdb_ctx *ctx, double *logp_N);

/* The matrices SVGDOptions::LogIntermediateMatrices prints (SVGD.hpp:346-365; filled in ComputePhi, :407-454), for the current X and
 * the scale Kernel::Step would compute from it: K is n x n column-major with K(j, i) = k(x_j, x_i); gradK is (n d) x n column-major
 * with the block (j d .. j d + d, i) = grad_x k(x, x_i) at x = x_j.  An inspection path: one rank, n^2 (d + 1) doubles <= 2 GiB; the
 * step itself never forms these matrices.  scale_out (may be NULL) receives the scale (A(0,0) for the Hessian method). */
int svgdb_compute_kernel_matrices(svgdb_ctx *ctx, double *K_nxn, double *gradK_ndxn, double *scale_out);

/* Model::EvaluateLogModelGrad for every particle (Model/Model.hpp:335-338): G is d x N. */
int svgdb_compute_log_model_grad(svgdb_ctx *ctx, double *G_dxN);

/* Optimizer state for checkpoint / mid-trajectory parity: state1 = sum of squares or 2nd moment,
 * state2 = 1st moment (Adam), both d x N; counter = Adam's step count (Adam.hpp:98-110). */
int svgdb_get_opt_state(svgdb_ctx *ctx, double *state1_dxN, double *state2_dxN, uint64_t *counter);
int svgdb_set_opt_state(svgdb_ctx *ctx, const double *state1_dxN, const double *state2_dxN, uint64_t counter);

int svgdb_sync(svgdb_ctx *ctx);

/* ---- measurement ------------------------------------------------------------------------ */
int svgdb_set_profiling(svgdb_ctx *ctx, int enabled); /* CUDA events around each phase */
int svgdb_get_stats(svgdb_ctx *ctx, svgdb_stats *out);
int svgdb_reset_stats(svgdb_ctx *ctx);
/* Time `iters` Steps with CUDA events on the ctx stream (device time, ms). */
int svgdb_time_steps(svgdb_ctx *ctx, int64_t iters, float *ms_out);

/* Pinned host staging buffers for the particle matrix (plain cudaMallocHost / cudaFreeHost). */
int svgdb_host_alloc(void **out, size_t bytes);
int svgdb_host_free(void *ptr);

/* Measurement aid (SVGDB_PRECISION_TC32): average time in ms of `reps` back-to-back launches of one hot kernel on the
 * context's current particles, CUDA events on the context's stream.  which = 0: the tensor-core distance pass of the
 * median bandwidth (GaussianRBFKernel.hpp:168-188), collecting the bracket median * (1 +- rel_halfwidth) of the last
 * step (variant 0 = product path, 1 = no counting, 2 = counting without collecting); which = 1: the pair-interaction
 * pass (SVGD.hpp:407-454) with its operand preparation.  Particles and optimizer state are not modified. */
int svgdb_time_kernel(svgdb_ctx *ctx, int which, int reps, int variant, double rel_halfwidth, float *ms_out);

/* Micro-probes for roofline denominators the driver does not measure.  what = 0: FP64 DMMA
 * (mma.sync m8n8k4) TFLOP/s from a register-resident issue loop on every SM. */
int svgdb_probe_peak(int device, int what, double *out);

#ifdef __cplusplus
}
#endif
#endif /* SVGD_B200_H */
