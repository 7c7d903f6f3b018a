/*
 * svgd_oracle.h — CPU restatement of the SVGDCpp hot path (TEST INFRASTRUCTURE).
 *
 * This is the parity ORACLE, not product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` leg may build, load or call it.
 * The product path (svgdcpp_b200/csrc, include/) never links or imports it.
 *
 * Parity status: PINNED.  The restatement reproduces every printed digit of the
 * reference's own published example outputs (examples/README.md:7-12 and cell 4 of
 * both example notebooks); see tests/test_oracle_golden.py and tests/golden/.
 * The reference itself cannot be compiled here: it needs Eigen and CppAD, neither is
 * installed and there is no network (DESIGN.md "Oracle").
 *
 * Layout convention everywhere: the particle matrix is the reference's m x n
 * column-major Eigen::MatrixXd (SVGD.hpp:176), i.e. particle i occupies
 * X[i*d .. i*d+d-1].  All arithmetic is IEEE double like the reference.
 *
 * All file:line citations are relative to /root/reference/include/SVGDCpp/.
 */
#ifndef SVGD_ORACLE_H
#define SVGD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_OPT_ADAGRAD = 0, ORACLE_OPT_ADAM = 1, ORACLE_OPT_RMSPROP = 2 };
enum { ORACLE_SCALE_MEDIAN = 0, ORACLE_SCALE_HESSIAN = 1, ORACLE_SCALE_FIXED = 2 };

/* Eigen::MatrixXd::Random(rows, cols) * scale with the unseeded glibc rand() the
 * reference examples rely on (examples/multivariate_normal/mvn_example.cpp:23). */
void oracle_eigen_random(double *out, size_t count, double scale, int reseed, unsigned seed);

/* Dense inverse by LU with partial pivoting: stand-in for Eigen's .inverse() on a
 * dynamic matrix (Model/MultivariateNormal.hpp:59).  Returns 0, or -1 if singular. */
int oracle_lu_inverse(const double *A, int d, double *Ainv);

/* Kernel/GaussianRBFKernel.hpp:222-254.  Mutates v (like nth_element). */
double oracle_median(double *v, size_t n);

/* Kernel/GaussianRBFKernel.hpp:168-188: a = log(n) / median(sqrt(D2))^2 over all n*n
 * ordered pairs.  work must hold n*n doubles (or NULL: allocated inside). */
double oracle_rbf_median_scale(const double *X, long n, int d, double *work);

/* grad log p for an unweighted, unnormalised sum of C Gaussians
 * (Model/MultivariateNormal.hpp:56-61, Model/Model.hpp:55-92,451-454).
 * means: C x d, covs: C x d x d (each symmetric).  lse != 0 evaluates the same
 * quantity through log-sum-exp (finite where the reference's log(exp()) underflows).
 * Returns 0 or -1 (singular covariance). */
int oracle_mvn_sum_logp_grad(const double *X, long n, int d, int C, const double *means,
                             const double *covs, int lse, double *G);

/* log p(x_i) of the same model (Model.hpp:305-308); lse as above. */
int oracle_mvn_sum_logp(const double *X, long n, int d, int C, const double *means, const double *covs, int lse, double *logp);

/* SVGD.hpp:407-454, literal eq. 8 double loop with k = exp(-a |x_j - x_i|^2)
 * (Kernel/GaussianRBFKernel.hpp:75-81) and grad_{x_j} k = -2 a (x_j - x_i) k. */
void oracle_phi(const double *X, const double *G, long n, int d, double a, double *phi);

/* GaussianRBFKernel::ComputeScale, ScaleMethod::Hessian (Kernel/GaussianRBFKernel.hpp:189-210):
 * A = 1/(2 d n) sum_i -Hessian(log p)(x_i), a full d x d matrix.  The reference tapes the Hessian with CppAD
 * (Model.hpp:366-370); for the sum of Gaussians it is, in closed form, with y_c = P_c (x - mu_c) and softmax weights w_c,
 *   -Hessian(log p) = sum_c w_c P_c - sum_c w_c y_c y_c^T + ybar ybar^T,  ybar = sum_c w_c y_c.
 * Parity of this branch is NOT pinned by any reference output: the closed form is checked against
 * finite differences of oracle_mvn_sum_logp_grad in tests/test_oracle_golden.py. */
int oracle_rbf_hessian_scale(const double *X, long n, int d, int C, const double *means, const double *covs, int lse,
                             double *A);

/* SVGD::ComputePhi with a matrix-valued kernel scale: k(x, x') = exp(-(x - x')^T A (x - x')),
 * grad_x k = -(A + A^T)(x - x') k  (Kernel/GaussianRBFKernel.hpp:75-81 taped by CppAD). */
void oracle_phi_matrix(const double *X, const double *G, long n, int d, const double *A, double *phi);

/* The intermediate matrices of SVGD::ComputePhi (SVGD.hpp:434-448) that SVGDOptions::LogIntermediateMatrices prints (:346-365):
 * K[i n + j] = kernel_matrix_(j, i), dK[(i n + j) d + c] = kernel_grad_matrix_(j d + c, i); A is the d x d scale matrix. */
void oracle_kernel_matrices(const double *X, long n, int d, const double *A, double *K, double *dK);

/* Optimizer increments (the driver ADDS the result, SVGD.hpp:393).
 * Optimizer/Adam.hpp:75-96, AdaGrad.hpp:60-65, RMSProp.hpp:69-74.
 * state1 = sum of squares / 2nd moment, state2 = 1st moment (Adam only). */
void oracle_opt_step(int kind, size_t count, const double *phi, double lr, double beta1,
                     double beta2, double eps, uint64_t *counter, double *state1,
                     double *state2, double *delta);

/* SVGD.hpp:396-399: X = max(min(X, ub), lb) per coordinate (lb/ub length d). */
void oracle_clamp(double *X, long n, int d, const double *lb, const double *ub);

typedef struct {
    long n;
    int d;
    int iters;
    int n_components;       /* C >= 1 */
    const double *means;    /* C x d */
    const double *covs;     /* C x d x d */
    int lse;                /* evaluate mixture gradient through log-sum-exp */
    int scale_method;       /* ORACLE_SCALE_MEDIAN | ORACLE_SCALE_HESSIAN | ORACLE_SCALE_FIXED */
    double fixed_a;
    int opt_kind;
    double lr, beta1, beta2, eps;
    const double *lb, *ub;  /* NULL => unchecked */
} oracle_config;

/* SVGD::Initialize + Run (SVGD.hpp:268-296, 338-400).  X is updated in place.
 * a_trace (iters doubles) and phi_last (n*d) may be NULL.  Returns 0 / -1. */
int oracle_svgd_run(const oracle_config *cfg, double *X, double *a_trace, double *phi_last);

/* ---- timed CPU baselines (bench.py cpu_baseline / --impl reference only) ---------- */

/* (R) reference-shaped iteration: omp parallel for over i (SVGD.hpp:418-431),
 * K (n x n) and grad K ((d n) x n) materialised, the two GEMMs of SVGD.hpp:453
 * including the [I I .. I] indexer product, nth_element median over n*n distances.
 * Scratch is allocated inside (d*n*n doubles!).  Returns seconds spent, or <0. */
double oracle_refshape_iterations_omp(const oracle_config *cfg, double *X, int iters, int threads);

/* (O) optimised CPU: blocked Gram form, nothing materialised except the n*n median
 * vector.  Same results to rounding.  Returns seconds spent, or <0. */
double oracle_blocked_iterations_omp(const oracle_config *cfg, double *X, int iters, int threads);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
