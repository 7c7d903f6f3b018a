/*
 * svgd_oracle.c — CPU restatement of the SVGDCpp hot path.  TEST INFRASTRUCTURE ONLY
 * (see svgd_oracle.h).  Plain C, IEEE double, loop-for-loop after the reference; the
 * derivatives CppAD would produce are written in closed form.
 * Citations are relative to /root/reference/include/SVGDCpp/.
 */
#include "svgd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Eigen's random<double>() is x + (y-x)*rand()/RAND_MAX with x=-1, y=1, filled in
 * storage (column-major) order; `3 * Random(dim, n)` then scales it
 * (examples/multivariate_normal/mvn_example.cpp:23). */
void oracle_eigen_random(double *out, size_t count, double scale, int reseed, unsigned seed)
{
    if (reseed) srand(seed);
    for (size_t t = 0; t < count; ++t) {
        double r = (double)rand();
        out[t] = scale * (-1.0 + 2.0 * r / (double)RAND_MAX);
    }
}

int oracle_lu_inverse(const double *A, int d, double *Ainv)
{
    double *lu = (double *)malloc(sizeof(double) * d * d);
    int *piv = (int *)malloc(sizeof(int) * d);
    if (!lu || !piv) { free(lu); free(piv); return -1; }
    memcpy(lu, A, sizeof(double) * d * d); /* row-major d x d (symmetric inputs: either) */
    for (int i = 0; i < d; ++i) piv[i] = i;
    for (int k = 0; k < d; ++k) {
        int p = k;
        double best = fabs(lu[k * d + k]);
        for (int r = k + 1; r < d; ++r)
            if (fabs(lu[r * d + k]) > best) { best = fabs(lu[r * d + k]); p = r; }
        if (best == 0.0) { free(lu); free(piv); return -1; }
        if (p != k) {
            for (int c = 0; c < d; ++c) { double t = lu[k * d + c]; lu[k * d + c] = lu[p * d + c]; lu[p * d + c] = t; }
            int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
        }
        for (int r = k + 1; r < d; ++r) {
            double f = lu[r * d + k] / lu[k * d + k];
            lu[r * d + k] = f;
            for (int c = k + 1; c < d; ++c) lu[r * d + c] -= f * lu[k * d + c];
        }
    }
    /* solve for each unit vector */
    double *y = (double *)malloc(sizeof(double) * d);
    for (int col = 0; col < d; ++col) {
        for (int r = 0; r < d; ++r) {
            double s = (piv[r] == col) ? 1.0 : 0.0;
            for (int c = 0; c < r; ++c) s -= lu[r * d + c] * y[c];
            y[r] = s;
        }
        for (int r = d - 1; r >= 0; --r) {
            double s = y[r];
            for (int c = r + 1; c < d; ++c) s -= lu[r * d + c] * Ainv[c * d + col];
            Ainv[r * d + col] = s / lu[r * d + r];
        }
    }
    free(y); free(lu); free(piv);
    return 0;
}

/* nth_element stand-in: after the call v[k] is the k-th order statistic, everything
 * before it is <= and everything after it is >= (the property ComputeMedian relies on). */
static void select_kth(double *v, size_t n, size_t k)
{
    ptrdiff_t lo = 0, hi = (ptrdiff_t)n - 1, kk = (ptrdiff_t)k;
    while (lo < hi) {
        double pivot = v[lo + (hi - lo) / 2];
        ptrdiff_t i = lo, j = hi;
        while (i <= j) {
            while (v[i] < pivot) ++i;
            while (v[j] > pivot) --j;
            if (i <= j) {
                double t = v[i]; v[i] = v[j]; v[j] = t;
                ++i; --j;
            }
        }
        /* v[lo..j] <= pivot <= v[i..hi], j < i */
        if (kk <= j) hi = j;
        else if (kk >= i) lo = i;
        else return;
    }
}

/* Kernel/GaussianRBFKernel.hpp:222-254 */
double oracle_median(double *v, size_t n)
{
    if (n % 2 == 0) {
        size_t h = n / 2;
        select_kth(v, n, h);
        double b = v[h];
        double a = v[0];
        for (size_t t = 1; t < h; ++t) if (v[t] > a) a = v[t]; /* max_element of the lower half */
        return (a + b) / 2.0;
    }
    size_t h = n / 2;
    select_kth(v, n, h);
    return v[h];
}

/* Kernel/GaussianRBFKernel.hpp:179-187 */
double oracle_rbf_median_scale(const double *X, long n, int d, double *work)
{
    double *dist = work ? work : (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    double *diag = (double *)malloc(sizeof(double) * (size_t)n);
    if (!dist || !diag) { if (!work) free(dist); free(diag); return NAN; }
    for (long i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += X[i * d + k] * X[i * d + k];
        diag[i] = s; /* squared_coord_matrix_.diagonal() */
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        for (long j = 0; j < n; ++j) {
            double s;
            if (i == j) s = diag[i]; /* the Gram diagonal IS the replicated diagonal: D2 = 0 exactly */
            else {
                s = 0.0;
                for (int k = 0; k < d; ++k) s += X[i * d + k] * X[j * d + k];
            }
            double d2 = diag[i] + diag[j] - 2.0 * s;
            dist[(size_t)i * n + j] = sqrt(d2);
        }
    }
    double med = oracle_median(dist, (size_t)n * (size_t)n);
    double a = log((double)n) / pow(med, 2);
    free(diag);
    if (!work) free(dist);
    return a;
}

int oracle_mvn_sum_logp_grad(const double *X, long n, int d, int C, const double *means,
                             const double *covs, int lse, double *G)
{
    double *prec = (double *)malloc(sizeof(double) * (size_t)C * d * d);
    if (!prec) return -1;
    for (int c = 0; c < C; ++c)
        if (oracle_lu_inverse(covs + (size_t)c * d * d, d, prec + (size_t)c * d * d)) { free(prec); return -1; }
    int rc = 0;
#pragma omp parallel
    {
        double *y = (double *)malloc(sizeof(double) * (size_t)C * d); /* P_c (x - mu_c) */
        double *h = (double *)malloc(sizeof(double) * (size_t)C);     /* -q_c / 2 */
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            const double *x = X + i * d;
            for (int c = 0; c < C; ++c) {
                const double *P = prec + (size_t)c * d * d;
                const double *mu = means + (size_t)c * d;
                double q = 0.0;
                for (int r = 0; r < d; ++r) {
                    double s = 0.0;
                    for (int k = 0; k < d; ++k) s += P[r * d + k] * (x[k] - mu[k]);
                    y[c * d + r] = s;
                    q += (x[r] - mu[r]) * s;
                }
                h[c] = -0.5 * q;
            }
            /* log p = log sum_c exp(h_c); grad = sum_c w_c (-y_c), w = exp(h_c)/sum exp(h) */
            double shift = 0.0;
            if (lse) { shift = h[0]; for (int c = 1; c < C; ++c) if (h[c] > shift) shift = h[c]; }
            double tot = 0.0;
            for (int c = 0; c < C; ++c) { h[c] = exp(h[c] - shift); tot += h[c]; }
            for (int r = 0; r < d; ++r) {
                double s = 0.0;
                for (int c = 0; c < C; ++c) s += h[c] * (-y[c * d + r]);
                G[i * d + r] = s / tot; /* NaN when every exp underflowed and lse == 0, like the reference */
            }
        }
        free(y); free(h);
    }
    free(prec);
    return rc;
}

/* Model::EvaluateLogModel (Model.hpp:305-308) of the sum of unnormalised Gaussians (MultivariateNormal.hpp:56-61, Model.hpp:55-92). */
int oracle_mvn_sum_logp(const double *X, long n, int d, int C, const double *means, const double *covs, int lse, double *logp)
{
    double *prec = (double *)malloc(sizeof(double) * (size_t)C * d * d);
    double *h = (double *)malloc(sizeof(double) * (size_t)C);
    if (!prec || !h) { free(prec); free(h); return -1; }
    for (int c = 0; c < C; ++c)
        if (oracle_lu_inverse(covs + (size_t)c * d * d, d, prec + (size_t)c * d * d)) { free(prec); free(h); return -1; }
    for (long i = 0; i < n; ++i) {
        const double *x = X + i * d;
        for (int c = 0; c < C; ++c) {
            const double *P = prec + (size_t)c * d * d, *mu = means + (size_t)c * d;
            double q = 0.0;
            for (int r = 0; r < d; ++r) {
                double s = 0.0;
                for (int k = 0; k < d; ++k) s += P[r * d + k] * (x[k] - mu[k]);
                q += (x[r] - mu[r]) * s;
            }
            h[c] = -0.5 * q;
        }
        double shift = 0.0, tot = 0.0;
        if (lse) { shift = h[0]; for (int c = 1; c < C; ++c) if (h[c] > shift) shift = h[c]; }
        for (int c = 0; c < C; ++c) tot += exp(h[c] - shift);
        logp[i] = shift + log(tot); /* -inf when every exp underflowed and lse == 0, like the reference */
    }
    free(prec); free(h);
    return 0;
}

/* SVGD.hpp:435-453 */
void oracle_phi(const double *X, const double *G, long n, int d, double a, double *phi)
{
#pragma omp parallel
    {
        double *acc_gk = (double *)malloc(sizeof(double) * d);
        double *acc_dk = (double *)malloc(sizeof(double) * d);
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            const double *xi = X + i * d; /* kernel location (SVGD.hpp:441) */
            for (int k = 0; k < d; ++k) { acc_gk[k] = 0.0; acc_dk[k] = 0.0; }
            for (long j = 0; j < n; ++j) {
                const double *xj = X + j * d;
                double q = 0.0;
                for (int k = 0; k < d; ++k) { double df = xj[k] - xi[k]; q += df * (a * df); }
                double kv = exp(-q);                       /* kernel_matrix_(j, i) */
                for (int k = 0; k < d; ++k) {
                    acc_gk[k] += G[j * d + k] * kv;        /* log_model_grad_matrix_ * kernel_matrix_ */
                    acc_dk[k] += -2.0 * a * (xj[k] - xi[k]) * kv; /* indexer * kernel_grad_matrix_ */
                }
            }
            for (int k = 0; k < d; ++k) phi[i * d + k] = (1.0 / (double)n) * (acc_gk[k] + acc_dk[k]);
        }
        free(acc_gk); free(acc_dk);
    }
}

int oracle_rbf_hessian_scale(const double *X, long n, int d, int C, const double *means, const double *covs, int lse,
                             double *A)
{
    double *prec = (double *)malloc(sizeof(double) * (size_t)C * d * d);
    double *y = (double *)malloc(sizeof(double) * (size_t)C * d);
    double *h = (double *)malloc(sizeof(double) * (size_t)C);
    double *ybar = (double *)malloc(sizeof(double) * (size_t)d);
    if (!prec || !y || !h || !ybar) { free(prec); free(y); free(h); free(ybar); return -1; }
    for (int c = 0; c < C; ++c)
        if (oracle_lu_inverse(covs + (size_t)c * d * d, d, prec + (size_t)c * d * d)) { free(prec); free(y); free(h); free(ybar); return -1; }
    for (size_t t = 0; t < (size_t)d * d; ++t) A[t] = 0.0;
    for (long i = 0; i < n; ++i) { /* GaussianRBFKernel.hpp:203-206: hessian_sum += -EvaluateLogModelHessian(x_i) */
        const double *x = X + i * d;
        for (int c = 0; c < C; ++c) {
            const double *P = prec + (size_t)c * d * d, *mu = means + (size_t)c * d;
            double q = 0.0;
            for (int r = 0; r < d; ++r) {
                double s = 0.0;
                for (int k = 0; k < d; ++k) s += P[r * d + k] * (x[k] - mu[k]);
                y[c * d + r] = s;
                q += (x[r] - mu[r]) * s;
            }
            h[c] = -0.5 * q;
        }
        double shift = 0.0, tot = 0.0;
        if (lse) { shift = h[0]; for (int c = 1; c < C; ++c) if (h[c] > shift) shift = h[c]; }
        for (int c = 0; c < C; ++c) { h[c] = exp(h[c] - shift); tot += h[c]; }
        for (int r = 0; r < d; ++r) {
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += (h[c] / tot) * y[c * d + r];
            ybar[r] = s;
        }
        for (int r = 0; r < d; ++r)
            for (int k = 0; k < d; ++k) {
                double s = ybar[r] * ybar[k];
                for (int c = 0; c < C; ++c) {
                    /* the symmetric part of P_c is what a Hessian sees */
                    const double *P = prec + (size_t)c * d * d;
                    s += (h[c] / tot) * (0.5 * (P[r * d + k] + P[k * d + r]) - y[c * d + r] * y[c * d + k]);
                }
                A[r * d + k] += s;
            }
    }
    for (size_t t = 0; t < (size_t)d * d; ++t) A[t] *= 1.0 / (2.0 * (double)d * (double)n); /* :208 */
    free(prec); free(y); free(h); free(ybar);
    return 0;
}

void oracle_phi_matrix(const double *X, const double *G, long n, int d, const double *A, double *phi)
{
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * d);
        double *df = (double *)malloc(sizeof(double) * d);
        double *Ad = (double *)malloc(sizeof(double) * d);
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            const double *xi = X + i * d;
            for (int k = 0; k < d; ++k) acc[k] = 0.0;
            for (long j = 0; j < n; ++j) {
                const double *xj = X + j * d;
                double q = 0.0;
                for (int k = 0; k < d; ++k) df[k] = xj[k] - xi[k];
                for (int r = 0; r < d; ++r) { /* (A + A^T) diff, and diff^T A diff */
                    double s = 0.0, st = 0.0;
                    for (int k = 0; k < d; ++k) { s += A[r * d + k] * df[k]; st += A[k * d + r] * df[k]; }
                    Ad[r] = s + st;
                    q += df[r] * s;
                }
                double kv = exp(-q);
                for (int k = 0; k < d; ++k) acc[k] += G[j * d + k] * kv - Ad[k] * kv;
            }
            for (int k = 0; k < d; ++k) phi[i * d + k] = (1.0 / (double)n) * acc[k];
        }
        free(acc); free(df); free(Ad);
    }
}

void oracle_kernel_matrices(const double *X, long n, int d, const double *A, double *K, double *dK)
{
    /* SVGD.hpp:434-448: kernel_matrix_(j, i) = k(x_j, x_i), kernel_grad_matrix_.block(j d, i, d, 1) = grad k(x_j, x_i), with the
     * kernel located at x_i (Kernel/GaussianRBFKernel.hpp:75-81): k = exp(-diff^T A diff), grad = -(A + A^T) diff k. */
    double *df = (double *)malloc(sizeof(double) * d);
    for (long i = 0; i < n; ++i) {
        for (long j = 0; j < n; ++j) {
            double q = 0.0;
            for (int k = 0; k < d; ++k) df[k] = X[j * d + k] - X[i * d + k];
            for (int r = 0; r < d; ++r) {
                double s = 0.0;
                for (int k = 0; k < d; ++k) s += A[r * d + k] * df[k];
                q += df[r] * s;
            }
            double kv = exp(-q);
            K[i * n + j] = kv;
            for (int r = 0; r < d; ++r) {
                double s = 0.0, st = 0.0;
                for (int k = 0; k < d; ++k) { s += A[r * d + k] * df[k]; st += A[k * d + r] * df[k]; }
                dK[(i * n + j) * d + r] = -(s + st) * kv;
            }
        }
    }
    free(df);
}

void oracle_opt_step(int kind, size_t count, const double *phi, double lr, double beta1,
                     double beta2, double eps, uint64_t *counter, double *s1, double *s2,
                     double *delta)
{
    if (kind == ORACLE_OPT_ADAM) { /* Optimizer/Adam.hpp:75-96 */
        for (size_t t = 0; t < count; ++t) {
            s2[t] = beta1 * s2[t] + (1 - beta1) * phi[t];
            s1[t] = beta2 * s1[t] + (1 - beta2) * (phi[t] * phi[t]);
        }
        ++*counter;
        double c1 = 1.0 - pow(beta1, (double)*counter);
        double c2 = 1.0 - pow(beta2, (double)*counter);
        for (size_t t = 0; t < count; ++t)
            delta[t] = lr * (1.0 / (eps + sqrt(s1[t] / c2))) * (s2[t] / c1);
    } else if (kind == ORACLE_OPT_ADAGRAD) { /* Optimizer/AdaGrad.hpp:60-65 */
        for (size_t t = 0; t < count; ++t) {
            s1[t] += phi[t] * phi[t];
            delta[t] = lr * (1.0 / (eps + sqrt(s1[t]))) * phi[t];
        }
    } else { /* Optimizer/RMSProp.hpp:69-74 (beta1 is the decay) */
        for (size_t t = 0; t < count; ++t) {
            s1[t] = beta1 * s1[t] + (1 - beta1) * (phi[t] * phi[t]);
            delta[t] = lr * (1.0 / (eps + sqrt(s1[t]))) * phi[t];
        }
    }
}

void oracle_clamp(double *X, long n, int d, const double *lb, const double *ub)
{
    for (long i = 0; i < n; ++i)
        for (int k = 0; k < d; ++k) {
            double v = X[i * d + k];
            v = v < ub[k] ? v : ub[k]; /* .min(upper) first ... */
            v = v > lb[k] ? v : lb[k]; /* ... then .max(lower), SVGD.hpp:398 */
            X[i * d + k] = v;
        }
}

int oracle_svgd_run(const oracle_config *cfg, double *X, double *a_trace, double *phi_last)
{
    long n = cfg->n;
    int d = cfg->d;
    size_t cnt = (size_t)n * d;
    double *G = (double *)malloc(sizeof(double) * cnt);
    double *phi = (double *)malloc(sizeof(double) * cnt);
    double *delta = (double *)malloc(sizeof(double) * cnt);
    double *s1 = (double *)calloc(cnt, sizeof(double)); /* Optimizer::Initialize() zeroes state */
    double *s2 = (double *)calloc(cnt, sizeof(double));
    double *work = cfg->scale_method == ORACLE_SCALE_MEDIAN ? (double *)malloc(sizeof(double) * (size_t)n * n) : NULL;
    uint64_t counter = 0;
    int rc = 0;
    double *Amat = NULL;
    if (!G || !phi || !delta || !s1 || !s2) rc = -1;
    for (int it = 0; it < cfg->iters && !rc; ++it) {
        /* SVGD::Step: kernel Step (bandwidth from the CURRENT X) happens before ComputePhi, SVGD.hpp:378-393 */
        double a = cfg->scale_method == ORACLE_SCALE_MEDIAN ? oracle_rbf_median_scale(X, n, d, work) : cfg->fixed_a;
        if (cfg->scale_method == ORACLE_SCALE_HESSIAN) {
            if (!Amat) Amat = (double *)malloc(sizeof(double) * (size_t)d * d);
            rc = Amat ? oracle_rbf_hessian_scale(X, n, d, cfg->n_components, cfg->means, cfg->covs, cfg->lse, Amat) : -1;
            if (rc) break;
            a = Amat[0];
        }
        if (a_trace) a_trace[it] = a;
        rc = oracle_mvn_sum_logp_grad(X, n, d, cfg->n_components, cfg->means, cfg->covs, cfg->lse, G);
        if (rc) break;
        if (cfg->scale_method == ORACLE_SCALE_HESSIAN) oracle_phi_matrix(X, G, n, d, Amat, phi);
        else oracle_phi(X, G, n, d, a, phi);
        oracle_opt_step(cfg->opt_kind, cnt, phi, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, &counter, s1, s2, delta);
        for (size_t t = 0; t < cnt; ++t) X[t] += delta[t];
        if (cfg->lb && cfg->ub) oracle_clamp(X, n, d, cfg->lb, cfg->ub);
    }
    if (phi_last && !rc) memcpy(phi_last, phi, sizeof(double) * cnt);
    free(G); free(phi); free(delta); free(s1); free(s2); free(work); free(Amat);
    return rc;
}

/* ------------------------------------------------------------------------------------
 * Timed CPU baselines.
 * ---------------------------------------------------------------------------------- */

double oracle_refshape_iterations_omp(const oracle_config *cfg, double *X, int iters, int threads)
{
    long n = cfg->n;
    int d = cfg->d;
    size_t cnt = (size_t)n * d;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double *G = (double *)malloc(sizeof(double) * cnt);
    double *K = (double *)malloc(sizeof(double) * (size_t)n * n);        /* kernel_matrix_(j, i) at K[i*n + j] */
    double *dK = (double *)malloc(sizeof(double) * (size_t)n * n * d);   /* kernel_grad_matrix_ column i at dK[i*(n*d) ..] */
    double *work = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *phi = (double *)malloc(sizeof(double) * cnt);
    double *delta = (double *)malloc(sizeof(double) * cnt);
    double *s1 = (double *)calloc(cnt, sizeof(double));
    double *s2 = (double *)calloc(cnt, sizeof(double));
    uint64_t counter = 0;
    if (!G || !K || !dK || !work || !phi || !delta || !s1 || !s2) {
        free(G); free(K); free(dK); free(work); free(phi); free(delta); free(s1); free(s2);
        return -1.0;
    }
    double t0 = now_s();
    for (int it = 0; it < iters; ++it) {
        double a = cfg->scale_method == ORACLE_SCALE_MEDIAN ? oracle_rbf_median_scale(X, n, d, work) : cfg->fixed_a;
        oracle_mvn_sum_logp_grad(X, n, d, cfg->n_components, cfg->means, cfg->covs, cfg->lse, G);
#pragma omp parallel for schedule(static) /* SVGD.hpp:418 */
        for (long i = 0; i < n; ++i) {
            const double *xi = X + i * d;
            for (long j = 0; j < n; ++j) {
                const double *xj = X + j * d;
                double q = 0.0;
                for (int k = 0; k < d; ++k) { double df = xj[k] - xi[k]; q += df * (a * df); }
                double kv = exp(-q);
                K[(size_t)i * n + j] = kv;
                double *g = dK + ((size_t)i * n + j) * d;
                for (int k = 0; k < d; ++k) g[k] = -2.0 * a * (xj[k] - xi[k]) * kv;
            }
        }
        /* (1/n) (G K + [I I .. I] dK): the indexer product multiplies by the dense
         * d x (d n) matrix of stacked identities, zeros included (SVGD.hpp:181,453). */
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            const double *Ki = K + (size_t)i * n;
            const double *dKi = dK + (size_t)i * n * d;
            for (int r = 0; r < d; ++r) {
                double s = 0.0;
                for (long j = 0; j < n; ++j) s += G[j * d + r] * Ki[j];
                double s2g = 0.0;
                for (long j = 0; j < n; ++j)
                    for (int k = 0; k < d; ++k) s2g += ((k == r) ? 1.0 : 0.0) * dKi[j * d + k];
                phi[i * d + r] = (1.0 / (double)n) * (s + s2g);
            }
        }
        oracle_opt_step(cfg->opt_kind, cnt, phi, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, &counter, s1, s2, delta);
        for (size_t t = 0; t < cnt; ++t) X[t] += delta[t];
        if (cfg->lb && cfg->ub) oracle_clamp(X, n, d, cfg->lb, cfg->ub);
    }
    double t1 = now_s();
    free(G); free(K); free(dK); free(work); free(phi); free(delta); free(s1); free(s2);
    return t1 - t0;
}

double oracle_blocked_iterations_omp(const oracle_config *cfg, double *X, int iters, int threads)
{
    long n = cfg->n;
    int d = cfg->d;
    size_t cnt = (size_t)n * d;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double *G = (double *)malloc(sizeof(double) * cnt);
    double *V = (double *)malloc(sizeof(double) * cnt);
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *work = cfg->scale_method == ORACLE_SCALE_MEDIAN ? (double *)malloc(sizeof(double) * (size_t)n * n) : NULL;
    double *phi = (double *)malloc(sizeof(double) * cnt);
    double *delta = (double *)malloc(sizeof(double) * cnt);
    double *s1 = (double *)calloc(cnt, sizeof(double));
    double *s2 = (double *)calloc(cnt, sizeof(double));
    uint64_t counter = 0;
    if (!G || !V || !r || !phi || !delta || !s1 || !s2) return -1.0;
    double t0 = now_s();
    for (int it = 0; it < iters; ++it) {
        double a = cfg->scale_method == ORACLE_SCALE_MEDIAN ? oracle_rbf_median_scale(X, n, d, work) : cfg->fixed_a;
        oracle_mvn_sum_logp_grad(X, n, d, cfg->n_components, cfg->means, cfg->covs, cfg->lse, G);
        for (long i = 0; i < n; ++i) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) { s += X[i * d + k] * X[i * d + k]; V[i * d + k] = G[i * d + k] - 2.0 * a * X[i * d + k]; }
            r[i] = s;
        }
        /* phi_i = (1/n) [ sum_j K_ji v_j + 2 a x_i sum_j K_ji ],  K_ji = exp(-a (r_i + r_j - 2 x_i.x_j)) */
#pragma omp parallel for schedule(dynamic, 8)
        for (long i = 0; i < n; ++i) {
            const double *xi = X + i * d;
            double acc[512];
            double ksum = 0.0;
            for (int k = 0; k < d; ++k) acc[k] = 0.0;
            for (long j = 0; j < n; ++j) {
                const double *xj = X + j * d;
                double s = 0.0;
                for (int k = 0; k < d; ++k) s += xi[k] * xj[k];
                double d2 = r[i] + r[j] - 2.0 * s;
                double kv = exp(-a * (d2 > 0.0 ? d2 : 0.0));
                ksum += kv;
                const double *vj = V + j * d;
                for (int k = 0; k < d; ++k) acc[k] += kv * vj[k];
            }
            for (int k = 0; k < d; ++k) phi[i * d + k] = (acc[k] + 2.0 * a * xi[k] * ksum) / (double)n;
        }
        oracle_opt_step(cfg->opt_kind, cnt, phi, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, &counter, s1, s2, delta);
        for (size_t t = 0; t < cnt; ++t) X[t] += delta[t];
        if (cfg->lb && cfg->ub) oracle_clamp(X, n, d, cfg->lb, cfg->ub);
    }
    double t1 = now_s();
    free(G); free(V); free(r); free(work); free(phi); free(delta); free(s1); free(s2);
    return t1 - t0;
}
