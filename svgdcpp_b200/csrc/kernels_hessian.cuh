// kernels_hessian.cuh — ScaleMethod::Hessian of the Gaussian RBF kernel (Kernel/GaussianRBFKernel.hpp:189-210):
//   A = 1/(2 d n) sum_i -Hessian(log p)(x_i),   k(x, x') = exp(-(x - x')^T A (x - x')),   grad_x k = -2 A (x - x') k.
//
// The pair-interaction kernels are reused unchanged through a change of variables.  With A = R^T R (Cholesky, R upper
// triangular), y = R x and g^ = R^-T g:
//   phi_i = (1/n) sum_j [ k_ji g_j - 2 A (x_j - x_i) k_ji ] = R^T phi^_i,
//   phi^_i = (1/n) [ sum_j k_ji (g^_j - 2 y_j) + 2 y_i sum_j k_ji ],   k_ji = exp(-|y_j - y_i|^2)
// i.e. exactly the scalar-bandwidth form with a = 1 on the transformed particles.  So a Hessian-scaled step is:
// Hessian sum (this file) -> A, R on the host (d x d) -> Y = X R^T, G^ = G R^-1 (row_times_matrix) -> the existing
// pair kernel in its "write phi" mode -> phi = phi^ R -> optimizer / clamp on the original particles (opt_apply).
//
// The Hessian of log p has a closed form for the models with a device gradient (sum of Gaussians, Model.hpp:55-92):
// with y_c = P_c (x - mu_c) and softmax weights w_c of -q_c/2,
//   -Hessian(log p) = sum_c w_c P_c - sum_c w_c y_c y_c^T + ybar ybar^T,   ybar = sum_c w_c y_c
// (the reference tapes it with CppAD, Model.hpp:366-370).  User models behind the gradient hook have no Hessian here.
#pragma once
#include "kernels_f64.cuh"

namespace svgdb {

// Partial sums over this rank's particles [row0, row0 + n_rows):
//   Hsum[a][b] += sum_i ( ybar_a ybar_b - sum_c w_ic y_ica y_icb ),   Wsum[c] += sum_i w_ic
// (the sum_c w_c P_c term is applied on the host from Wsum).  grid.x = particle chunks, grid.y = 64 x 64 tiles of (a, b);
// 256 threads, each owning 16 elements of the tile.  Dynamic shared memory: (C + 1) * d + 2 * C doubles.
__global__ void __launch_bounds__(256)
mvn_sum_hessian_f64_kernel(const double *__restrict__ X, int d, int64_t row0, int64_t n_rows, int C, const double *__restrict__ means,
                           const double *__restrict__ prec, int tiles_per_dim, double *__restrict__ Hsum, double *__restrict__ Wsum)
{
    extern __shared__ double sh[];
    double *y = sh;              // [C][d]
    double *ybar = y + (size_t)C * d; // [d]
    double *h = ybar + d;        // [C]  -q_c/2, then the weights
    double *wacc = h + C;        // [C]  this block's sum of weights
    const int ta = blockIdx.y / tiles_per_dim, tb = blockIdx.y % tiles_per_dim;
    const int t = threadIdx.x;
    double acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
    for (int c = t; c < C; c += blockDim.x) wacc[c] = 0.0;
    __syncthreads();
    for (int64_t li = blockIdx.x; li < n_rows; li += gridDim.x) {
        const double *x = X + (row0 + li) * d;
        for (int idx = t; idx < C * d; idx += blockDim.x) { // y_c = P_c (x - mu_c)
            const int c = idx / d, r = idx - c * d;
            const double *P = prec + ((size_t)c * d + r) * d, *mu = means + (size_t)c * d;
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += P[k] * (x[k] - mu[k]);
            y[idx] = s;
        }
        __syncthreads();
        if (t < C) { // -q_c / 2
            const double *mu = means + (size_t)t * d;
            double q = 0.0;
            for (int r = 0; r < d; ++r) q += (x[r] - mu[r]) * y[t * d + r];
            h[t] = -0.5 * q;
        }
        __syncthreads();
        if (t == 0) { // softmax through log-sum-exp (finite where the literal form underflows, like the gradient kernel)
            double shift = h[0];
            for (int c = 1; c < C; ++c) shift = fmax(shift, h[c]);
            double tot = 0.0;
            for (int c = 0; c < C; ++c) { h[c] = exp(h[c] - shift); tot += h[c]; }
            for (int c = 0; c < C; ++c) { h[c] /= tot; wacc[c] += h[c]; }
        }
        __syncthreads();
        for (int r = t; r < d; r += blockDim.x) {
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += h[c] * y[c * d + r];
            ybar[r] = s;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int el = t + 256 * e, a = ta * 64 + (el >> 6), b = tb * 64 + (el & 63);
            if (a < d && b < d) {
                double s = ybar[a] * ybar[b];
                for (int c = 0; c < C; ++c) s -= h[c] * y[c * d + a] * y[c * d + b];
                acc[e] += s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int el = t + 256 * e, a = ta * 64 + (el >> 6), b = tb * 64 + (el & 63);
        if (a < d && b < d && acc[e] != 0.0) atomicAdd(&Hsum[(size_t)a * d + b], acc[e]);
    }
    if (blockIdx.y == 0)
        for (int c = t; c < C; c += blockDim.x)
            if (wacc[c] != 0.0) atomicAdd(&Wsum[c], wacc[c]);
}

// out[i][c] = sum_k in[i][k] M[k][c]  for rows [0, n_rows); M is d x d row-major (d <= a few hundred: served by L1/L2).
__global__ void row_times_matrix_f64_kernel(const double *__restrict__ in, const double *__restrict__ M, int64_t n_rows, int d,
                                            double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * d) return;
    const int64_t i = idx / d;
    const int c = (int)(idx - i * d);
    const double *row = in + i * d;
    double s = 0.0;
    for (int k = 0; k < d; ++k) s += row[k] * M[(size_t)k * d + c];
    out[idx] = s;
}

// X_out = clamp(X + optimizer(phi)) for this rank's rows (Optimizer/*.hpp, SVGD.hpp:393-399); phi is indexed by local row.
__global__ void opt_apply_f64_kernel(const double *__restrict__ X, const double *__restrict__ phi, int64_t row0, int64_t n_rows, int d,
                                     OptParams opt, double *__restrict__ s1, double *__restrict__ s2, const double *__restrict__ lb,
                                     const double *__restrict__ ub, double *__restrict__ X_out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * d) return;
    const int c = (int)(idx % d);
    const double xn = X[row0 * d + idx] + opt_increment(opt, phi[idx], s1, s2, idx);
    X_out[row0 * d + idx] = clamp_coord(xn, lb, ub, c);
}

} // namespace svgdb
