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
// (the sum_c w_c P_c term is applied on the host from Wsum).  grid.x = groups of PT particles (grid-stride), grid.y = 64 x 64 tiles
// of (a, b); 256 threads, each owning 16 elements of the tile.  P_c is symmetric (up to the rounding of the host inversion), so y_c = P_c (x - mu_c)
// is read column-wise: consecutive threads read consecutive addresses, and one read of P serves the PT particles of the group.
// Dynamic shared memory: hessian_smem_doubles(PT, C, d) doubles.
__host__ __device__ inline size_t hessian_smem_doubles(int PT, int C, int d)
{
    return (size_t)PT * d * (C + 2) + (size_t)PT * C + (size_t)C;
}

template <int PT>
__global__ void __launch_bounds__(256)
mvn_sum_hessian_f64_kernel(const double *__restrict__ X, int d, int64_t row0, int64_t n_rows, int C, const double *__restrict__ means,
                           const double *__restrict__ prec, int tiles_per_dim, double *__restrict__ Hsum, double *__restrict__ Wsum)
{
    extern __shared__ double sh[];
    double *xs = sh;                          // [PT][d]     particles of the group
    double *y = xs + (size_t)PT * d;          // [PT][C][d]  y_c = P_c (x - mu_c)
    double *ybar = y + (size_t)PT * C * d;    // [PT][d]
    double *h = ybar + (size_t)PT * d;        // [PT][C]     -q_c / 2, then the softmax weights
    double *wacc = h + (size_t)PT * C;        // [C]         this block's sum of weights
    const int ta = blockIdx.y / tiles_per_dim, tb = blockIdx.y % tiles_per_dim;
    const int t = threadIdx.x;
    double acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
    for (int c = t; c < C; c += 256) wacc[c] = 0.0;
    const int64_t n_groups = (n_rows + PT - 1) / PT;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int np = (int)min((int64_t)PT, n_rows - g * PT);
        __syncthreads(); // the previous group's tile stage is done with xs / y / ybar / h
        for (int idx = t; idx < PT * d; idx += 256) xs[idx] = idx < np * d ? X[(row0 + g * PT) * d + idx] : 0.0;
        __syncthreads();
        for (int item = t; item < C * d; item += 256) {
            const int c = item / d, r = item - c * d;
            const double *P = prec + (size_t)c * d * d + r, *mu = means + (size_t)c * d;
            double s[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) s[p] = 0.0;
            for (int k = 0; k < d; ++k) {
                const double pk = P[(size_t)k * d], mk = mu[k];
#pragma unroll
                for (int p = 0; p < PT; ++p) s[p] += pk * (xs[p * d + k] - mk);
            }
#pragma unroll
            for (int p = 0; p < PT; ++p) y[((size_t)p * C + c) * d + r] = s[p];
        }
        __syncthreads();
        for (int item = t; item < PT * C; item += 256) { // -q_c / 2
            const int p = item / C, c = item - p * C;
            const double *mu = means + (size_t)c * d;
            double q = 0.0;
            for (int r = 0; r < d; ++r) q += (xs[p * d + r] - mu[r]) * y[((size_t)p * C + c) * d + r];
            h[item] = -0.5 * q;
        }
        __syncthreads();
        if (t < np) { // softmax through log-sum-exp (finite where the literal form underflows, like the gradient kernel)
            double *hp = h + (size_t)t * C;
            double shift = hp[0];
            for (int c = 1; c < C; ++c) shift = fmax(shift, hp[c]);
            double tot = 0.0;
            for (int c = 0; c < C; ++c) { hp[c] = exp(hp[c] - shift); tot += hp[c]; }
            for (int c = 0; c < C; ++c) hp[c] /= tot;
        }
        __syncthreads();
        if (blockIdx.y == 0)
            for (int c = t; c < C; c += 256)
                for (int p = 0; p < np; ++p) wacc[c] += h[(size_t)p * C + c];
        for (int item = t; item < PT * d; item += 256) {
            const int p = item / d, r = item - p * d;
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += h[(size_t)p * C + c] * y[((size_t)p * C + c) * d + r];
            ybar[item] = s;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int el = t + 256 * e, a = ta * 64 + (el >> 6), b = tb * 64 + (el & 63);
            if (a < d && b < d) {
                double s = 0.0;
                for (int p = 0; p < np; ++p) {
                    s += ybar[p * d + a] * ybar[p * d + b];
                    for (int c = 0; c < C; ++c) s -= h[(size_t)p * C + c] * y[((size_t)p * C + c) * d + a] * y[((size_t)p * C + c) * d + b];
                }
                acc[e] += s;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int el = t + 256 * e, a = ta * 64 + (el >> 6), b = tb * 64 + (el & 63);
        if (a < d && b < d && acc[e] != 0.0) atomicAdd(&Hsum[(size_t)a * d + b], acc[e]);
    }
    __syncthreads();
    if (blockIdx.y == 0)
        for (int c = t; c < C; c += 256)
            if (wacc[c] != 0.0) atomicAdd(&Wsum[c], wacc[c]);
}

// out[i][c] = sum_k in[i][k] M[k][c]  for rows [0, n_rows); M is d x d row-major (d <= a few hundred: served by L1/L2).
// One thread owns column c of 8 consecutive rows, so each M[k][c] is read once for 8 outputs and the row values are warp broadcasts.
// APPLY: instead of storing, the result is phi of local row i: X_out = clamp(X + optimizer(phi)) (Optimizer/*.hpp, SVGD.hpp:393-399).
struct RowApply {
    const double *X;
    double *X_out;
    int64_t row0;
    OptParams opt;
    double *s1, *s2;
    const double *lb, *ub;
};

template <bool APPLY>
__global__ void __launch_bounds__(256)
row_times_matrix_f64_kernel(const double *__restrict__ in, const double *__restrict__ M, int64_t n_rows, int d, double *__restrict__ out, RowApply ap)
{
    constexpr int PR = 8;
    const int64_t cols = ((int64_t)d + 31) & ~(int64_t)31; // a warp stays inside one row block
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t blk = idx / cols;
    const int c = (int)(idx - blk * cols);
    const int64_t i0 = blk * PR;
    if (i0 >= n_rows || c >= d) return;
    const int np = (int)min((int64_t)PR, n_rows - i0);
    double s[PR];
#pragma unroll
    for (int p = 0; p < PR; ++p) s[p] = 0.0;
    const double *row = in + i0 * d;
    if (np == PR) {
        for (int k = 0; k < d; ++k) {
            const double m = M[(size_t)k * d + c];
#pragma unroll
            for (int p = 0; p < PR; ++p) s[p] += row[(size_t)p * d + k] * m;
        }
    } else {
        for (int k = 0; k < d; ++k) {
            const double m = M[(size_t)k * d + c];
            for (int p = 0; p < np; ++p) s[p] += row[(size_t)p * d + k] * m;
        }
    }
#pragma unroll
    for (int p = 0; p < PR; ++p) {
        if (p >= np) break;
        const int64_t o = (i0 + p) * d + c;
        if (APPLY) {
            const double xn = ap.X[ap.row0 * d + o] + opt_increment(ap.opt, s[p], ap.s1, ap.s2, o);
            ap.X_out[ap.row0 * d + o] = clamp_coord(xn, ap.lb, ap.ub, c);
        } else {
            out[o] = s[p];
        }
    }
}

inline unsigned row_times_matrix_blocks(int64_t n_rows, int d)
{
    const int64_t cols = ((int64_t)d + 31) & ~(int64_t)31;
    return (unsigned)((((n_rows + 7) / 8) * cols + 255) / 256);
}

// Inspection path of SVGDOptions::LogIntermediateMatrices (SVGD.hpp:346-365, 407-454): the n x n kernel matrix and its gradients
// in the reference's own layouts, for small n.  One thread per pair (j, i):
//   K[i n + j] = k(x_j, x_i) = exp(-(x_j - x_i)^T A (x_j - x_i))              == kernel_matrix_(j, i), column-major
//   dK[(i n + j) d + c] = -((A + A^T)(x_j - x_i))_c k(x_j, x_i)               == kernel_grad_matrix_(j d + c, i), column-major
// (A = a I for the scalar scales).  The hot path never forms either matrix.
__global__ void kernel_matrices_f64_kernel(const double *__restrict__ X, int64_t n, int d, const double *__restrict__ A, double *__restrict__ K,
                                           double *__restrict__ dK)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int64_t i = idx / n, j = idx - i * n;
    const double *xi = X + i * d, *xj = X + j * d;
    double q = 0.0;
    for (int r = 0; r < d; ++r) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += A[(size_t)r * d + k] * (xj[k] - xi[k]);
        q += (xj[r] - xi[r]) * s;
    }
    const double kv = exp(-q);
    K[idx] = kv;
    for (int r = 0; r < d; ++r) {
        double s = 0.0, st = 0.0;
        for (int k = 0; k < d; ++k) {
            const double df = xj[k] - xi[k];
            s += A[(size_t)r * d + k] * df;
            st += A[(size_t)k * d + r] * df;
        }
        dK[idx * d + r] = -(s + st) * kv;
    }
}

// log p(x_i) of a sum of C unnormalised Gaussians for this rank's particles (Model::EvaluateLogModel, Model.hpp:305-308, for the
// models of Model.hpp:55-92), through log-sum-exp like the gradient kernel.  An inspection entry point: one thread per particle.
__global__ void mvn_sum_logp_f64_kernel(const double *__restrict__ X, int d, int64_t row0, int64_t n_rows, int C, const double *__restrict__ means,
                                        const double *__restrict__ prec, double *__restrict__ logp)
{
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_rows) return;
    const double *x = X + (row0 + li) * d;
    double top = -INFINITY, tot = 0.0; // online log-sum-exp over the components
    for (int c = 0; c < C; ++c) {
        const double *P = prec + (size_t)c * d * d, *mu = means + (size_t)c * d;
        double q = 0.0;
        for (int r = 0; r < d; ++r) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += P[(size_t)r * d + k] * (x[k] - mu[k]);
            q += (x[r] - mu[r]) * s;
        }
        const double h = -0.5 * q;
        if (h > top) { tot = tot * exp(top - h) + 1.0; top = h; }
        else tot += exp(h - top);
    }
    logp[li] = top + log(tot);
}

} // namespace svgdb
