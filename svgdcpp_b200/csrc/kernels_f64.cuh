// kernels_f64.cuh — FP64 (DMMA tensor-core) kernels of the SVGD step for sm_100a.
//
// Reference semantics being reproduced (paths relative to /root/reference/include/SVGDCpp/):
//   SVGD.hpp:407-454            phi_i = (1/n) sum_j [ k(x_j,x_i) g_j + grad_{x_j} k(x_j,x_i) ]
//   Kernel/GaussianRBFKernel.hpp:75-81   k = exp(-(x-x')^T A (x-x')), A = a I
//   Kernel/GaussianRBFKernel.hpp:168-188 a = log(n) / median(|x_i - x_j|)^2 over all n^2 ordered pairs
//   Model/MultivariateNormal.hpp:56-61, Model/Model.hpp:55-92   unweighted sum of unnormalised Gaussians
//   Optimizer/Adam.hpp:75-96, AdaGrad.hpp:60-65, RMSProp.hpp:69-74;  clamp SVGD.hpp:396-399
//
// Algebra (DESIGN.md "Rewrite"): with v_j = g_j - 2 a x_j, r_i = |x_i|^2 and
//   K_ji = exp(-a (r_i + r_j - 2 x_i.x_j)),   phi_i = (1/n) [ sum_j K_ji v_j + 2 a x_i sum_j K_ji ]
// so one pass needs two contractions (X X^T and K V) and never stores K or grad K.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svgdb {

enum { OPT_ADAGRAD = 0, OPT_ADAM = 1, OPT_RMSPROP = 2 };

struct OptParams {
    int kind;
    double lr, beta1, beta2, eps;
    double bias1, bias2; // 1 - beta^t for Adam (host pow(), Adam.hpp:93-96)
};

// D[8x8] += A[8x4] * B[4x8]; lane = 4*g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One optimizer update on one coordinate; returns the increment the driver adds (SVGD.hpp:393).
__device__ __forceinline__ double opt_increment(const OptParams &o, double phi, double *s1, double *s2, int64_t idx)
{
    if (o.kind == OPT_ADAM) {
        double m = o.beta1 * s2[idx] + (1 - o.beta1) * phi;
        double v = o.beta2 * s1[idx] + (1 - o.beta2) * (phi * phi);
        s2[idx] = m;
        s1[idx] = v;
        return o.lr * (1.0 / (o.eps + sqrt(v / o.bias2))) * (m / o.bias1);
    } else if (o.kind == OPT_ADAGRAD) {
        double s = s1[idx] + phi * phi;
        s1[idx] = s;
        return o.lr * (1.0 / (o.eps + sqrt(s))) * phi;
    } else {
        double s = o.beta1 * s1[idx] + (1 - o.beta1) * (phi * phi);
        s1[idx] = s;
        return o.lr * (1.0 / (o.eps + sqrt(s))) * phi;
    }
}

__device__ __forceinline__ double clamp_coord(double x, const double *lb, const double *ub, int k)
{
    if (lb != nullptr) {
        x = fmin(x, ub[k]); // min with the upper bound first, then max with the lower (SVGD.hpp:398)
        x = fmax(x, lb[k]);
    }
    return x;
}

// ---------------------------------------------------------------------------------------------
// r_i = |x_i|^2 (one warp per particle).
// ---------------------------------------------------------------------------------------------
__global__ void rownorm_f64_kernel(const double *__restrict__ X, int64_t n, int d, double *__restrict__ r)
{
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int k = lane; k < d; k += 32) { double v = X[row * d + k]; s += v * v; }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) r[row] = s;
}

// V = G - 2 a X for rows [row0, row0 + n_rows); G is indexed by local row, V by global row.
__global__ void make_v_f64_kernel(const double *__restrict__ X, const double *__restrict__ G,
                                  const double *__restrict__ a_ptr, int64_t row0, int64_t n_rows, int d,
                                  double *__restrict__ V)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * d) return;
    double a = *a_ptr;
    V[row0 * d + idx] = G[idx] - 2.0 * a * X[row0 * d + idx];
}

// ---------------------------------------------------------------------------------------------
// grad log p for a sum of C unnormalised Gaussians, evaluated through an online log-sum-exp over
// the components (finite wherever at least one component is representable in log space; the
// reference's literal log(sum exp) turns NaN once every exp underflows).
// One block = PT particles; thread r-loop over output coordinates; P_c symmetric so P[k][r] is
// read coalesced along r.  smem: diffT[d][PT], Y[PT][d], Gacc[PT][d] doubles.
// ---------------------------------------------------------------------------------------------
template <int PT>
__global__ void __launch_bounds__(128)
mvn_sum_grad_f64_kernel(const double *__restrict__ X, int64_t n_total, int d, int64_t row0, int64_t n_rows,
                        int C, const double *__restrict__ means, const double *__restrict__ prec,
                        double *__restrict__ G)
{
    extern __shared__ double sm[];
    double *diffT = sm;                      // [d][PT]
    double *Y = diffT + (size_t)d * PT;      // [PT][d]
    double *Gacc = Y + (size_t)PT * d;       // [PT][d]
    double *q = Gacc + (size_t)PT * d;       // [PT]
    double *mrun = q + PT, *srun = mrun + PT, *wgt = srun + PT, *scl = wgt + PT;
    const int tid = threadIdx.x;
    const int64_t p0 = row0 + (int64_t)blockIdx.x * PT; // first global particle of this block
    (void)n_total;

    for (int t = tid; t < PT * d; t += blockDim.x) Gacc[t] = 0.0;
    if (tid < PT) { mrun[tid] = -INFINITY; srun[tid] = 0.0; }

    for (int c = 0; c < C; ++c) {
        const double *mu = means + (size_t)c * d;
        const double *P = prec + (size_t)c * d * d;
        __syncthreads();
        for (int t = tid; t < PT * d; t += blockDim.x) {
            int p = t / d, k = t - p * d;
            int64_t row = p0 + p;
            diffT[(size_t)k * PT + p] = (row < row0 + n_rows) ? X[row * d + k] - mu[k] : 0.0;
        }
        if (tid < PT) q[tid] = 0.0;
        __syncthreads();
        double qpart[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) qpart[p] = 0.0;
        for (int r = tid; r < d; r += blockDim.x) {
            double acc[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) acc[p] = 0.0;
            for (int k = 0; k < d; ++k) {
                double pk = P[(size_t)k * d + r];
                const double2 *dv = reinterpret_cast<const double2 *>(diffT + (size_t)k * PT);
#pragma unroll
                for (int p2 = 0; p2 < PT / 2; ++p2) {
                    double2 v = dv[p2];
                    acc[2 * p2] = fma(pk, v.x, acc[2 * p2]);
                    acc[2 * p2 + 1] = fma(pk, v.y, acc[2 * p2 + 1]);
                }
            }
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                Y[(size_t)p * d + r] = acc[p];
                qpart[p] = fma(diffT[(size_t)r * PT + p], acc[p], qpart[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            double v = qpart[p];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) atomicAdd(&q[p], v);
        }
        __syncthreads();
        if (tid < PT) {
            double h = -0.5 * q[tid];
            double m_new = fmax(mrun[tid], h);
            double sc = (mrun[tid] == -INFINITY) ? 0.0 : exp(mrun[tid] - m_new);
            double w = exp(h - m_new);
            srun[tid] = srun[tid] * sc + w;
            mrun[tid] = m_new;
            scl[tid] = sc;
            wgt[tid] = w;
        }
        __syncthreads();
        for (int t = tid; t < PT * d; t += blockDim.x) {
            int p = t / d;
            Gacc[t] = Gacc[t] * scl[p] - wgt[p] * Y[t];
        }
    }
    __syncthreads();
    for (int t = tid; t < PT * d; t += blockDim.x) {
        int p = t / d, k = t - p * d;
        int64_t row = p0 + p;
        if (row < row0 + n_rows) G[(row - row0) * d + k] = Gacc[t] / srun[p];
    }
}

// ---------------------------------------------------------------------------------------------
// The mixture gradient as library GEMMs (launch_grad_gemm in svgd_b200_api.cu): the two streaming kernels around each
// Y = (X - mu_c) Sigma_c^-1.  Same online log-sum-exp as mvn_sum_grad_f64_kernel, state kept per particle in global memory.
// ---------------------------------------------------------------------------------------------
// D[t] = X[t] - mu[t mod d] over the rank's rows (X already offset to its first row)
__global__ void __launch_bounds__(256)
grad_diff_kernel(const double *__restrict__ X, int64_t cnt, int d, const double *__restrict__ mu, double *__restrict__ D)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += (int64_t)gridDim.x * blockDim.x) D[t] = X[t] - mu[t % d];
}

// one warp per particle: q = D_i . Y_i, h = -q / 2, (m, s) <- online log-sum-exp, G_i <- G_i e^(m - m') - e^(h - m') Y_i; the last
// component divides by s.  c = 0 initialises (G, m, s are not read).
__global__ void __launch_bounds__(256)
grad_accumulate_kernel(const double *__restrict__ D, const double *__restrict__ Y, int64_t rows, int d, int c, int last,
                       double *__restrict__ mrun, double *__restrict__ srun, double *__restrict__ G)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= rows) return;
    const double *Di = D + i * d, *Yi = Y + i * d;
    double *Gi = G + i * d;
    double q = 0.0;
    for (int r = lane; r < d; r += 32) q = fma(Di[r], Yi[r], q);
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const double h = -0.5 * q;
    const double m_old = c == 0 ? -INFINITY : mrun[i], s_old = c == 0 ? 0.0 : srun[i];
    const double m_new = fmax(m_old, h);
    const double sc = (m_old == -INFINITY) ? 0.0 : exp(m_old - m_new);
    const double w = exp(h - m_new);
    const double s_new = s_old * sc + w;
    for (int r = lane; r < d; r += 32) {
        const double g = (c == 0 ? 0.0 : Gi[r] * sc) - w * Yi[r];
        Gi[r] = last ? g / s_new : g;
    }
    if (lane == 0 && !last) { mrun[i] = m_new; srun[i] = s_new; }
}

// ---------------------------------------------------------------------------------------------
// Shared tile loader: rows [row_base, row_base+64) x cols [col_base, col_base+ncols) of a
// particle-contiguous matrix into smem with leading dimension ld, zero-filled outside (n, d)
// and up to ncols_pad columns.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tile64(double *dst, int ld, const double *__restrict__ src, int64_t n, int d,
                                            int64_t row_base, int col_base, int ncols_pad, int tid, int nthreads)
{
    for (int t = tid; t < 64 * ncols_pad; t += nthreads) {
        int rr = t / ncols_pad, cc = t - rr * ncols_pad;
        int64_t row = row_base + rr;
        int col = col_base + cc;
        dst[rr * ld + cc] = (row < n && col < d) ? src[row * d + col] : 0.0;
    }
}

struct PhiArgs {
    const double *X;      // [N][d] all particles (this step's positions)
    const double *V;      // [N][d] g_j - 2 a x_j
    const double *r;      // [N] |x_j|^2
    const double *a_ptr;  // device scalar: kernel scale a
    int64_t n_total;
    int d;
    int64_t row0, n_rows; // this GPU's particle rows
    OptParams opt;
    double *s1, *s2;      // optimizer state, local rows [n_rows][d]
    const double *lb, *ub; // bounds per coordinate or nullptr
    double *X_out;        // [N][d] next positions (rows [row0,row0+n_rows) written) or nullptr
    double *phi_out;      // [n_rows][d] if non-null: write phi, skip the update
};

// ---------------------------------------------------------------------------------------------
// Fused pair-interaction + optimizer kernel (FP64, DMMA m8n8k4).
// CTA = 4 warps = 64 rows i (16 per warp); loops over all j in tiles of 64.  grid.y = output column
// groups of DC (<= 64) coordinates; S is contracted over the full d in k-chunks of <= 64.
//   MMA-1  S = X_i X_j^T                 (A = X_i rows, B = X_j rows)
//   E = exp(-a max(r_i + r_j - 2 S, 0)), E_ii = 1, masked for j >= N; row sums by quad shuffles
//   MMA-2  Phi += E V_j                   (the C fragments of S are re-used in place as A fragments:
//          accumulator slot t of an 8-wide j block stands for j = 2t (+1), V rows are read to match)
// Epilogue: phi = (Phi + 2 a x_i rowsum)/n -> Adam/AdaGrad/RMSProp increment -> clamp -> X_out.
// ---------------------------------------------------------------------------------------------
template <int DC>
__global__ void __launch_bounds__(128, 2) phi_f64_kernel(PhiArgs p)
{
    constexpr int NCB = DC / 8;
    extern __shared__ double sm[];
    const int d = p.d;
    const int kc_max = d < 64 ? ((d + 3) & ~3) : 64;     // k extent staged per chunk
    const int ldx = ((kc_max + 15) & ~15) + 4;           // == 4 mod 16: conflict-free A/B fragment reads
    constexpr int ldv = DC + 2;                          // == 2 mod 8: conflict-free V fragment reads
    double *Xi = sm;                 // [64][ldx]
    double *Xj = Xi + 64 * ldx;      // [64][ldx]
    double *Vj = Xj + 64 * ldx;      // [64][ldv]
    double *rj = Vj + 64 * ldv;      // [64]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int64_t i0 = p.row0 + (int64_t)blockIdx.x * 64;
    const int64_t row_end = p.row0 + p.n_rows;
    const int col0 = blockIdx.y * DC;
    const bool single_chunk = d <= 64;
    const double a = *p.a_ptr;
    const int64_t N = p.n_total;

    double phi_acc[2][NCB][2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) phi_acc[rb][cb][0] = phi_acc[rb][cb][1] = 0.0;
    double rowsum[2] = {0.0, 0.0};

    const int64_t irow[2] = {i0 + warp * 16 + g, i0 + warp * 16 + 8 + g};
    double ri[2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) ri[rb] = irow[rb] < N ? p.r[irow[rb]] : 0.0;

    if (single_chunk) load_tile64(Xi, ldx, p.X, N, d, i0, 0, kc_max, tid, 128);

    for (int64_t j0 = 0; j0 < N; j0 += 64) {
        double S[2][8][2];
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int jb = 0; jb < 8; ++jb) S[rb][jb][0] = S[rb][jb][1] = 0.0;

        for (int kc = 0; kc < d; kc += 64) {
            __syncthreads(); // everyone is done with the previous contents of Xj / Vj / Xi
            if (!single_chunk) load_tile64(Xi, ldx, p.X, N, d, i0, kc, kc_max, tid, 128);
            load_tile64(Xj, ldx, p.X, N, d, j0, kc, kc_max, tid, 128);
            if (kc == 0) {
                load_tile64(Vj, ldv, p.V, N, d, j0, col0, DC, tid, 128);
                if (tid < 64) rj[tid] = (j0 + tid < N) ? p.r[j0 + tid] : 0.0;
            }
            __syncthreads();
            const double *xa0 = Xi + (warp * 16 + g) * ldx + t;
            const double *xa1 = xa0 + 8 * ldx;
            const double *xb = Xj + g * ldx + t;
#pragma unroll 4
            for (int k = 0; k < kc_max; k += 4) {
                double a0 = xa0[k], a1 = xa1[k];
#pragma unroll
                for (int jb = 0; jb < 8; ++jb) {
                    double b = xb[jb * 8 * ldx + k];
                    dmma884(S[0][jb][0], S[0][jb][1], a0, b);
                    dmma884(S[1][jb][0], S[1][jb][1], a1, b);
                }
            }
        }

        // E = exp(-a D2), in place
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int jb = 0; jb < 8; ++jb)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int jl = jb * 8 + 2 * t + h;
                    int64_t j = j0 + jl;
                    double d2 = fmax(ri[rb] + rj[jl] - 2.0 * S[rb][jb][h], 0.0);
                    double e = exp(-a * d2);
                    if (j == irow[rb]) e = 1.0;   // k(x_i, x_i) = exp(0) exactly, like the reference
                    if (j >= N) e = 0.0;
                    S[rb][jb][h] = e;
                    rowsum[rb] += e;
                }

        // Phi += E V_j
#pragma unroll
        for (int jb = 0; jb < 8; ++jb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double *vb = Vj + (jb * 8 + 2 * t + h) * ldv + g;
                double e0 = S[0][jb][h], e1 = S[1][jb][h];
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb) {
                    double b = vb[cb * 8];
                    dmma884(phi_acc[0][cb][0], phi_acc[0][cb][1], e0, b);
                    dmma884(phi_acc[1][cb][0], phi_acc[1][cb][1], e1, b);
                }
            }
    }

    // row sums: the 4 lanes of a quad hold disjoint column subsets of the same row
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        rowsum[rb] += __shfl_xor_sync(0xffffffffu, rowsum[rb], 1);
        rowsum[rb] += __shfl_xor_sync(0xffffffffu, rowsum[rb], 2);
    }

    const double inv_n = 1.0 / (double)N;
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        int64_t i = irow[rb];
        if (i >= row_end) continue;
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int c = col0 + cb * 8 + 2 * t + h;
                if (c >= d) continue;
                double x = p.X[i * d + c];
                double phi = inv_n * (phi_acc[rb][cb][h] + 2.0 * a * x * rowsum[rb]);
                int64_t li = (i - p.row0) * d + c;
                if (p.phi_out != nullptr) {
                    p.phi_out[li] = phi;
                } else {
                    double xn = x + opt_increment(p.opt, phi, p.s1, p.s2, li);
                    p.X_out[i * d + c] = clamp_coord(xn, p.lb, p.ub, c);
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Pairwise squared-distance pass for the exact median (FP64, DMMA).  Persistent CTAs stride over
// 64x64 tiles.  sym != 0: only tiles tj >= ti are visited and off-diagonal tiles weigh 2 (D2 is
// bitwise symmetric: same products, same k order); sym == 0: rectangular rows [row0,row0+n_rows)
// x all columns, weight 1 (row-sharded multi-GPU).
// Keys: the IEEE bit pattern of D2 >= 0 is order preserving.  For every element
//     key <  lo      -> below += w, max_below = max(max_below, key)
//     lo <= key < hi -> MODE_HIST: hist[(key - lo) >> shift] += w
//                       MODE_COLLECT: append key w times to cand[] (count kept even past capacity)
// ---------------------------------------------------------------------------------------------
enum { MODE_HIST = 0, MODE_COLLECT = 1 };
constexpr int HIST_BINS = 4096;

struct DistArgs {
    const double *X;
    const double *r;
    int64_t n_total;
    int d;
    int64_t row0, n_rows;
    int sym;
    int64_t n_tiles_i, n_tiles_j, n_work; // tile counts and number of tile pairs
    int work_offset, work_stride;         // this rank takes tile pairs offset, offset + stride, ... (cyclic over ranks)
    uint64_t lo, hi;
    int shift;
    unsigned long long *below;      // [1]
    unsigned long long *max_below;  // [1]
    unsigned long long *hist;       // [HIST_BINS]           (MODE_HIST)
    unsigned long long *cand;       // [capacity]            (MODE_COLLECT)
    unsigned long long *cand_count; // [1]
    uint64_t capacity;
};

__device__ __forceinline__ void decode_tile(const DistArgs &p, int64_t w, int64_t &ti, int64_t &tj)
{
    if (!p.sym) {
        ti = w / p.n_tiles_j;
        tj = w - ti * p.n_tiles_j;
        return;
    }
    // row-major upper triangle of a T x T grid: row ti holds T - ti tiles
    const double T = (double)p.n_tiles_j;
    double f = ((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)w)) * 0.5;
    ti = (int64_t)f;
    if (ti < 0) ti = 0;
    if (ti >= p.n_tiles_j) ti = p.n_tiles_j - 1;
    auto start = [&](int64_t row) { return row * p.n_tiles_j - row * (row - 1) / 2; };
    while (ti > 0 && start(ti) > w) --ti;
    while (start(ti + 1) <= w) ++ti;
    tj = ti + (w - start(ti));
}

template <int MODE>
__global__ void __launch_bounds__(128, 2) dist_pass_f64_kernel(DistArgs p)
{
    extern __shared__ double sm[];
    const int d = p.d;
    const int kc_max = d < 64 ? ((d + 3) & ~3) : 64;
    const int ldx = ((kc_max + 15) & ~15) + 4;
    double *Xi = sm;
    double *Xj = Xi + 64 * ldx;
    double *ris = Xj + 64 * ldx; // [64]
    double *rjs = ris + 64;      // [64]
    unsigned int *shist = reinterpret_cast<unsigned int *>(rjs + 64); // [HIST_BINS] (MODE_HIST)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int64_t N = p.n_total;

    if (MODE == MODE_HIST) {
        for (int b = tid; b < HIST_BINS; b += 128) shist[b] = 0u;
    }
    unsigned long long below = 0ull, maxb = 0ull;

    for (int64_t w = (int64_t)blockIdx.x * p.work_stride + p.work_offset; w < p.n_work; w += (int64_t)gridDim.x * p.work_stride) {
        int64_t ti, tj;
        decode_tile(p, w, ti, tj);
        const int64_t i0 = p.row0 + ti * 64, j0 = tj * 64;
        const unsigned int wgt = (p.sym && ti != tj) ? 2u : 1u;
        const int64_t i_end = p.row0 + p.n_rows;

        double S[2][8][2];
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int jb = 0; jb < 8; ++jb) S[rb][jb][0] = S[rb][jb][1] = 0.0;

        for (int kc = 0; kc < d; kc += 64) {
            __syncthreads();
            load_tile64(Xi, ldx, p.X, N, d, i0, kc, kc_max, tid, 128);
            load_tile64(Xj, ldx, p.X, N, d, j0, kc, kc_max, tid, 128);
            if (kc == 0 && tid < 64) {
                ris[tid] = (i0 + tid < N) ? p.r[i0 + tid] : 0.0;
                rjs[tid] = (j0 + tid < N) ? p.r[j0 + tid] : 0.0;
            }
            __syncthreads();
            const double *xa0 = Xi + (warp * 16 + g) * ldx + t;
            const double *xa1 = xa0 + 8 * ldx;
            const double *xb = Xj + g * ldx + t;
#pragma unroll 4
            for (int k = 0; k < kc_max; k += 4) {
                double a0 = xa0[k], a1 = xa1[k];
#pragma unroll
                for (int jb = 0; jb < 8; ++jb) {
                    double b = xb[jb * 8 * ldx + k];
                    dmma884(S[0][jb][0], S[0][jb][1], a0, b);
                    dmma884(S[1][jb][0], S[1][jb][1], a1, b);
                }
            }
        }

#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
            const int il = warp * 16 + rb * 8 + g;
            const int64_t i = i0 + il;
#pragma unroll
            for (int jb = 0; jb < 8; ++jb)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jl = jb * 8 + 2 * t + h;
                    const int64_t j = j0 + jl;
                    const bool valid = (i < i_end) && (j < N);
                    double d2 = fmax(ris[il] + rjs[jl] - 2.0 * S[rb][jb][h], 0.0);
                    if (i == j) d2 = 0.0; // the Gram diagonal is the row norm: exactly 0 in the reference
                    const unsigned long long key = (unsigned long long)__double_as_longlong(d2);
                    const bool is_below = valid && key < p.lo;
                    const bool in_range = valid && key >= p.lo && key < p.hi;
                    if (is_below) { below += wgt; maxb = key > maxb ? key : maxb; }
                    if (MODE == MODE_HIST) {
                        if (in_range) atomicAdd(&shist[(unsigned int)((key - p.lo) >> p.shift)], wgt);
                    } else {
                        const unsigned int mask = __ballot_sync(0xffffffffu, in_range);
                        if (mask) {
                            const int leader = __ffs(mask) - 1;
                            unsigned long long base = 0ull;
                            if (lane == leader)
                                base = atomicAdd(p.cand_count, (unsigned long long)__popc(mask) * wgt);
                            base = __shfl_sync(0xffffffffu, base, leader);
                            if (in_range) {
                                unsigned long long slot = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u)) * wgt;
                                if (slot < p.capacity) p.cand[slot] = key;
                                if (wgt == 2u && slot + 1 < p.capacity) p.cand[slot + 1] = key;
                            }
                        }
                    }
                }
        }
    }

    // block-level reduction of the below counters
    for (int o = 16; o; o >>= 1) {
        below += __shfl_xor_sync(0xffffffffu, below, o);
        unsigned long long other = __shfl_xor_sync(0xffffffffu, maxb, o);
        maxb = other > maxb ? other : maxb;
    }
    if (lane == 0) {
        if (below) atomicAdd(p.below, below);
        if (maxb) atomicMax(p.max_below, maxb);
    }
    if (MODE == MODE_HIST) {
        __syncthreads();
        for (int b = tid; b < HIST_BINS; b += 128) {
            unsigned int c = shist[b];
            if (c) atomicAdd(&p.hist[b], (unsigned long long)c);
        }
    }
}

// Register-resident DMMA issue loop: 16 independent accumulator pairs per warp (roofline probe).
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double *sink)
{
    double c[16][2];
#pragma unroll
    for (int q = 0; q < 16; ++q) c[q][0] = c[q][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) dmma884(c[q][0], c[q][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += c[q][0] + c[q][1];
    if (s == 123.456) *sink = s;
}

} // namespace svgdb
