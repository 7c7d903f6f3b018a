// kernels_tc32.cuh — pieces shared by the tensor-core (tcgen05 / TMEM / TMA) path of SVGDB_PRECISION_TC32:
// constants, the column-sum kernel used to centre the particles, the FP64 optimizer kernel that follows the
// pair-interaction kernel, and small device helpers.  The two big kernels live in kernels_phi_tc.cuh
// (pair interaction) and kernels_dist_tc.cuh (distance pass of the median bandwidth); arithmetic and error
// bounds are described there and in DESIGN.md "Precision modes".
//
// Reference semantics: SVGD.hpp:407-454, Kernel/GaussianRBFKernel.hpp:75-81,168-188 (see kernels_f64.cuh).
#pragma once
#include "kernels_f64.cuh"
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace svgdb {
namespace tc {

constexpr int TC_D = 64;        // padded particle dimension (d <= 64 in this path)
constexpr int TC_ONES_ROW = 64; // phi_buf column that holds sum_j E
constexpr int TC_PHI_LD = 80;   // phi_buf row: [0,64) sum_j E v, 64 = sum_j E
constexpr int TC_TILE = 128;
constexpr float TC_E_UNSCALE = 1.0f / 32768.0f; // E is stored as fp16(2^15 k)

// ---- operand preparation ---------------------------------------------------------------------------
__global__ void colsum_kernel(const double *__restrict__ X, int64_t n, int d, double *__restrict__ sum)
{
    // grid-stride over rows; thread k-lane accumulates column k (d <= 64, blockDim = 64 x 4)
    int k = threadIdx.x & 63, sub = threadIdx.x >> 6;
    double acc = 0.0;
    if (k < d)
        for (int64_t row = (int64_t)blockIdx.x * 4 + sub; row < n; row += (int64_t)gridDim.x * 4) acc += X[row * d + k];
    if (k < d) atomicAdd(&sum[k], acc);
}

// The operands are centred on colsum / n.  Any common shift gives the same distances and the same phi, so when only the first
// rows of a new particle set have arrived (svgdb_step_host) their mean stands in for the mean of all of them: sum *= factor.
__global__ void colsum_scale_kernel(double *__restrict__ sum, int d, double factor)
{
    if ((int)threadIdx.x < d) sum[threadIdx.x] *= factor;
}

// Three-term bf16 split of a scalar (24 significant bits): v ~= t0 + t1 + t2.
__device__ __forceinline__ void split3_bf16(double v, __nv_bfloat16 &t0, __nv_bfloat16 &t1, __nv_bfloat16 &t2)
{
    t0 = __float2bfloat16_rn((float)v);
    double rem = isfinite(__bfloat162float(t0)) ? v - (double)__bfloat162float(t0) : 0.0;
    t1 = __float2bfloat16_rn((float)rem);
    rem -= (double)__bfloat162float(t1);
    t2 = __float2bfloat16_rn((float)rem);
}

struct OptTcArgs {
    const double *X;
    const double *colsum;
    const float *phi_buf;
    const double *a_ptr;
    int64_t n_total, row0, n_rows; // rows [row0, row0 + n_rows) of the particle matrix
    int64_t state_row0;            // first row of this rank: optimizer state and phi_out are indexed from it
    int d;
    OptParams opt;
    double *s1, *s2;
    const double *lb, *ub;
    double *X_out, *phi_out;
    const int *miss; // device flag of an optimistic step whose predicted median bracket missed: nothing persistent may change
};

// phi = (Phi + 2 a x~ rowsum)/n in FP64, then the optimizer increment and clamp (FP64 state).
__global__ void opt_update_tc32_kernel(OptTcArgs p)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n_rows * p.d) return;
    if (p.miss != nullptr && *p.miss != 0) return;
    int64_t li = idx / p.d;
    int c = (int)(idx - li * p.d);
    int64_t i = p.row0 + li;
    const int64_t sidx = (i - p.state_row0) * p.d + c;
    const double a = *p.a_ptr;
    double x = p.X[i * p.d + c];
    double xc = x - p.colsum[c] / (double)p.n_total;
    double acc = (double)p.phi_buf[i * TC_PHI_LD + c];
    double rowsum = (double)p.phi_buf[i * TC_PHI_LD + TC_ONES_ROW];
    double phi = (1.0 / (double)p.n_total) * (acc + 2.0 * a * xc * rowsum);
    if (p.phi_out != nullptr) {
        p.phi_out[sidx] = phi;
    } else {
        double xn = x + opt_increment(p.opt, phi, p.s1, p.s2, sidx);
        p.X_out[i * p.d + c] = clamp_coord(xn, p.lb, p.ub, c);
    }
}

// Development aid: timeline of CTA 0 of the pair-interaction kernel (SVGDB_TC_TRACE=<file>).
// trace slots: role 0 = MMA issuer of i-tile 0, 1 / 2 = one exp warp of i-tile 0 / 1; 8 events per j-tile, 64 j-tiles
// Compiled in only with -DSVGDB_TC_TRACE_BUILD (python -m svgdcpp_b200.build --trace): the predicates cost the exp warps ~15 % of their
// instructions per unit even when no trace buffer is given.
#ifdef SVGDB_TC_TRACE_BUILD
#define TC_TRACE(role, t, ev)                                                                      \
    do {                                                                                           \
        if (p.trace != nullptr && blockIdx.x == 0 && (t) < 64) p.trace[((role) * 64 + (t)) * 8 + (ev)] = clock64(); \
    } while (0)
#else
#define TC_TRACE(role, t, ev) do { } while (0)
#endif

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- helpers of the distance pass (kernels_dist_tc.cuh) -------------------------------------------------
constexpr int TC_WBUF = 512;  // candidate distances (fp32) staged per warp before one global reservation

// keys are the IEEE bits of (double)D2 (order preserving), shared with the FP64 path's select
__device__ __forceinline__ unsigned long long dist_key(float d2)
{
    return (unsigned long long)__double_as_longlong((double)fmaxf(d2, 0.0f));
}

} // namespace tc
} // namespace svgdb
