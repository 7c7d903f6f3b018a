// kernels_tc32.cuh — the tensor-core (tcgen05 / TMEM / TMA) path of the SVGD step for sm_100a.
//
// Same algebra as kernels_f64.cuh, evaluated in "FP32-class" arithmetic (SVGDB_PRECISION_TC32):
//   * particles are centred (x~ = x - mean; K and the repulsive term only see differences) and split
//     into two bf16 terms x~ = hi + lo; S = x~_i . x~_j is one bf16 tensor-core contraction over
//     K = 3*64: [hi|hi|lo]_i . [hi|lo|hi]_j = hi.hi + hi.lo + lo.hi (the dropped lo.lo term is 2^-18
//     relative), accumulated in fp32 in TMEM;
//   * E = exp2(2c S - c r_i - c r_j), c = a log2(e), with r = |x~|^2 from FP64, one MUFU.EX2 per pair;
//   * E is scaled by 2^15, rounded to fp16 (11-bit significand; every k >= 2^-29 stays a normal number)
//     and written back to TMEM as the A operand of the second contraction against
//     V^T = [v_hi | 1 | v_lo] (v = g - 2 a x~ split in two fp16 terms; the ones column yields the row
//     sum with the SAME rounded E, so the k(x_i,x_i) = 1 self term cancels exactly in the repulsion);
//   * the optimizer, clamp and the particle state stay FP64 (opt_update_tc32_kernel).
// Error bound and measurements: DESIGN.md "Precision modes".
//
// Reference semantics: SVGD.hpp:407-454, Kernel/GaussianRBFKernel.hpp:75-81,168-188 (see kernels_f64.cuh).
#pragma once
#include "kernels_f64.cuh"
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace svgdb {
namespace tc {

constexpr int TC_D = 64;        // padded particle dimension (d <= 64 in this path)
constexpr int TC_KCH = 3;       // 64-wide K chunks of the distance contraction: hi.hi, hi.lo, lo.hi
constexpr int TC_KTOT = TC_KCH * 64;
constexpr int TC_NVH = 80;      // V^T rows: [0,64) v_hi, 64 = ones, [65,80) zero
constexpr int TC_NV = 144;      //           [80,144) v_lo
constexpr int TC_ONES_ROW = 64;
constexpr int TC_PHI_LD = 80;   // phi_buf row: [0,64) sum_j E v, 64 = sum_j E
constexpr int TC_TILE = 128;
constexpr float TC_E_SCALE_LOG2 = 15.0f; // E is stored as fp16(2^15 E)
constexpr float TC_E_UNSCALE = 1.0f / 32768.0f;

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

// one warp per particle: centred bf16 split into the A and B operand layouts, r = |x~|^2
__global__ void split_kernel(const double *__restrict__ X, const double *__restrict__ colsum, int64_t n, int64_t n_pad, int d,
                             __nv_bfloat16 *__restrict__ XA, __nv_bfloat16 *__restrict__ XB, double *__restrict__ rt,
                             float *__restrict__ rf)
{
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int k = lane + 32 * h;
        double xc = 0.0;
        if (row < n && k < d) xc = X[row * d + k] - colsum[k] / (double)n;
        s += xc * xc;
        float xf = (float)xc;
        __nv_bfloat16 hi = __float2bfloat16_rn(xf);
        __nv_bfloat16 lo = __float2bfloat16_rn((float)(xc - (double)__bfloat162float(hi)));
        __nv_bfloat16 *a = XA + row * TC_KTOT, *b = XB + row * TC_KTOT;
        a[k] = hi; a[64 + k] = hi; a[128 + k] = lo;
        b[k] = hi; b[64 + k] = lo; b[128 + k] = hi;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { rt[row] = s; rf[row] = (float)s; }
}

// V^T (bf16, [TC_NV][ldn]) and beta = -a log2(e) r from the FP64 V = G - 2 a X (uncentred) of all particles:
// v~ = V + 2 a mean.  One block = 64 particles, transposed through shared memory.
__global__ void __launch_bounds__(256)
make_vt_kernel(const double *__restrict__ V, const double *__restrict__ colsum, const double *__restrict__ rt,
               const double *__restrict__ a_ptr, int64_t n, int64_t ldn, int d, __half *__restrict__ VT,
               float *__restrict__ beta)
{
    __shared__ __half tile[TC_NV][64 + 2];
    const double a = *a_ptr;
    const int64_t j0 = (int64_t)blockIdx.x * 64;
    for (int t = threadIdx.x; t < 64 * 64; t += blockDim.x) {
        int jl = t >> 6, c = t & 63;
        int64_t j = j0 + jl;
        float hi = 0.f, lo = 0.f;
        if (j < n && c < d) {
            double v = V[j * d + c] + 2.0 * a * (colsum[c] / (double)n);
            hi = __half2float(__float2half_rn((float)v));
            lo = (float)(v - (double)hi);
        }
        tile[c][jl] = __float2half_rn(hi);
        tile[TC_NVH + c][jl] = __float2half_rn(lo);
    }
    for (int t = threadIdx.x; t < 16 * 64; t += blockDim.x) {
        int rr = TC_ONES_ROW + (t >> 6), jl = t & 63;
        tile[rr][jl] = __float2half_rn((rr == TC_ONES_ROW && j0 + jl < n) ? 1.f : 0.f);
    }
    if (threadIdx.x < 64) {
        int64_t j = j0 + threadIdx.x;
        if (j < ldn) beta[j] = (j < n) ? (float)(-a * 1.4426950408889634 * rt[j]) : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < TC_NV * 64; t += blockDim.x) {
        int rr = t >> 6, jl = t & 63;
        if (j0 + jl < ldn) VT[(int64_t)rr * ldn + j0 + jl] = tile[rr][jl];
    }
}

struct OptTcArgs {
    const double *X;
    const double *colsum;
    const float *phi_buf;
    const double *a_ptr;
    int64_t n_total, row0, n_rows;
    int d;
    OptParams opt;
    double *s1, *s2;
    const double *lb, *ub;
    double *X_out, *phi_out;
};

// phi = (Phi + 2 a x~ rowsum)/n in FP64, then the optimizer increment and clamp (FP64 state).
__global__ void opt_update_tc32_kernel(OptTcArgs p)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n_rows * p.d) return;
    int64_t li = idx / p.d;
    int c = (int)(idx - li * p.d);
    int64_t i = p.row0 + li;
    const double a = *p.a_ptr;
    double x = p.X[i * p.d + c];
    double xc = x - p.colsum[c] / (double)p.n_total;
    double acc = (double)p.phi_buf[i * TC_PHI_LD + c];
    double rowsum = (double)p.phi_buf[i * TC_PHI_LD + TC_ONES_ROW];
    double phi = (1.0 / (double)p.n_total) * (acc + 2.0 * a * xc * rowsum);
    if (p.phi_out != nullptr) {
        p.phi_out[idx] = phi;
    } else {
        double xn = x + opt_increment(p.opt, phi, p.s1, p.s2, idx);
        p.X_out[i * p.d + c] = clamp_coord(xn, p.lb, p.ub, c);
    }
}

// ---- the fused pair-interaction kernel ---------------------------------------------------------------
struct PhiTcArgs {
    const float *beta;   // [n_pad] -c r_j (0 beyond n)
    const double *a_ptr;
    float *phi_buf;      // [n_pad][TC_PHI_LD], zeroed; partial sums are added atomically
    int64_t n_total, row0, n_rows;
    int n_jtiles, jsplit;
    int *err;
};

constexpr uint32_t TC_A_BYTES = TC_KCH * 16384;            // resident X_i tile
constexpr uint32_t TC_B_BYTES = TC_KCH * 16384;            // one X_j tile
constexpr uint32_t TC_V_BYTES = 2 * TC_NV * 128;           // one V^T tile (two 64-wide K chunks)
constexpr uint32_t TC_PHI_SMEM = TC_A_BYTES + 2 * TC_B_BYTES + 2 * TC_V_BYTES + 256 + 1024;

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// grid.x = n_itiles * jsplit; 320 threads: warps 0-3 / 4-7 = exp warpgroups for S buffers 0 / 1 (thread =
// TMEM lane = row i), warp 8 = TMA producer, warp 9 = MMA issuer.  TMEM: S0 [0,128) S1 [128,256) Phi [256,400);
// E_b (fp16 pairs) overwrites the first 64 columns of S_b in place.
__global__ void __launch_bounds__(320, 1)
phi_tc32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapV, PhiTcArgs p)
{
    const int it = blockIdx.x / p.jsplit, js = blockIdx.x - it * p.jsplit;
    const int tps = (p.n_jtiles + p.jsplit - 1) / p.jsplit;
    const int jbeg = js * tps;
    const int nt = min(p.n_jtiles, jbeg + tps) - jbeg;
    if (nt <= 0) return;
    const int64_t i0 = p.row0 + (int64_t)it * TC_TILE;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + TC_A_BYTES;
    uint8_t *sV = sB + 2 * TC_B_BYTES;
    uint64_t *bars = (uint64_t *)(sV + 2 * TC_V_BYTES);
    uint64_t *a_full = bars + 0, *b_full = bars + 1, *b_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 7;
    uint64_t *s_full = bars + 9, *e_ready = bars + 11, *phi_full = bars + 13;
    uint32_t *tmem_holder = (uint32_t *)(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1);
            mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1);
            mbar_init(s_full + s, 1); mbar_init(e_ready + s, 128);
        }
        mbar_init(phi_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t tS[2] = {tmem, tmem + 128};
    const uint32_t tPhi = tmem + 256;

    if (warp == 8) {
        if (lane == 0) { // ---- TMA producer ------------------------------------------------------------
            mbar_arrive_expect_tx(a_full, TC_A_BYTES);
            for (int c = 0; c < TC_KCH; ++c) tma_load_2d(sA + c * 16384, &mapA, c * 64, (int)i0, a_full);
            for (int t = 0; t < nt; ++t) {
                const int slot = t & 1, ph = (t >> 1) & 1;
                const int j0 = (jbeg + t) * TC_TILE;
                if (!mbar_wait(b_empty + slot, ph ^ 1, p.err, 10)) break;
                mbar_arrive_expect_tx(b_full + slot, TC_B_BYTES);
                for (int c = 0; c < TC_KCH; ++c) tma_load_2d(sB + slot * TC_B_BYTES + c * 16384, &mapB, c * 64, j0, b_full + slot);
                if (!mbar_wait(v_empty + slot, ph ^ 1, p.err, 11)) break;
                mbar_arrive_expect_tx(v_full + slot, TC_V_BYTES);
                for (int c = 0; c < 2; ++c) tma_load_2d(sV + slot * TC_V_BYTES + c * TC_NV * 128, &mapV, j0 + c * 64, 0, v_full + slot);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) { // ---- MMA issuer ----------------------------------------------------------------
            const uint32_t idesc_s = make_idesc_bf16(TC_TILE, TC_TILE);
            const uint32_t idesc_v = make_idesc_f16(TC_TILE, TC_NV);
            bool ok = mbar_wait(a_full, 0, p.err, 20);
            auto issue_s = [&](int t) -> bool {
                const int slot = t & 1, ph = (t >> 1) & 1;
                if (!mbar_wait(b_full + slot, ph, p.err, 21)) return false;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < TC_KCH; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint64_t da = make_desc_k_sw128(smem_u32(sA + c * 16384) + k * 32);
                        uint64_t db = make_desc_k_sw128(smem_u32(sB + slot * TC_B_BYTES + c * 16384) + k * 32);
                        umma_bf16_ss(tS[slot], da, db, idesc_s, (c | k) ? 1u : 0u);
                    }
                umma_commit(b_empty + slot);
                umma_commit(s_full + slot);
                return true;
            };
            if (ok) ok = issue_s(0);
            for (int t = 0; ok && t < nt; ++t) {
                if (t + 1 < nt && !issue_s(t + 1)) { ok = false; break; }
                const int slot = t & 1, ph = (t >> 1) & 1;
                if (!mbar_wait(e_ready + slot, ph, p.err, 22)) { ok = false; break; }
                if (!mbar_wait(v_full + slot, ph, p.err, 23)) { ok = false; break; }
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint64_t db = make_desc_k_sw128(smem_u32(sV + slot * TC_V_BYTES + c * TC_NV * 128) + k * 32);
                        umma_bf16_ts(tPhi, tS[slot] + (c * 4 + k) * 8, db, idesc_v, (t | c | k) ? 1u : 0u);
                    }
                umma_commit(v_empty + slot);
            }
            if (ok) umma_commit(phi_full);
        }
    } else { // ---- exp warpgroups ----------------------------------------------------------------------------
        const int b = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int64_t i = i0 + row;
        const float alpha = ((i < p.n_total) ? p.beta[i] : 0.f) + TC_E_SCALE_LOG2;
        const float two_c = (float)(2.0 * (*p.a_ptr) * 1.4426950408889634);
        bool ok = true;
        for (int t = b; ok && t < nt; t += 2) {
            const int ph = (t >> 1) & 1;
            if (!mbar_wait(s_full + b, ph, p.err, 30 + b)) { ok = false; break; }
            tc_fence_after();
            const int64_t j0 = (int64_t)(jbeg + t) * TC_TILE;
            const int64_t dcol = i - j0; // column of k(x_i, x_i) in this tile, if inside [0,128)
            const bool tile_has_diag = (j0 < i0 + TC_TILE) && (j0 + TC_TILE > i0);
#pragma unroll 1
            for (int c0 = 0; c0 < TC_TILE; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tS[b] + lane_base + c0, r);
                float4 bq[8];
                const float4 *bp = reinterpret_cast<const float4 *>(p.beta + j0 + c0);
#pragma unroll
                for (int q = 0; q < 8; ++q) bq[q] = __ldg(bp + q);
                tmem_ld_wait();
                float e[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    e[4 * q + 0] = ex2_approx(fmaf(__uint_as_float(r[4 * q + 0]), two_c, alpha + bq[q].x));
                    e[4 * q + 1] = ex2_approx(fmaf(__uint_as_float(r[4 * q + 1]), two_c, alpha + bq[q].y));
                    e[4 * q + 2] = ex2_approx(fmaf(__uint_as_float(r[4 * q + 2]), two_c, alpha + bq[q].z));
                    e[4 * q + 3] = ex2_approx(fmaf(__uint_as_float(r[4 * q + 3]), two_c, alpha + bq[q].w));
                }
                if (tile_has_diag) {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (dcol == c0 + q) e[q] = 32768.0f; // k(x_i, x_i) = exp(0) exactly, like the reference
                }
                uint32_t packed[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) packed[q] = pack_f16x2(e[2 * q], e[2 * q + 1]);
                tmem_st16(tS[b] + lane_base + c0 / 2, packed);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(e_ready + b);
        }
        if (b == 0 && ok) { // ---- flush Phi: TMEM -> global partial sums ------------------------------------
            if (mbar_wait(phi_full, 0, p.err, 40)) {
                tc_fence_after();
                const bool valid = i < p.row0 + p.n_rows;
                float *dst = p.phi_buf + i * TC_PHI_LD;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t hi[16], lo[16];
                    tmem_ld16(tPhi + lane_base + c0, hi);
                    tmem_ld16(tPhi + lane_base + TC_NVH + c0, lo);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) atomicAdd(dst + c0 + q, (__uint_as_float(hi[q]) + __uint_as_float(lo[q])) * TC_E_UNSCALE);
                    }
                }
                uint32_t rs[16];
                tmem_ld16(tPhi + lane_base + TC_ONES_ROW, rs);
                tmem_ld_wait();
                if (valid) atomicAdd(dst + TC_ONES_ROW, __uint_as_float(rs[0]) * TC_E_UNSCALE);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

} // namespace tc
} // namespace svgdb
