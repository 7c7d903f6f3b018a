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
//     [v_hi | 1] and [v_lo | 0] (v = g - 2 a x~ split in two fp16 terms, both accumulated into the same
//     TMEM columns; the ones column yields the row sum with the SAME rounded E, so the k(x_i,x_i) = 1
//     self term cancels exactly in the repulsion);
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
constexpr int TC_NV = 160;      //           [80,144) v_lo, [144,160) zero  (hi and lo blocks are contracted into
constexpr int TC_ONES_ROW = 64; //           the SAME 80 accumulator columns: Phi += E v_hi^T ; Phi += E v_lo^T)
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

// Three-term bf16 split of a scalar (24 significant bits): v ~= t0 + t1 + t2.
__device__ __forceinline__ void split3_bf16(double v, __nv_bfloat16 &t0, __nv_bfloat16 &t1, __nv_bfloat16 &t2)
{
    t0 = __float2bfloat16_rn((float)v);
    double rem = isfinite(__bfloat162float(t0)) ? v - (double)__bfloat162float(t0) : 0.0;
    t1 = __float2bfloat16_rn((float)rem);
    rem -= (double)__bfloat162float(t1);
    t2 = __float2bfloat16_rn((float)rem);
}

// Operand rows for the tensor-core contractions, one warp per particle.  Both layouts are K = 3 x 64 bf16:
//     A_i = [ sa*hi_i | sa*lo_i | u1 u2 u3 1 1 1 0.. ]        B_j = [ sb*hi_j | sb*lo_j | 1 1 1 w1 w2 w3 0.. ]
// with y = scale * x~ = hi + lo (two bf16 terms) and the MMAs  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi + A_ex.B_ex
// so that the accumulator IS the quantity the epilogue needs (no per-pair arithmetic left):
//   MODE_DIST (scale = 1, sa = -2, sb = 1, u = w = |x~|^2):   S = |x~_i|^2 + |x~_j|^2 - 2 x~_i.x~_j = D2_ij
//   MODE_PHI  (scale = sqrt(2c), sa = sb = 1, u = 15 - c r_i, w = -c r_j):   S = log2( 2^15 k(x_j, x_i) )
// Padding rows (row >= n) get u = w = +inf in MODE_DIST (their distances never count) and zeros in MODE_PHI.
enum { SPLIT_DIST = 0, SPLIT_PHI = 1 };
__global__ void split_kernel(const double *__restrict__ X, const double *__restrict__ colsum, const double *__restrict__ a_ptr,
                             int64_t n, int64_t n_rows_alloc, int d, int mode, __nv_bfloat16 *__restrict__ XA,
                             __nv_bfloat16 *__restrict__ XB, double *__restrict__ rt)
{
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (row >= n_rows_alloc) return;
    const double c = mode == SPLIT_PHI ? (*a_ptr) * 1.4426950408889634 : 0.0; // a log2(e)
    const double scale = mode == SPLIT_PHI ? sqrt(2.0 * c) : 1.0;
    const float sa = mode == SPLIT_PHI ? 1.0f : -2.0f;
    __nv_bfloat16 *a = XA + row * TC_KTOT, *b = XB + row * TC_KTOT;
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int k = lane + 32 * h;
        double xc = 0.0;
        if (row < n && k < d) xc = X[row * d + k] - colsum[k] / (double)n;
        s += xc * xc;
        const double y = scale * xc;
        __nv_bfloat16 hi = __float2bfloat16_rn((float)y);
        __nv_bfloat16 lo = __float2bfloat16_rn((float)(y - (double)__bfloat162float(hi)));
        a[k] = __float2bfloat16_rn(sa * __bfloat162float(hi)); // exact: a power-of-two multiple
        a[64 + k] = __float2bfloat16_rn(sa * __bfloat162float(lo));
        b[k] = hi;
        b[64 + k] = lo;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && mode == SPLIT_DIST) rt[row] = s;
    // the 16 extra K columns (the remaining 48 of the chunk are never contracted, but keep them defined)
    double u, w;
    if (mode == SPLIT_DIST) { u = w = (row < n) ? s : (double)INFINITY; }
    else { u = (row < n) ? 15.0 - c * s : 0.0; w = (row < n) ? -c * s : 0.0; }
    __nv_bfloat16 u0, u1, u2, w0, w1, w2;
    split3_bf16(u, u0, u1, u2);
    split3_bf16(w, w0, w1, w2);
    const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int k = lane + 32 * h;
        __nv_bfloat16 av = zero, bv = zero;
        if (k == 0) { av = u0; bv = one; }
        if (k == 1) { av = u1; bv = one; }
        if (k == 2) { av = u2; bv = one; }
        if (k == 3) { av = one; bv = w0; }
        if (k == 4) { av = one; bv = w1; }
        if (k == 5) { av = one; bv = w2; }
        a[128 + k] = av;
        b[128 + k] = bv;
    }
}

// V^T (fp16, [TC_NV][ldn]) from the FP64 V = G - 2 a X (uncentred) of all particles:
// v~ = V + 2 a mean.  One block = 64 particles, transposed through shared memory.
__global__ void __launch_bounds__(256)
make_vt_kernel(const double *__restrict__ V, const double *__restrict__ colsum, const double *__restrict__ rt,
               const double *__restrict__ a_ptr, int64_t n, int64_t ldn, int d, __half *__restrict__ VT)
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
        tile[TC_NVH + rr][jl] = __float2half_rn(0.f); // the lo block carries no ones row
    }
    (void)rt;
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
// One CTA owns TWO 128-row i-tiles (256 particles) and a range of 128-column j-tiles.  Warps 0-3 / 4-7 are the
// exp warpgroups of i-tile 0 / 1 (thread = TMEM lane = row), warp 8 is the TMA producer, warp 9 issues every
// tcgen05.mma (both warp-uniform, one elected lane executes the instruction).  Both i-tiles contract against the
// same X_j / V_j tiles in shared memory, and the MMA order  PV0(t) S0(t+1) PV1(t) S1(t+1)  gives each warpgroup
// a full  PV + S  window for its 128x128 exponentials (FlashAttention-4 style ping-pong).
// The accumulator of the first contraction already is log2(2^15 k(x_j,x_i)) (see split_kernel), so the exp stage
// is one MUFU.EX2 per pair plus the fp16 pack.
//   TMEM   S_w [128 w, +128) fp32;  E_w = fp16 pairs over the first 64 columns of S_w;  Phi_w [256 + 80 w, +80)
//   smem   A_w 2 x 48 KB resident (hi | lo | extra);  X_j ring 4 x 16 KB chunks (hi, lo, extra per tile);
//          V_j ring 6 x 10 KB chunks (4 per tile: hi j[0,64) | hi j[64,128) | lo j[0,64) | lo j[64,128))
struct PhiTcArgs {
    float *phi_buf;      // [n_pad + 256][TC_PHI_LD], zeroed; partial sums are added atomically
    int64_t n_total, row0, n_rows;
    int n_jtiles, jsplit;
    int *err;
    long long *trace; // optional timeline of CTA 0 (development aid): [role][tile][event] clock64 values
};

// trace slots: role 0 = MMA issuer, 1 = exp WG0 thread 0, 2 = exp WG1 thread 0; 8 events per tile, 64 tiles
#define TC_TRACE(role, t, ev)                                                                      \
    do {                                                                                           \
        if (p.trace != nullptr && blockIdx.x == 0 && (t) < 64) p.trace[((role) * 64 + (t)) * 8 + (ev)] = clock64(); \
    } while (0)

constexpr uint32_t TC_CHUNK = 16384;                     // 128 rows x 128 B
constexpr uint32_t TC_A_BYTES = TC_KCH * TC_CHUNK;       // one resident X_i tile
constexpr int TC_NB = 4;                                 // X_j chunk ring
constexpr uint32_t TC_VCHUNK = TC_NVH * 128;             // 80 rows x 128 B
constexpr int TC_NVS = 6;                                // V chunk ring
constexpr uint32_t TC_PHI_SMEM = 2 * TC_A_BYTES + TC_NB * TC_CHUNK + TC_NVS * TC_VCHUNK + 512 + 1024;

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// S_w = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo + A_ex.B_ex for one i-tile: 13 MMAs (K = 16 each) issued by the elected
// lane as the X_j chunks (ring slots q = 3t + {0: hi, 1: lo, 2: extra}) become available.
template <class WaitFull, class Release>
__device__ __forceinline__ bool issue_dist_mmas(uint32_t d, uint32_t a_lo, uint32_t b_lo0, uint32_t idesc, int t, bool wait_b, bool release_b,
                                                WaitFull wait_full, Release release)
{
    const uint32_t a_hi_d = a_lo, a_lo_d = a_lo + (TC_CHUNK >> 4), a_ex_d = a_lo + 2 * (TC_CHUNK >> 4);
#pragma unroll
    for (int c = 0; c < TC_KCH; ++c) {
        const int q = TC_KCH * t + c, slot = q % TC_NB, ph = (q / TC_NB) & 1;
        if (wait_b && !wait_full(slot, ph)) return false;
        if (c == 0) tc_fence_after();
        const uint32_t bl = b_lo0 + slot * (TC_CHUNK >> 4);
        if (elect_one()) {
            if (c == 0) { // B_hi: against A_hi (starts the accumulation) and A_lo
                umma_f16_ss2<false>(d, a_hi_d, bl, idesc);
                umma_f16_ss2<true>(d, a_hi_d + 2, bl + 2, idesc);
                umma_f16_ss2<true>(d, a_hi_d + 4, bl + 4, idesc);
                umma_f16_ss2<true>(d, a_hi_d + 6, bl + 6, idesc);
                umma_f16_ss2<true>(d, a_lo_d, bl, idesc);
                umma_f16_ss2<true>(d, a_lo_d + 2, bl + 2, idesc);
                umma_f16_ss2<true>(d, a_lo_d + 4, bl + 4, idesc);
                umma_f16_ss2<true>(d, a_lo_d + 6, bl + 6, idesc);
            } else if (c == 1) { // B_lo: against A_hi
                umma_f16_ss2<true>(d, a_hi_d, bl, idesc);
                umma_f16_ss2<true>(d, a_hi_d + 2, bl + 2, idesc);
                umma_f16_ss2<true>(d, a_hi_d + 4, bl + 4, idesc);
                umma_f16_ss2<true>(d, a_hi_d + 6, bl + 6, idesc);
            } else { // 16 extra K columns: norms / exponent offsets
                umma_f16_ss2<true>(d, a_ex_d, bl, idesc);
            }
            if (release_b) release(slot);
        }
        __syncwarp();
    }
    return true;
}

__global__ void __launch_bounds__(320, 1)
phi_tc32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapV, PhiTcArgs p)
{
    const int ip = blockIdx.x / p.jsplit, js = blockIdx.x - ip * p.jsplit;
    const int tps = (p.n_jtiles + p.jsplit - 1) / p.jsplit;
    const int jbeg = js * tps;
    const int nt = min(p.n_jtiles, jbeg + tps) - jbeg;
    if (nt <= 0) return;
    const int64_t i0 = p.row0 + (int64_t)ip * (2 * TC_TILE);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                              // [2][TC_A_BYTES]
    uint8_t *sB = sA + 2 * TC_A_BYTES;               // [TC_NB][TC_CHUNK]
    uint8_t *sV = sB + TC_NB * TC_CHUNK;             // [TC_NVS][TC_VCHUNK]
    uint64_t *bars = (uint64_t *)(sV + TC_NVS * TC_VCHUNK);
    uint64_t *a_full = bars;                // 1
    uint64_t *b_full = bars + 1;            // TC_NB
    uint64_t *b_empty = b_full + TC_NB;     // TC_NB
    uint64_t *v_full = b_empty + TC_NB;     // TC_NVS
    uint64_t *v_empty = v_full + TC_NVS;    // TC_NVS
    uint64_t *s_full = v_empty + TC_NVS;    // 2
    uint64_t *e_ready = s_full + 2;         // 2
    uint64_t *phi_full = e_ready + 2;       // 1
    uint32_t *tmem_holder = (uint32_t *)(phi_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < TC_NB; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < TC_NVS; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(s_full + s, 1); mbar_init(e_ready + s, 128); }
        mbar_init(phi_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 8) { // ---- TMA producer: the whole warp runs the loop (uniform registers), one elected lane issues
        if (elect_one()) {
            mbar_arrive_expect_tx(a_full, 2 * TC_A_BYTES);
            for (int w = 0; w < 2; ++w)
                for (int c = 0; c < TC_KCH; ++c)
                    tma_load_2d(sA + w * TC_A_BYTES + c * TC_CHUNK, &mapA, c * 64, (int)(i0 + w * TC_TILE), a_full);
        }
        __syncwarp();
        bool ok = true;
        for (int t = 0; ok && t < nt; ++t) {
            const int j0 = (jbeg + t) * TC_TILE;
            for (int c = 0; ok && c < TC_KCH; ++c) {
                const int q = TC_KCH * t + c, slot = q % TC_NB, ph = (q / TC_NB) & 1;
                if (!mbar_wait(b_empty + slot, ph ^ 1, p.err, 10)) { ok = false; break; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(b_full + slot, TC_CHUNK);
                    tma_load_2d(sB + slot * TC_CHUNK, &mapB, c * 64, j0, b_full + slot);
                }
                __syncwarp();
            }
            for (int c = 0; ok && c < 4; ++c) {
                const int q = 4 * t + c, slot = q % TC_NVS, ph = (q / TC_NVS) & 1;
                if (!mbar_wait(v_empty + slot, ph ^ 1, p.err, 11)) { ok = false; break; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(v_full + slot, TC_VCHUNK);
                    tma_load_2d(sV + slot * TC_VCHUNK, &mapV, j0 + (c & 1) * 64, (c >> 1) * TC_NVH, v_full + slot);
                }
                __syncwarp();
            }
        }
    } else if (warp == 9) { // ---- MMA issuer: warp-uniform control flow, tcgen05 instructions from one elected lane
        const uint32_t idesc_s = make_idesc_bf16(TC_TILE, TC_TILE);
        const uint32_t idesc_v = make_idesc_f16(TC_TILE, TC_NVH);
        bool ok = mbar_wait(a_full, 0, p.err, 20);
        const uint32_t a_lo0 = desc_lo_k_sw128(smem_u32(sA)), b_lo0 = desc_lo_k_sw128(smem_u32(sB)), v_lo0 = desc_lo_k_sw128(smem_u32(sV));
        auto issue_s = [&](int w, int t) -> bool { // S_w(t)
            bool r = issue_dist_mmas(tmem + w * 128, a_lo0 + w * (TC_A_BYTES >> 4), b_lo0, idesc_s, t, w == 0, w == 1,
                                     [&](int slot, int ph) { return mbar_wait(b_full + slot, ph, p.err, 21); },
                                     [&](int slot) { umma_commit(b_empty + slot); });
            if (r && elect_one()) umma_commit(s_full + w);
            __syncwarp();
            return r;
        };
        auto issue_pv = [&](int w, int t) -> bool { // Phi_w += E_w(t) . [v_hi ; v_lo]
            if (!mbar_wait(e_ready + w, t & 1, p.err, 22 + w)) return false;
            const uint32_t d = tmem + 256 + w * TC_NVH, e = tmem + w * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int q = 4 * t + c, slot = q % TC_NVS, ph = (q / TC_NVS) & 1;
                if (w == 0 && !mbar_wait(v_full + slot, ph, p.err, 24)) return false;
                if (c == 0) tc_fence_after();
                const uint32_t vl = v_lo0 + slot * (TC_VCHUNK >> 4), ea = e + (c & 1) * 32;
                if (elect_one()) {
                    if (c == 0) umma_f16_ts2r(d, ea, vl, idesc_v, t ? 1u : 0u); else umma_f16_ts2<true>(d, ea, vl, idesc_v);
                    umma_f16_ts2<true>(d, ea + 8, vl + 2, idesc_v);
                    umma_f16_ts2<true>(d, ea + 16, vl + 4, idesc_v);
                    umma_f16_ts2<true>(d, ea + 24, vl + 6, idesc_v);
                    if (w == 1) umma_commit(v_empty + slot);
                }
                __syncwarp();
            }
            return true;
        };
        if (ok) ok = issue_s(0, 0) && issue_s(1, 0);
        for (int t = 0; ok && t < nt; ++t) {
            if (lane == 0) TC_TRACE(0, t, 0);
            ok = issue_pv(0, t);
            if (lane == 0) TC_TRACE(0, t, 1);
            if (ok && t + 1 < nt) ok = issue_s(0, t + 1);
            if (lane == 0) TC_TRACE(0, t, 2);
            if (ok) ok = issue_pv(1, t);
            if (lane == 0) TC_TRACE(0, t, 3);
            if (ok && t + 1 < nt) ok = issue_s(1, t + 1);
            if (lane == 0) TC_TRACE(0, t, 4);
        }
        if (ok && elect_one()) umma_commit(phi_full);
        __syncwarp();
    } else { // ---- exp warpgroups ------------------------------------------------------------------------------
        const int w = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + w * 128 + lane_base;
        const int64_t iw0 = i0 + w * TC_TILE;
        const int64_t i = iw0 + row;
        bool ok = true;
        for (int t = 0; ok && t < nt; ++t) {
            const int64_t j0 = (int64_t)(jbeg + t) * TC_TILE;
            const int dcol = (int)(i - j0); // column of k(x_i, x_i) in this tile, if inside [0,128)
            const bool tile_has_diag = (j0 < iw0 + TC_TILE) && (j0 + TC_TILE > iw0);
            if (row == 0) TC_TRACE(1 + w, t, 1);
            if (!mbar_wait(s_full + w, t & 1, p.err, 30 + w)) { ok = false; break; }
            if (row == 0) TC_TRACE(1 + w, t, 2);
            tc_fence_after();
            // 32-column chunk c: exponentials of rr[] -> fp16 pairs over S columns [16c, 16c+16) (already read)
            auto exp_chunk = [&](const uint32_t (&rr)[32], int c) {
                uint32_t packed[16];
                if (!tile_has_diag) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        packed[q] = pack_f16x2(ex2_approx(__uint_as_float(rr[2 * q])), ex2_approx(__uint_as_float(rr[2 * q + 1])));
                } else { // k(x_i, x_i) = exp(0) exactly, like the reference (2^15 after the fp16 scaling)
                    const int dq = dcol - c * 32;
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        float e0 = ex2_approx(__uint_as_float(rr[2 * q])), e1 = ex2_approx(__uint_as_float(rr[2 * q + 1]));
                        if (dq == 2 * q) e0 = 32768.0f;
                        if (dq == 2 * q + 1) e1 = 32768.0f;
                        packed[q] = pack_f16x2(e0, e1);
                    }
                }
                tmem_st16(tS + c * 16, packed);
            };
            uint32_t r0[32], r1[32];
            tmem_ld32(tS, r0);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                tmem_ld_wait();
                tmem_ld32(tS + (2 * cc + 1) * 32, r1); // overlaps the math on the even chunk
                exp_chunk(r0, 2 * cc);
                tmem_ld_wait();
                if (cc == 0) tmem_ld32(tS + 64, r0);
                exp_chunk(r1, 2 * cc + 1);
            }
            if (row == 0) TC_TRACE(1 + w, t, 3);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(e_ready + w);
            if (row == 0) TC_TRACE(1 + w, t, 4);
        }
        if (ok && mbar_wait(phi_full, 0, p.err, 40)) { // ---- flush Phi_w: TMEM -> global partial sums
            tc_fence_after();
            const bool valid = i < p.row0 + p.n_rows;
            float *dst = p.phi_buf + i * TC_PHI_LD;
            const uint32_t tP = tmem + 256 + w * TC_NVH + lane_base;
#pragma unroll 1
            for (int c0 = 0; c0 < TC_NVH; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tP + c0, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c0 + q <= TC_ONES_ROW) atomicAdd(dst + c0 + q, __uint_as_float(v[q]) * TC_E_UNSCALE);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// ---- pairwise squared distances on the tensor cores, for the exact median --------------------------------
// Same counting / collecting contract as dist_pass_f64_kernel (kernels_f64.cuh) on fp32 D2, which the MMA
// delivers directly (split_kernel, SPLIT_DIST: the row norms ride in the extra K columns; padding rows give
// +inf and never count).  Keys are the IEEE bits of (double)D2, so the host-side bracket logic and the radix
// select (select.cuh) are shared with the FP64 path; lo_f / hi_f are the float images of the key bounds
// (d2 >= lo_f  <=>  (double)d2 >= lo).
// Layout as in phi_tc32_kernel: two i-tiles per CTA, one counting warpgroup each, TWO S buffers per warpgroup
// (TMEM [256 w + 128 (t&1), +128)), X_j ring of 4 chunks.
struct DistTcArgs {
    int64_t n_total, row0, n_rows;
    int sym, n_jtiles, jsplit;
    int pair_offset, pair_stride; // this rank owns i-pairs offset, offset + stride, ... (cyclic: balances the triangle)
    float lo_f, hi_f;
    unsigned long long lo_key;
    int shift;
    unsigned long long *below, *hist, *cand, *cand_count; // (no max-below tracking: the host re-brackets instead)
    unsigned long long capacity;
    int *err;
    long long *trace;
};

constexpr int TC_WBUF = 512;  // candidate distances (fp32) staged per warp before one global reservation
constexpr int TC_PRIV = 40;   // per-thread staging slots: compacted once a thread holds more than TC_PRIV - 32
constexpr uint32_t TC_DIST_SMEM_BASE = 2 * TC_A_BYTES + TC_NB * TC_CHUNK + 8 * TC_WBUF * 4 + 256 * TC_PRIV * 4 + 512 + 1024;
constexpr uint32_t TC_DIST_SMEM_HIST = TC_DIST_SMEM_BASE; // the histogram aliases the (unused) warp staging buffers
static_assert(HIST_BINS * 4 <= 8 * TC_WBUF * 4, "histogram must fit the warp staging area");

__device__ __forceinline__ unsigned long long dist_key(float d2)
{
    return (unsigned long long)__double_as_longlong((double)fmaxf(d2, 0.0f));
}
// staging bypass for a chunk that overflows the warp buffer (very wide bracket): reserve straight in the global list
__device__ __noinline__ void dist_append_global(float d2, unsigned int wgt, const DistTcArgs *p)
{
    const unsigned long long key = dist_key(d2);
    const unsigned long long g = atomicAdd(p->cand_count, (unsigned long long)wgt);
    if (g < p->capacity) p->cand[g] = key;
    if (wgt == 2u && g + 1 < p->capacity) p->cand[g + 1] = key;
}

// warp-collective and out of line: move `count` staged distances to the global candidate list as keys
__device__ __noinline__ void dist_flush(const float *mybuf, unsigned int count, const DistTcArgs *p)
{
    __syncwarp();
    const unsigned int lane = threadIdx.x & 31;
    unsigned long long base = 0ull;
    if (lane == 0 && count) base = atomicAdd(p->cand_count, (unsigned long long)count);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned int q = lane; q < count; q += 32)
        if (base + q < p->capacity) p->cand[base + q] = dist_key(mybuf[q]);
    __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(320, 1)
dist_tc32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ DistTcArgs p)
{
    const int ipl = blockIdx.x / p.jsplit, js = blockIdx.x - ipl * p.jsplit;
    const int ip = p.pair_offset + p.pair_stride * ipl;
    const int jfirst = p.sym ? 2 * ip : 0; // tile-level upper triangle when symmetric
    const int len = p.n_jtiles - jfirst;
    const int tps = (len + p.jsplit - 1) / p.jsplit;
    const int jbeg = jfirst + js * tps;
    const int nt = min(p.n_jtiles, jbeg + tps) - jbeg;
    if (nt <= 0) return;
    const int64_t i0 = p.row0 + (int64_t)ip * (2 * TC_TILE);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + 2 * TC_A_BYTES;
    float *wbuf = (float *)(sB + TC_NB * TC_CHUNK); // [8][TC_WBUF]   per-warp staging
    float *priv = wbuf + 8 * TC_WBUF;               // [8][TC_PRIV][32] per-thread staging, lane-interleaved
    uint64_t *bars = (uint64_t *)(priv + 256 * TC_PRIV);
    uint64_t *a_full = bars;
    uint64_t *b_full = bars + 1;
    uint64_t *b_empty = b_full + TC_NB;
    uint64_t *s_full = b_empty + TC_NB;  // [2 wg][2 buf]
    uint64_t *s_free = s_full + 4;       // [2 wg][2 buf]
    uint32_t *tmem_holder = (uint32_t *)(s_free + 4);
    unsigned int *shist = (unsigned int *)wbuf; // [HIST_BINS] (MODE_HIST only: aliases the warp staging buffers)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < TC_NB; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(s_full + s, 1); mbar_init(s_free + s, 128); }
        fence_barrier_init();
    }
    if (MODE == MODE_HIST)
        for (int b = threadIdx.x; b < HIST_BINS; b += blockDim.x) shist[b] = 0u;
    if (warp == 8) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 8) { // ---- TMA producer (whole warp, one elected lane issues)
        if (elect_one()) {
            mbar_arrive_expect_tx(a_full, 2 * TC_A_BYTES);
            for (int w = 0; w < 2; ++w)
                for (int c = 0; c < TC_KCH; ++c)
                    tma_load_2d(sA + w * TC_A_BYTES + c * TC_CHUNK, &mapA, c * 64, (int)(i0 + w * TC_TILE), a_full);
        }
        __syncwarp();
        bool ok = true;
        for (int t = 0; ok && t < nt; ++t) {
            const int j0 = (jbeg + t) * TC_TILE;
            for (int c = 0; ok && c < TC_KCH; ++c) {
                const int q = TC_KCH * t + c, slot = q % TC_NB, ph = (q / TC_NB) & 1;
                if (!mbar_wait(b_empty + slot, ph ^ 1, p.err, 50)) { ok = false; break; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(b_full + slot, TC_CHUNK);
                    tma_load_2d(sB + slot * TC_CHUNK, &mapB, c * 64, j0, b_full + slot);
                }
                __syncwarp();
            }
        }
    } else if (warp == 9) { // ---- MMA issuer: S_w(t) into buffer (w, t & 1); warp-uniform, one elected lane issues
        const uint32_t idesc_s = make_idesc_bf16(TC_TILE, TC_TILE);
        bool ok = mbar_wait(a_full, 0, p.err, 60);
        const uint32_t a_lo0 = desc_lo_k_sw128(smem_u32(sA)), b_lo0 = desc_lo_k_sw128(smem_u32(sB));
        for (int t = 0; ok && t < nt; ++t) {
            const int buf = t & 1, bph = (t >> 1) & 1;
            for (int w = 0; ok && w < 2; ++w) {
                if (lane == 0) TC_TRACE(0, t, 3 * w);
                if (!mbar_wait(s_free + 2 * w + buf, bph ^ 1, p.err, 62)) { ok = false; break; }
                if (lane == 0) TC_TRACE(0, t, 3 * w + 1);
                ok = issue_dist_mmas(tmem + w * 256 + buf * 128, a_lo0 + w * (TC_A_BYTES >> 4), b_lo0, idesc_s, t, w == 0, w == 1,
                                     [&](int slot, int ph) { return mbar_wait(b_full + slot, ph, p.err, 61); },
                                     [&](int slot) { umma_commit(b_empty + slot); });
                if (ok && elect_one()) umma_commit(s_full + 2 * w + buf);
                __syncwarp();
                if (lane == 0) TC_TRACE(0, t, 3 * w + 2);
            }
        }
    } else { // ---- counting warpgroups: thread = row i
        const int w = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int itile = 2 * ip + w;
        const int64_t iw0 = i0 + w * TC_TILE;
        const int64_t i = iw0 + row;
        const bool row_valid = i < p.row0 + p.n_rows;
        float *mybuf = wbuf + warp * TC_WBUF;
        const uint32_t priv_base = smem_u32(priv + warp * (TC_PRIV * 32) + lane); // this thread's lane-interleaved slots
        const uint32_t wbuf_base = smem_u32(mybuf);
        unsigned int count = 0; // warp-uniform fill level of mybuf
        unsigned long long below = 0ull;
        // rows outside this rank's range never count: give them an empty bracket with nothing below it
        const float lo = row_valid ? p.lo_f : -INFINITY, hi = row_valid ? p.hi_f : -INFINITY;
        uint32_t paddr = priv_base; // next free private slot (slot e of this thread lives at priv_base + 128 e)
        unsigned int cur_wgt = 1u;  // weight of the entries currently staged in the private slots
        // warp-collective: move the private entries (all of weight cur_wgt) into the warp buffer / histogram
        auto compact = [&]() {
            const uint32_t mine = (paddr - priv_base) >> 7;
            const unsigned int tot = __reduce_add_sync(0xffffffffu, mine);
            if (tot) { // warp-uniform
                if (MODE == MODE_HIST) {
                    for (uint32_t e = 0; e < mine; ++e)
                        atomicAdd(&shist[(unsigned int)((dist_key(lds_f32(priv_base + 128u * e)) - p.lo_key) >> p.shift)], cur_wgt);
                } else {
                    if (count + tot * cur_wgt > (unsigned int)TC_WBUF) { dist_flush(mybuf, count, &p); count = 0; }
                    if (tot * cur_wgt > (unsigned int)TC_WBUF) { // more than an empty buffer holds: straight to global
                        for (uint32_t e = 0; e < mine; ++e) dist_append_global(lds_f32(priv_base + 128u * e), cur_wgt, &p);
                    } else {
                        uint32_t incl = mine; // inclusive scan over lanes
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        uint32_t dst = wbuf_base + 4u * (count + (incl - mine) * cur_wgt);
                        for (uint32_t e = 0; e < mine; ++e) {
                            const float v = lds_f32(priv_base + 128u * e);
                            sts_f32(dst, v); dst += 4u;
                            if (cur_wgt == 2u) { sts_f32(dst, v); dst += 4u; }
                        }
                        count += tot * cur_wgt;
                    }
                }
                paddr = priv_base;
                __syncwarp();
            }
        };
        for (int t = 0; t < nt; ++t) {
            const int buf = t & 1, bph = (t >> 1) & 1;
            const int tj = jbeg + t;
            const int64_t j0 = (int64_t)tj * TC_TILE;
            // symmetric mode: tiles below the diagonal are covered by their transposes (weight 2)
            const unsigned int wgt = !p.sym ? 1u : (tj < itile ? 0u : (tj == itile ? 1u : 2u));
            const bool tile_has_diag = (j0 < iw0 + TC_TILE) && (j0 + TC_TILE > iw0);
            const int dcol = (int)(i - j0);
            if (row == 0) TC_TRACE(1 + w, t, 0);
            if (!mbar_wait(s_full + 2 * w + buf, bph, p.err, 70 + w)) break;
            if (row == 0) TC_TRACE(1 + w, t, 1);
            tc_fence_after();
            const uint32_t tS = tmem + w * 256 + buf * 128 + lane_base;
            if (wgt == 0u) { // nothing to count: hand the buffer straight back
                tc_fence_before();
                mbar_arrive(s_free + 2 * w + buf);
                continue;
            }
            if (wgt != cur_wgt) { compact(); cur_wgt = wgt; }
            unsigned int cnt4[4] = {0u, 0u, 0u, 0u}; // four independent counters: no 128-long dependent add chain
            // Per distance: two compares, a predicated count and a predicated store into the thread's private staging
            // column (no branch, no vote: a vote + branch per element serialises the warp at ~80 cycles each).
            auto visit = [&](float d2, unsigned int &cnt) {
                asm volatile("{\n\t.reg .pred pb, pi;\n\t"
                             "setp.lt.f32 pb, %2, %3;\n\t"
                             "@pb add.u32 %0, %0, 1;\n\t"
                             "setp.lt.and.f32 pi, %2, %4, !pb;\n\t"
                             "@pi st.shared.f32 [%1], %2;\n\t"
                             "@pi add.u32 %1, %1, 128;\n\t}"
                             : "+r"(cnt), "+r"(paddr)
                             : "f"(d2), "f"(lo), "f"(hi)
                             : "memory");
            };
            auto count_chunk = [&](const uint32_t (&rr)[32], int c) {
                if (!tile_has_diag) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) visit(__uint_as_float(rr[q]), cnt4[q & 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) visit(dcol == c * 32 + q ? 0.0f : __uint_as_float(rr[q]), cnt4[q & 3]); // |x_i - x_i|^2 = 0 exactly
                }
                // the next chunk may add up to 32 entries per thread: compact when any thread could overflow
                if (__any_sync(0xffffffffu, paddr - priv_base > (uint32_t)(TC_PRIV - 32) * 128u)) compact();
            };
            uint32_t r0[32], r1[32];
            tmem_ld32(tS, r0);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                tmem_ld_wait();
                tmem_ld32(tS + (2 * cc + 1) * 32, r1);
                count_chunk(r0, 2 * cc);
                tmem_ld_wait();
                if (cc == 0) {
                    tmem_ld32(tS + 64, r0);
                } else { // S is in registers: hand the buffer back to the MMA issuer
                    tc_fence_before();
                    mbar_arrive(s_free + 2 * w + buf);
                }
                count_chunk(r1, 2 * cc + 1);
            }
            below += (unsigned long long)(cnt4[0] + cnt4[1] + cnt4[2] + cnt4[3]) * wgt;
            if (row == 0) TC_TRACE(1 + w, t, 2);
        }
        compact();
        if (MODE == MODE_COLLECT && count) dist_flush(mybuf, count, &p);
        for (int o = 16; o; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below) atomicAdd(p.below, below);
    }
    tc_fence_before();
    __syncthreads();
    if (MODE == MODE_HIST)
        for (int b = threadIdx.x; b < HIST_BINS; b += blockDim.x) {
            unsigned int c = shist[b];
            if (c) atomicAdd(&p.hist[b], (unsigned long long)c);
        }
    if (warp == 8) tmem_dealloc(tmem, 512);
}

} // namespace tc
} // namespace svgdb
