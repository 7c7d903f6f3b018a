// kernels_phi_tc.cuh — the persistent tensor-core pair-interaction kernel of SVGDB_PRECISION_TC32 (sm_100a).
//
//   phi_i = (1/n) [ sum_j k(x_j,x_i) v_j + 2 a x~_i sum_j k(x_j,x_i) ],  v_j = g_j - 2 a x~_j        (SVGD.hpp:407-454)
//
// Arithmetic (error budget: DESIGN.md "Precision modes"):
//   * y = sqrt(2c) (x - mean), c = a log2(e).  Row side  y_i = hi_i + lo_i  (two fp16 terms, 22 bits),
//     column side  y^_j = fp16(y_j).  The first contraction is  S = y_i . y^_j  (fp16 products are exact in
//     fp32, fp32 accumulation in TMEM) and the exponent  S + u_i + w_j  with  u_i = 15 - |y_i|^2/2,
//     w_j = -|y^_j|^2/2  equals  15 - |y_i - y^_j|^2 / 2 <= 15: the kernel matrix is the EXACT Gaussian
//     kernel between the particles and their fp16-rounded images (2^-12 relative per coordinate).
//   * E = 2^15 k rounded to fp16 (normal down to k = 2^-29), one MUFU.EX2 per pair (optionally a share of
//     them as a Cody-Waite + degree-4 polynomial on the FMA pipe); the row sum adds the SAME rounded values
//     (FHADD), so k(x_i,x_i) = 1 cancels exactly in the repulsive term.
//   * second contraction  Phi += E . [v_hi ; v_lo]  (v split in two fp16 terms).
//
// Mapping onto the SM (one persistent CTA per SM, 10 warps):
//   * TMEM (512 columns):  S_b [64 b, +64) fp32 for the four (i-tile w, j-half k) units b = 2w + k (E_b aliases its
//     first 32 columns as fp16 pairs), Phi_w [256 + 64 w, +64) fp32,  A_w [384 + 64 w, +64) = the row operand [hi | lo] of i-tile w, written
//     ONCE per segment by the exp warps (tcgen05.st) -> every MMA runs in TS mode: only the column operand
//     is read from shared memory (64 B/clk instead of the 128 B/clk an SS MMA at M = N = 128 needs, which is
//     the whole shared-memory bandwidth of the SM and was the limiter of the first version of this kernel).
//   * shared memory: 4 stages of { X^_j tile 16 KB | V_j tile 4 x 8 KB | w_j 512 B }, one TMA producer warp.
//   * warp 9 issues every tcgen05.mma (M=128, N=64, K=16): per unit 8 MMAs for S and 8 for Phi, four units in
//     flight (see the kernel's comment); warps 0-3 / 4-7 are the exp warpgroups of i-tile 0 / 1.
//   * work: the (i-pair, j-tile) rectangle is linearised and cut into one contiguous range per CTA; a CTA
//     walks its range segment by segment (segment = one i-pair), so operand load, pipeline fill and the
//     flush of Phi happen ~3 times per SM instead of once per (i-pair, j-split) CTA.
#pragma once
#include "kernels_tc32.cuh"
#include <cuda_fp8.h>
#include <type_traits>

namespace svgdb {
namespace tc {

constexpr uint32_t P2_VBOX = 8192;                    // 64 coordinates x 64 particles fp16
constexpr uint32_t P2_V_BYTES = 4 * P2_VBOX;          // hi j[0,64) | hi j[64,128) | lo j[0,64) | lo j[64,128)
constexpr uint32_t P2_W_BYTES = 4096;                 // 128 particles x 16 fp16 exponent-offset columns (no-swizzle core-matrix order)
constexpr uint32_t P2_AEX_BYTES = 4096;               // per i-tile: 128 rows x 16 fp16 exponent-offset columns
// Two arithmetic variants of the pair kernel (DESIGN.md "Precision modes"):
//   fast    : S = (hi_i + lo_i) . hi_j,  E = fp16(2^15 k),  Phi += E . (v_hi + v_lo)                  16 + 1 MMAs per 128 x 64 unit
//   PRECISE : S = hi_i.hi_j + lo_i.hi_j + hi_i.lo_j (both particles 22 bits),  E = E_hi + E_lo (two fp16 terms),
//             Phi += E_hi.v_hi + E_hi.v_lo + E_lo.v_hi                                                  24 + 1 MMAs per unit
//   F8LO (FAST only): the two correction terms -- lo_i . y^_j of the first contraction and E . v_lo of the second, each 2^-11 of its
//             main term -- run as kind::f8f6f4 on e5m2 copies (lo_i 2^10 and y^_j 2^-10; the top byte of the fp16 E and e5m2(v_lo)):
//             13 MMAs per unit instead of 17; adds zero-mean noise of ~1e-4 relative per kernel value / 3e-5 per product.
template <bool PRECISE, bool F8LO = false>
struct P2Cfg {
    static_assert(!(PRECISE && F8LO), "the fp8 term belongs to the FAST variant");
    static constexpr int STAGES = (PRECISE || F8LO) ? 3 : 4;
    static constexpr uint32_t XB_BYTES = PRECISE ? 32768u : 16384u; // 128 particles x 64 fp16 hi (+ 64 fp16 lo): 128 B rows, SWIZZLE_128B boxes
    static constexpr uint32_t X8_BYTES = F8LO ? 8192u : 0u;         // 128 particles x 64 e5m2: 64 B rows, SWIZZLE_64B
    static constexpr uint32_t V8_BYTES = F8LO ? 8192u : 0u;         // v_lo as e5m2: two boxes (j-halves) of 64 coordinates x 64 particles
    static constexpr uint32_t STAGE = XB_BYTES + P2_V_BYTES + P2_W_BYTES + X8_BYTES + V8_BYTES; // 52 KB / 68 KB / 68 KB (1024-aligned)
    static constexpr uint32_t TX = F8LO ? STAGE - 2 * P2_VBOX : STAGE; // (the fp16 v_lo boxes are not fetched in the F8LO variant)
    static constexpr uint32_t ONES_BYTES = F8LO ? 2048u : 0u;      // 16 rows x 64 fp16 ones: the B operand of the row-sum MMAs (TCSUM)
    static constexpr uint32_t SMEM = STAGES * STAGE + 2 * P2_AEX_BYTES + ONES_BYTES + 256 + 1024;
};
// K-major operand of 16 fp16 columns WITHOUT swizzle: 8 x 16 B core matrices; row r, column k lives at
// (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2   (LBO = 128 B between the two K halves, SBO = 256 B per 8 rows)
__host__ __device__ constexpr uint32_t p2_ex_offset(uint32_t r, uint32_t k) { return (r >> 3) * 256u + (k >> 3) * 128u + (r & 7u) * 16u + (k & 7u) * 2u; }
constexpr uint32_t DESC_HI_K_NOSW = (256u >> 4) | (1u << 14);  // SBO = 256 B, version 1, SWIZZLE_NONE
constexpr uint32_t DESC_LO_K_NOSW_LBO = (128u >> 4) << 16;     // LBO = 128 B
constexpr int P2_A_LD = 128;                          // XA2 row: [hi(64) | lo(64)] fp16
constexpr uint32_t P2_COL_PHI = 256, P2_COL_A = 384;

// ---- operand preparation ---------------------------------------------------------------------------
// Three-term fp16 split of a scalar (33 significant bits): v ~= t0 + t1 + t2.
__device__ __forceinline__ void split3_f16(double v, __half &t0, __half &t1, __half &t2)
{
    v = fmax(v, -60000.0); // below every exponent that matters, above the fp16 range limit
    t0 = __double2half(v);
    double rem = v - (double)__half2float(t0);
    t1 = __double2half(rem);
    rem -= (double)__half2float(t1);
    t2 = __double2half(rem);
}

// One warp per particle.  XA2[row] = [hi | lo] (row operand, read by the owning thread into TMEM),
// XB2[row] = hi (column operand, TMA).  The exponent offsets u = 15 - |hi+lo|^2/2 (row) and w = -|hi|^2/2
// (column; -60000 for padding rows, whose kernel values are then exactly 0) ride in a 16-column K chunk
//     UA[row] = [u0 u1 u2 1 1 1 0..]      WB[row] = [1 1 1 w0 w1 w2 0..]      (three-term fp16 splits)
// so that the accumulator of the first contraction IS the exponent.  WB is stored per 128-particle tile in the
// core-matrix order the MMA reads (p2_ex_offset), UA as plain rows.
// precise != 0: the column operand keeps both terms, XB2[row] = [hi | lo] (128 columns per row), and w = -|hi + lo|^2/2.
// XB8 != nullptr (F8LO): additionally the e5m2 copies -- XA2[row] bytes [128, 192) = e5m2(lo 2^10), XB8[row] = e5m2(hi 2^-10).
__global__ void split_phi2_kernel(const double *__restrict__ X, const double *__restrict__ colsum, const double *__restrict__ a_ptr,
                                  int64_t n, int64_t n_rows_a, int64_t n_rows_b, int d, __half *__restrict__ XA2,
                                  __half *__restrict__ XB2, __half *__restrict__ UA, __half *__restrict__ WB, int precise,
                                  uint8_t *__restrict__ XB8 = nullptr)
{
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows_a) return;
    const double c = (*a_ptr) * 1.4426950408889634; // a log2(e)
    const double scale = sqrt(2.0 * c);
    double s_full = 0.0, s_hi = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        double y = 0.0;
        if (row < n && k < d) y = scale * (X[row * d + k] - colsum[k] / (double)n);
        const __half hi = __double2half(y);
        const double hid = (double)__half2float(hi);
        const __half lo = __double2half(y - hid);
        const double full = hid + (double)__half2float(lo);
        s_full += full * full;
        s_hi += hid * hid;
        XA2[row * P2_A_LD + k] = hi;
        if (XB8 != nullptr) {
            reinterpret_cast<uint8_t *>(XA2 + row * P2_A_LD + 64)[k] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(lo) * 1024.0f, __NV_SATFINITE, __NV_E5M2);
            if (row < n_rows_b) XB8[row * 64 + k] = (uint8_t)__nv_cvt_float_to_fp8((float)hid * (1.0f / 1024.0f), __NV_SATFINITE, __NV_E5M2);
        } else {
            XA2[row * P2_A_LD + 64 + k] = lo;
        }
        if (row < n_rows_b) {
            if (precise) {
                XB2[row * 128 + k] = hi;
                XB2[row * 128 + 64 + k] = lo;
            } else {
                XB2[row * 64 + k] = hi;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        s_full += __shfl_xor_sync(0xffffffffu, s_full, o);
        s_hi += __shfl_xor_sync(0xffffffffu, s_hi, o);
    }
    if (lane < 16) {
        __half u0, u1, u2, w0, w1, w2;
        split3_f16((row < n) ? 15.0 - 0.5 * s_full : 0.0, u0, u1, u2);
        split3_f16((row < n) ? -0.5 * (precise ? s_full : s_hi) : -60000.0, w0, w1, w2);
        const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
        const __half ua = lane == 0 ? u0 : lane == 1 ? u1 : lane == 2 ? u2 : lane < 6 ? one : zero;
        const __half wb = lane < 3 ? one : lane == 3 ? w0 : lane == 4 ? w1 : lane == 5 ? w2 : zero;
        UA[row * 16 + lane] = ua;
        if (row < n_rows_b)
            *reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(WB) + (row >> 7) * P2_W_BYTES + p2_ex_offset((uint32_t)(row & 127), (uint32_t)lane)) = wb;
    }
}

// v~ = g - 2 a (x - mean) for this rank's rows, rounded once to fp32 (24 bits: more than the two fp16 terms it is split
// into keep); centred here, in FP64, so that a far-off particle cloud costs no precision.  V32 is indexed by global row:
// it is what the ranks all-gather (half the bytes of the FP64 V of the other precision mode).
__global__ void make_v32_kernel(const double *__restrict__ X, const double *__restrict__ G, const double *__restrict__ colsum,
                                const double *__restrict__ a_ptr, int64_t n, int64_t row0, int64_t n_rows, int d, float *__restrict__ V32)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * d) return;
    const int c = (int)(idx % d);
    const double a = *a_ptr;
    V32[row0 * d + idx] = (float)(G[idx] - 2.0 * a * (X[row0 * d + idx] - colsum[c] / (double)n));
}

// V^T (fp16, [128][ldn]): rows [0,64) v_hi, [64,128) v_lo of v~.  One block = 64 particles, transposed through shared memory.
// VT8 != nullptr (F8LO): additionally v_lo as e5m2, [64][ldn] bytes.
__global__ void __launch_bounds__(256)
make_vt2_kernel(const float *__restrict__ V32, int64_t n, int64_t ldn, int d, __half *__restrict__ VT, uint8_t *__restrict__ VT8 = nullptr)
{
    __shared__ __half tile[128][64 + 2];
    const int64_t j0 = (int64_t)blockIdx.x * 64;
    for (int t = threadIdx.x; t < 64 * 64; t += blockDim.x) {
        const int jl = t >> 6, c = t & 63;
        const int64_t j = j0 + jl;
        __half hi = __float2half_rn(0.f), lo = hi;
        if (j < n && c < d) {
            const float v = V32[j * d + c];
            hi = __float2half_rn(v);
            lo = __float2half_rn(v - __half2float(hi));
        }
        tile[c][jl] = hi;
        tile[64 + c][jl] = lo;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 128 * 64; t += blockDim.x) {
        const int rr = t >> 6, jl = t & 63;
        if (j0 + jl < ldn) {
            VT[(int64_t)rr * ldn + j0 + jl] = tile[rr][jl];
            if (VT8 != nullptr && rr >= 64)
                VT8[(int64_t)(rr - 64) * ldn + j0 + jl] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(tile[rr][jl]), __NV_SATFINITE, __NV_E5M2);
        }
    }
}

// ---- device helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// SS MMA with explicit descriptor words for both operands (any layout), accumulating
__device__ __forceinline__ void umma_f16_ss_desc(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc)
{
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
                 "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "setp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
                 : "memory");
}
// acc += lo half, acc2 += hi half of a packed f16x2 (FHADD: fp32 accumulation of the ROUNDED values)
__device__ __forceinline__ void acc_f16x2(float &acc_lo, float &acc_hi, uint32_t p)
{
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tadd.rn.f32.f16 %0, l, %0;\n\tadd.rn.f32.f16 %1, h, %1;\n\t}"
        : "+f"(acc_lo), "+f"(acc_hi)
        : "r"(p));
}
// fp16x2( e_lo - float(p.lo), e_hi - float(p.hi) ): the second fp16 term of a value whose first term is already packed in p
// (FHFMA: fp16 x fp16 + fp32, one instruction per element)
__device__ __forceinline__ uint32_t residual_f16x2(uint32_t p, float e_lo, float e_hi)
{
    float r0, r1;
    asm("{\n\t.reg .f16 a, b, m;\n\tmov.b32 {a, b}, %2;\n\tmov.b16 m, 0xBC00;\n\t"
        "fma.rn.f32.f16 %0, a, m, %3;\n\tfma.rn.f32.f16 %1, b, m, %4;\n\t}"
        : "=f"(r0), "=f"(r1)
        : "r"(p), "f"(e_lo), "f"(e_hi));
    return pack_f16x2(r0, r1);
}
// 2^x for x <= 15 on the FMA / ALU pipes: round-to-nearest split x = n + f (magic-number add), degree-4
// polynomial for 2^f on [-1/2, 1/2], exponent add.  x is clamped at
// -126 (results that small round to +0 in fp16 anyway; -inf from padding columns lands there too).
__device__ __forceinline__ float ex2_poly(float x)
{
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f; // 1.5 * 2^23: the integer part of x now sits in the low mantissa bits
    const float f = x - (t - 12582912.0f);
    float p = 9.666368515e-3f; // Chebyshev-node fit of 2^f on [-1/2, 1/2]: max rel error 3.6e-6 (fp16 rounding: 2.4e-4)
    p = fmaf(p, f, 5.592197584e-2f);
    p = fmaf(p, f, 2.402234904e-1f);
    p = fmaf(p, f, 6.931210452e-1f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

struct Phi2Args {
    float *phi_buf;        // [n_pad128 + 256][TC_PHI_LD], zeroed; [0,64) sum_j E v, [64] sum_j E; added atomically
    const __half *XA2;     // [n_pad128 + 256][128]
    const __half *UA;      // [n_pad128 + 256][16] row exponent-offset chunk
    const __half *WB;      // [n_pad128 / 128][4 KB] column exponent-offset chunks, core-matrix order
    int64_t row0, n_rows;  // this rank's rows
    int n_jtiles, n_ipairs;
    int max_seg;           // longest run of j-tiles accumulated in TMEM before the partial sums are flushed (see p2_segment)
    int poly;              // pairs (of 16) per 32-column chunk whose exponentials use ex2_poly
    int dbg;               // development: 1 = exp warps only hand the barriers on, 2 = TMEM load/store without the math
    int no_vlo;            // F8LO only: 1 = leave the E . v_lo term out altogether (v then carries one fp16 term, like E)
    int *err;
    long long *trace;
};

// segment s of CTA b: i-pair `ip`, j-tiles [jb, je)
struct P2Seg { int ip, jb, je; };
__device__ __forceinline__ bool p2_segment(const Phi2Args &p, long long &pos, long long end, P2Seg &s)
{
    if (pos >= end) return false;
    s.ip = (int)(pos / p.n_jtiles);
    s.jb = (int)(pos - (long long)s.ip * p.n_jtiles);
    // The tensor core adds into its fp32 accumulator with truncation: the bias of a sum grows with the number of additions made on
    // top of it.  Segments are therefore capped; the partial sums of an i-pair are combined by (round-to-nearest) atomics in phi_buf.
    const long long seg_end = min(min(end, (long long)(s.ip + 1) * p.n_jtiles), pos + p.max_seg);
    s.je = s.jb + (int)(seg_end - pos);
    pos = seg_end;
    return true;
}

// POLY of the 16 pairs of a 32-column chunk go through ex2_poly (0 = all MUFU).
//
// Pipeline units.  A (i-tile w, j-half k) pair of 128 x 64 pairs is one unit with its own S buffer b = 2w + k
// (TMEM columns [64 b, +64); E_b aliases its first 32 columns).  Per j-tile the MMA warp issues
//     [PV(b0) S'(b0)] [PV(b1) S'(b1)] [PV(b2) S'(b2)] [PV(b3) S'(b3)]          (S' = the next j-tile's S)
// so between the completion of S(b) and the issue of PV(b) lie three other units (1536 tensor-pipe cycles):
// the commit -> mbarrier -> exp warps -> mbarrier -> MMA warp round trip (~1.5k cycles measured with idle exp
// warps) no longer starves the tensor pipe, which it did with two 128-column buffers (one unit of slack).
// Sixteen exp warps: warp = (i-tile w, 32-column half h of a unit, row quadrant q) = 8 w + 4 h + q, i.e. four resident exp warps per
// scheduler (with two, a warp spends ~30 % of its time in dependent-issue and tcgen05.ld/st latencies that nothing fills);
// then the TMA producer and one MMA issuer per i-tile.
constexpr int P2_EWARPS = 16;
constexpr int P2_THREADS = (P2_EWARPS + 3) * 32;

// CL = 2: clusters of two CTAs walk the SAME j-tiles with two consecutive i-pairs (512 particle rows per cluster); every box of a
// stage is fetched from L2 once and multicast into both CTAs (each issues half of the boxes), a stage is refilled when the MMAs of
// both CTAs have released it.  The kernel pulls 7 GB per launch from L2 at the headline shape (4.1 TB/s): this halves it.
// TCSUM (with F8LO): the row sum Sum_j E_ij comes from the tensor core too -- four N = 16 MMAs per unit of E against a block of ones,
// accumulated in the 16 TMEM columns the e5m2 lo operand leaves free -- instead of one FHADD per pair in the exp warps.
template <int POLY, bool PRECISE, int CL = 1, bool F8LO = false, bool TCSUM = false>
__global__ void __launch_bounds__(P2_THREADS, 1)
phi2_tc32_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapV, const __grid_constant__ Phi2Args p,
                 const __grid_constant__ CUtensorMap mapB8, const __grid_constant__ CUtensorMap mapV8)
{
    using Cfg = P2Cfg<PRECISE, F8LO>;
    constexpr int P2_STAGES = Cfg::STAGES;
    constexpr uint32_t P2_XB_BYTES = Cfg::XB_BYTES, P2_STAGE = Cfg::STAGE, P2_TX = Cfg::TX;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);
    // the cluster's contiguous range of (i-pair group, j-tile) work units (p.n_ipairs counts groups of CL i-pairs; CTA r takes pair CL g + r)
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const long long n_clusters = gridDim.x / CL, cluster_id = blockIdx.x / CL;
    const long long units = (long long)p.n_ipairs * p.n_jtiles;
    const long long u_beg = units * cluster_id / n_clusters, u_end = units * (cluster_id + 1) / n_clusters;
    if (CL == 1 && u_beg >= u_end) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sAex = smem + P2_STAGES * P2_STAGE; // [2][P2_AEX_BYTES]
    uint8_t *sOnes = sAex + 2 * P2_AEX_BYTES;    // [Cfg::ONES_BYTES], 1024-aligned
    uint64_t *bars = (uint64_t *)(sOnes + Cfg::ONES_BYTES);
    uint64_t *full = bars;                  // P2_STAGES: TMA bytes landed
    uint64_t *empty = full + P2_STAGES;     // P2_STAGES: every MMA reading the stage has completed
    uint64_t *s_full = empty + P2_STAGES;   // 4: S_b complete
    uint64_t *e_ready = s_full + 4;         // 4: E_b written (one arrival per exp warp of the tile)
    uint64_t *phi_full = e_ready + 4;       // [2] every MMA of the segment on tile w complete
    uint64_t *a_ready = phi_full + 2;       // [2] row operand of tile w in TMEM, Phi_w flushed (one arrival per exp warp)
    uint32_t *tmem_holder = (uint32_t *)(a_ready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < P2_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 2 * CL); }
        for (int s = 0; s < 4; ++s) { mbar_init(s_full + s, 1); mbar_init(e_ready + s, 8); }
        for (int s = 0; s < 2; ++s) { mbar_init(phi_full + s, 1); mbar_init(a_ready + s, 8); }
        fence_barrier_init();
    }
    if (TCSUM) {
        for (uint32_t t = threadIdx.x; t < Cfg::ONES_BYTES / 4; t += P2_THREADS) reinterpret_cast<uint32_t *>(sOnes)[t] = 0x3C003C00u; // fp16 1.0
        fence_proxy_async();
    }
    if (warp == P2_EWARPS) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all(); // the peer's barriers are initialised before anything of ours can arrive on them
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { // SM clock of this launch = cycles / nanoseconds over CTA 0's lifetime
        p.trace[(1 * 64 + 63) * 8 + 0] = clock64();
        p.trace[(2 * 64 + 63) * 8 + 0] = (long long)globaltimer_ns();
    }

    if (warp == P2_EWARPS) { // ---- TMA producer (whole warp runs the loop, one elected lane issues)
        long long pos = u_beg;
        P2Seg sg;
        uint32_t g = 0;
        bool ok = true;
        while (ok && p2_segment(p, pos, u_end, sg)) {
            for (int jt = sg.jb; ok && jt < sg.je; ++jt, ++g) {
                const uint32_t slot = g % P2_STAGES, use = g / P2_STAGES;
                if (!mbar_wait(empty + slot, (use & 1) ^ 1, p.err, 10)) { ok = false; break; }
                if (elect_one()) {
                    uint8_t *st = smem + slot * P2_STAGE;
                    const int j0 = jt * TC_TILE;
                    mbar_arrive_expect_tx(full + slot, (F8LO && p.no_vlo) ? P2_TX - 8192u : P2_TX); // all of the stage's bytes, whoever fetches them
                    const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(p.WB) + (size_t)jt * P2_W_BYTES;
                    if (CL == 1) {
                        tma_load_2d(st, &mapB, 0, j0, full + slot);
                        if (PRECISE) tma_load_2d(st + 16384, &mapB, 64, j0, full + slot); // lo_j
#pragma unroll
                        for (int c = 0; c < (F8LO ? 2 : 4); ++c) // v_hi for both j-halves (+ v_lo in fp16 unless it goes in e5m2)
                            tma_load_2d(st + P2_XB_BYTES + c * P2_VBOX, &mapV, j0 + (c & 1) * 64, (c >> 1) * 64, full + slot);
                        bulk_load_1d(st + P2_XB_BYTES + P2_V_BYTES, wsrc, P2_W_BYTES, full + slot);
                        if (F8LO) {
                            uint8_t *s8 = st + P2_XB_BYTES + P2_V_BYTES + P2_W_BYTES;
                            tma_load_2d(s8, &mapB8, 0, j0, full + slot);                 // y^_j as e5m2, 128 particles x 64 B
                            if (!p.no_vlo) {
                                tma_load_2d(s8 + 8192, &mapV8, j0, 0, full + slot);          // v_lo as e5m2: 64 coordinates x particles [j0, j0 + 64)
                                tma_load_2d(s8 + 8192 + 4096, &mapV8, j0 + 64, 0, full + slot); // ... x particles [j0 + 64, j0 + 128)
                            }
                        }
                    } else if (crank == 0) { // CTA 0: the particle operand and its offset chunk; CTA 1: the four V boxes (about half of the bytes each)
                        tma_load_2d_mc(st, &mapB, 0, j0, full + slot, MC_MASK);
                        if (PRECISE) tma_load_2d_mc(st + 16384, &mapB, 64, j0, full + slot, MC_MASK);
                        bulk_load_1d_mc(st + P2_XB_BYTES + P2_V_BYTES, wsrc, P2_W_BYTES, full + slot, MC_MASK);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            tma_load_2d_mc(st + P2_XB_BYTES + c * P2_VBOX, &mapV, j0 + (c & 1) * 64, (c >> 1) * 64, full + slot, MC_MASK);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp > P2_EWARPS) { // ---- MMA issuer of i-tile wm: warp-uniform control flow, one elected lane issues
        const int wm = warp - P2_EWARPS - 1;
        const uint32_t idesc = make_idesc_f16(TC_TILE, 64);
        const uint32_t st_lo0 = desc_lo_k_sw128(smem_u32(smem));
        const uint32_t aex_lo0 = desc_lo_k_sw128(smem_u32(sAex)) | DESC_LO_K_NOSW_LBO;
        const uint32_t wb_lo0 = desc_lo_k_sw128(smem_u32(smem + P2_XB_BYTES + P2_V_BYTES)) | DESC_LO_K_NOSW_LBO;
        uint32_t g = 0;
        // S_b(gt) = [hi_i | lo_i] . hi_j over the 64 particles of j-half k: 8 TS MMAs (N = 64)
        auto issue_s = [&](int b, uint32_t gt) -> bool {
            const int w = b >> 1, k = b & 1;
            const uint32_t slot = gt % P2_STAGES, use = gt / P2_STAGES;
            if (k == 0 && !mbar_wait(full + slot, use & 1, p.err, 21)) return false;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t dS = tmem + b * 64, aT = tmem + P2_COL_A + w * 64;
                const uint32_t bl = st_lo0 + slot * (P2_STAGE >> 4) + k * (8192 >> 4);
                umma_f16_ts2<false>(dS, aT, bl, idesc);
                umma_f16_ts2<true>(dS, aT + 8, bl + 2, idesc);
                umma_f16_ts2<true>(dS, aT + 16, bl + 4, idesc);
                umma_f16_ts2<true>(dS, aT + 24, bl + 6, idesc);
                if (F8LO) { // (lo_i 2^10) . (y^_j 2^-10) in e5m2: K = 32 per MMA, 64-byte rows (SWIZZLE_64B), j-half k = rows 64 k .. of the tile
                    const uint32_t b8 = st_lo0 + slot * (P2_STAGE >> 4) + ((P2_XB_BYTES + P2_V_BYTES + P2_W_BYTES) >> 4) + k * (4096 >> 4);
                    const uint32_t idesc8 = make_idesc_bf16(TC_TILE, 64);
                    umma_f8_ts2<true>(dS, aT + 32, b8, idesc8);
                    umma_f8_ts2<true>(dS, aT + 40, b8 + 2, idesc8);
                } else {
                    umma_f16_ts2<true>(dS, aT + 32, bl, idesc);
                    umma_f16_ts2<true>(dS, aT + 40, bl + 2, idesc);
                    umma_f16_ts2<true>(dS, aT + 48, bl + 4, idesc);
                    umma_f16_ts2<true>(dS, aT + 56, bl + 6, idesc);
                }
                if (PRECISE) { // hi_i . lo_j
                    const uint32_t bo = bl + (16384 >> 4);
                    umma_f16_ts2<true>(dS, aT, bo, idesc);
                    umma_f16_ts2<true>(dS, aT + 8, bo + 2, idesc);
                    umma_f16_ts2<true>(dS, aT + 16, bo + 4, idesc);
                    umma_f16_ts2<true>(dS, aT + 24, bo + 6, idesc);
                }
                // + u_i + w_j: the 16-column exponent-offset chunks (no-swizzle operands, both from shared memory)
                umma_f16_ss_desc(dS, aex_lo0 + w * (P2_AEX_BYTES >> 4), DESC_HI_K_NOSW,
                                 wb_lo0 + slot * (P2_STAGE >> 4) + k * (2048 >> 4), DESC_HI_K_NOSW, idesc);
                umma_commit(s_full + b);
            }
            __syncwarp();
            return true;
        };
        // Phi_w += E_b(gt) . [v_hi ; v_lo] (64 particles of j-half k): 8 TS MMAs (N = 64)
        auto issue_pv = [&](int b, uint32_t gt, bool first, bool last) -> bool {
            const int w = b >> 1, k = b & 1;
            if (!mbar_wait(e_ready + b, gt & 1, p.err, 22 + b)) return false;
            if (lane == 0 && w == 0) TC_TRACE(0, gt, 3 + k);
            tc_fence_after();
            const uint32_t slot = gt % P2_STAGES;
            if (elect_one()) {
                const uint32_t dP = tmem + P2_COL_PHI + w * 64, e = tmem + b * 64;
                const uint32_t vh = st_lo0 + slot * (P2_STAGE >> 4) + (P2_XB_BYTES >> 4) + k * (P2_VBOX >> 4), vl = vh + 2 * (P2_VBOX >> 4);
                umma_f16_ts2r(dP, e, vh, idesc, (first && k == 0) ? 0u : 1u);
                umma_f16_ts2<true>(dP, e + 8, vh + 2, idesc);
                umma_f16_ts2<true>(dP, e + 32, vh + 4, idesc);
                umma_f16_ts2<true>(dP, e + 40, vh + 6, idesc);
                if (F8LO) { // E (top bytes of the fp16 values: e5m2) . v_lo (e5m2): K = 32 particles per MMA
                    const uint32_t v8 = st_lo0 + slot * (P2_STAGE >> 4) + ((P2_XB_BYTES + P2_V_BYTES + P2_W_BYTES + 8192) >> 4) + k * (4096 >> 4);
                    const uint32_t idesc8 = make_idesc_bf16(TC_TILE, 64);
                    if (!p.no_vlo) {
                        umma_f8_ts2<true>(dP, e + 16, v8, idesc8);
                        umma_f8_ts2<true>(dP, e + 48, v8 + 2, idesc8);
                    }
                } else {
                    umma_f16_ts2<true>(dP, e, vl, idesc);
                    umma_f16_ts2<true>(dP, e + 8, vl + 2, idesc);
                    umma_f16_ts2<true>(dP, e + 32, vl + 4, idesc);
                    umma_f16_ts2<true>(dP, e + 40, vl + 6, idesc);
                }
                if (PRECISE) { // E_lo . v_hi  (E_lo sits in the second 16 columns of each 32-column half)
                    umma_f16_ts2<true>(dP, e + 16, vh, idesc);
                    umma_f16_ts2<true>(dP, e + 24, vh + 2, idesc);
                    umma_f16_ts2<true>(dP, e + 48, vh + 4, idesc);
                    umma_f16_ts2<true>(dP, e + 56, vh + 6, idesc);
                }
                if (TCSUM) { // row sums: E_b . ones (N = 16; every column of the result is the row sum)
                    const uint32_t dR = tmem + P2_COL_A + w * 64 + 48, on = desc_lo_k_sw128(smem_u32(sOnes));
                    const uint32_t idesc16 = make_idesc_f16(TC_TILE, 16);
                    umma_f16_ts2r(dR, e, on, idesc16, (first && k == 0) ? 0u : 1u);
                    umma_f16_ts2<true>(dR, e + 8, on + 2, idesc16);
                    umma_f16_ts2<true>(dR, e + 32, on + 4, idesc16);
                    umma_f16_ts2<true>(dR, e + 40, on + 6, idesc16);
                }
                if (k == 1) {
                    if (CL == 1) umma_commit(empty + slot);
                    else umma_commit_mc(empty + slot, MC_MASK); // both MMA warps of both CTAs must be done with a stage before it is refilled
                }
                if (k == 1 && last) umma_commit(phi_full + w);
            }
            __syncwarp();
            return true;
        };
        long long pos = u_beg;
        P2Seg sg;
        bool ok = true;
        for (uint32_t seg = 0; ok && p2_segment(p, pos, u_end, sg); ++seg) {
            const int nt = sg.je - sg.jb;
            if (!mbar_wait(a_ready + wm, seg & 1, p.err, 20)) { ok = false; break; }
            const int b0 = 2 * wm, b1 = b0 + 1; // this tile's units: j-half 0 and 1
            ok = issue_s(b0, g) && issue_s(b1, g);
            for (int t = 0; ok && t < nt; ++t) {
                const bool more = t + 1 < nt, last = !more;
                if (lane == 0 && wm == 0) TC_TRACE(0, g + t, 0);
                ok = issue_pv(b0, g + t, t == 0, last) && (!more || issue_s(b0, g + t + 1));
                if (lane == 0 && wm == 0) TC_TRACE(0, g + t, 1);
                ok = ok && issue_pv(b1, g + t, t == 0, last) && (!more || issue_s(b1, g + t + 1));
                if (lane == 0 && wm == 0) TC_TRACE(0, g + t, 2);
            }
            g += nt;
        }
    } else { // ---- exp warps: thread = (particle row of i-tile w, 32-column half h of every unit) ---------------------
        const int w = warp >> 3, h = (warp >> 2) & 1;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tP = tmem + P2_COL_PHI + w * 64 + 32 * h + lane_base; // this warp flushes Phi columns [32 h, +32)
        const uint32_t tA = tmem + P2_COL_A + w * 64 + 32 * h + lane_base;   // ... and loads the hi (h = 0) / lo (h = 1) row operand
        const bool tracer = row == 0 && h == 0;
        long long pos = u_beg;
        P2Seg sg;
        uint32_t g = 0;
        bool ok = true;
        for (uint32_t seg = 0; ok && p2_segment(p, pos, u_end, sg); ++seg) {
            const int nt = sg.je - sg.jb;
            const int64_t iw0 = p.row0 + ((int64_t)sg.ip * CL + crank) * (2 * TC_TILE) + w * TC_TILE;
            const int64_t i = iw0 + row;
            { // row operand of particle i -> TMEM: hi half by the h = 0 warp, lo half by the h = 1 warp (previous segment complete: phi_full)
                const uint4 *src = reinterpret_cast<const uint4 *>(p.XA2 + i * P2_A_LD + 64 * h);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (F8LO && h == 1 && k == 1) break; // the e5m2 lo term is 64 bytes = 16 columns
                    uint32_t v[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 x = __ldg(src + 4 * k + q);
                        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                    }
                    tmem_st16(tA + 16 * k, v);
                }
                if (h == 0) { // exponent-offset chunk [u0 u1 u2 1 1 1 0..] -> shared memory, core-matrix order (SS operand)
                    const uint4 *usrc = reinterpret_cast<const uint4 *>(p.UA + i * 16);
                    const uint32_t aex = smem_u32(sAex + w * P2_AEX_BYTES) + p2_ex_offset((uint32_t)row, 0);
                    const uint4 ua0 = __ldg(usrc), ua1 = __ldg(usrc + 1);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(aex), "r"(ua0.x), "r"(ua0.y), "r"(ua0.z), "r"(ua0.w) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(aex + 128u), "r"(ua1.x), "r"(ua1.y), "r"(ua1.z), "r"(ua1.w) : "memory");
                    fence_proxy_async(); // generic-proxy stores -> visible to the MMA's async-proxy reads
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready + w);
            }
            float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f; // partial row sum (this warp's columns) of the rounded E, 4 chains
            // Units of this tile in order (j-tile t, half k); this warp owns columns [32 h, +32) of each.  Its E goes over the
            // first 16 of the 32 S columns it has just read (E_b = S_b columns [0,16) and [32,48)).  The S of the NEXT unit is
            // fetched while the E stores of this one drain, if it is already complete.
            const uint32_t nunits = 2u * (uint32_t)nt;
            // The one unit of the segment (if any) whose 32 columns of this warp, [J + 64 q, +32), meet its 32 rows [r_lo, +32): it holds
            // k(x_i, x_i); dq = this thread's diagonal column there (inside [0, 32) for the rows that have one).
            int q_diag = -1, dq = -1;
            {
                const int64_t J = (int64_t)sg.jb * TC_TILE + 32 * h, r_lo = iw0 + (warp & 3) * 32;
                const int64_t t = r_lo + 31 - J;
                if (t >= 0 && t / 64 < (int64_t)nunits) {
                    const int64_t j0 = J + 64 * (t / 64);
                    if (j0 < r_lo + 32 && j0 + 32 > r_lo) { q_diag = (int)(t / 64); dq = (int)(i - j0); }
                }
            }
            uint32_t r0[32];
            {
                if (tracer) TC_TRACE(1 + w, g, 1);
                if (!mbar_wait(s_full + 2 * w, g & 1, p.err, 40 + 2 * w)) { ok = false; break; }
                tc_fence_after();
                tmem_ld32(tmem + (2 * w) * 64 + 32 * h + lane_base, r0);
            }
            for (uint32_t q = 0; ok && q < nunits; ++q) {
                const int k = (int)(q & 1u);
                const uint32_t gt = g + (q >> 1);
                const int b = 2 * w + k;
                const uint32_t tS = tmem + b * 64 + 32 * h + lane_base;
                const bool has_diag = (int)q == q_diag; // warp-uniform
                tmem_ld_wait();
                if (tracer) TC_TRACE(1 + w, gt, 2 + 3 * k);
                uint32_t pk[16];
                uint32_t pl[PRECISE ? 16 : 1]; // E_lo = fp16(E - E_hi)
                auto exp_chunk = [&](auto diag_tag, auto sum_tag) {
                    constexpr bool DIAG = decltype(diag_tag)::value;
                    constexpr bool ROWSUM = decltype(sum_tag)::value && !TCSUM;
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const float x0 = __uint_as_float(r0[4 * q4]), x1 = __uint_as_float(r0[4 * q4 + 1]);
                        const float x2 = __uint_as_float(r0[4 * q4 + 2]), x3 = __uint_as_float(r0[4 * q4 + 3]);
                        float e0, e1, e2, e3;
                        if (2 * q4 < POLY) { e0 = ex2_poly(x0); e1 = ex2_poly(x1); } else { e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
                        if (2 * q4 + 1 < POLY) { e2 = ex2_poly(x2); e3 = ex2_poly(x3); } else { e2 = ex2_approx(x2); e3 = ex2_approx(x3); }
                        if (DIAG) { // k(x_i, x_i) = exp(0) exactly, like the reference (2^15 after the fp16 scaling)
                            if (dq == 4 * q4) e0 = 32768.0f;
                            if (dq == 4 * q4 + 1) e1 = 32768.0f;
                            if (dq == 4 * q4 + 2) e2 = 32768.0f;
                            if (dq == 4 * q4 + 3) e3 = 32768.0f;
                        }
                        pk[2 * q4] = pack_f16x2(e0, e1);
                        pk[2 * q4 + 1] = pack_f16x2(e2, e3);
                        if constexpr (PRECISE) {
                            // the row sum adds E itself (E_hi + E_lo represents it to 2^-22; the diagonal 2^15 is exact in both)
                            rs0 += e0; rs1 += e1; rs2 += e2; rs3 += e3;
                            pl[2 * q4] = residual_f16x2(pk[2 * q4], e0, e1);
                            pl[2 * q4 + 1] = residual_f16x2(pk[2 * q4 + 1], e2, e3);
                        } else if (ROWSUM) {
                            acc_f16x2(rs0, rs1, pk[2 * q4]);
                            acc_f16x2(rs2, rs3, pk[2 * q4 + 1]);
                        }
                    }
                };
                if (p.dbg != 0) {
#pragma unroll
                    for (int z = 0; z < 16; ++z) pk[z] = r0[z] ^ r0[z + 16];
                    if constexpr (PRECISE) {
#pragma unroll
                        for (int z = 0; z < 16; ++z) pl[z] = r0[z] & r0[z + 16];
                    }
                } else if (has_diag) {
                    exp_chunk(std::true_type{}, std::true_type{});
                } else {
                    exp_chunk(std::false_type{}, std::true_type{});
                }
                tmem_st16(tS, pk);
                if constexpr (PRECISE) tmem_st16(tS + 16, pl);
                if (F8LO && !p.no_vlo) { // e5m2 copy of E for the E . v_lo term: the top byte of each fp16 value (one PRMT per four values)
                    uint32_t e8[8];
#pragma unroll
                    for (int z = 0; z < 8; ++z) e8[z] = __byte_perm(pk[2 * z], pk[2 * z + 1], 0x7531);
                    tmem_st8(tS + 16, e8);
                }
                // E_b complete -> the tile's MMA warp first (PV(b), then the next S into this buffer: that chain is what this warp will wait for
                // two units from now), then the next unit's S.  (Polling for the next S before the arrival -- to overlap its TMEM load with the
                // store drain -- found it incomplete four times out of five and only delayed the hand-off by the poll.)
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(e_ready + b);
                if (tracer) TC_TRACE(1 + w, gt, 3 + 3 * k);
                if (q + 1 < nunits) {
                    const int kn = (int)((q + 1) & 1u);
                    const uint32_t gn = g + ((q + 1) >> 1);
                    if (tracer) TC_TRACE(1 + w, gn, 1 + 3 * kn);
                    if (!mbar_wait(s_full + 2 * w + kn, gn & 1, p.err, 40 + 2 * w + kn)) { ok = false; break; }
                    tc_fence_after();
                    tmem_ld32(tmem + (2 * w + kn) * 64 + 32 * h + lane_base, r0);
                }
            }
            g += nt;
            if (!ok || !mbar_wait(phi_full + w, seg & 1, p.err, 50)) { ok = false; break; }
            tc_fence_after();
            { // ---- flush this warp's half of Phi_w and its partial row sum: TMEM -> global partial sums
                const bool valid = i < p.row0 + p.n_rows;
                float *dst = p.phi_buf + i * TC_PHI_LD + 32 * h;
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tP + c0, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int z = 0; z < 16; ++z) atomicAdd(dst + c0 + z, __uint_as_float(v[z]) * TC_E_UNSCALE);
                    }
                }
                if (TCSUM) {
                    if (h == 0) {
                        uint32_t v[16];
                        tmem_ld16(tmem + P2_COL_A + w * 64 + 48 + lane_base, v);
                        tmem_ld_wait();
                        if (valid) atomicAdd(p.phi_buf + i * TC_PHI_LD + TC_ONES_ROW, __uint_as_float(v[0]) * TC_E_UNSCALE);
                    }
                } else if (valid) {
                    atomicAdd(p.phi_buf + i * TC_PHI_LD + TC_ONES_ROW, ((rs0 + rs1) + (rs2 + rs3)) * TC_E_UNSCALE);
                }
                tc_fence_before(); // the a_ready arrival of the next segment orders these loads before its first MMA
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all(); // nobody leaves while the peer may still multicast into its shared memory or arrive on its barriers
    if (warp == P2_EWARPS) tmem_dealloc(tmem, 512);
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        p.trace[(1 * 64 + 63) * 8 + 7] = clock64();
        p.trace[(2 * 64 + 63) * 8 + 7] = (long long)globaltimer_ns();
    }
}

} // namespace tc
} // namespace svgdb
