// kernels_dist_tc.cuh — persistent tensor-core distance pass for the exact median bandwidth (SVGDB_PRECISION_TC32).
//
// Same counting / collecting contract as dist_pass_f64_kernel (kernels_f64.cuh) and the same arithmetic as the first
// tensor-core version (dist_tc32_kernel): centred particles split in two bf16 terms, D2 = |x~_i|^2 + |x~_j|^2 - 2 x~_i.x~_j
// from hi.hi + lo.hi + hi.lo products with fp32 accumulation, the norms as three-term bf16 splits in a 16-column K chunk,
// keys = IEEE bits of (double)D2.  Reference semantics: Kernel/GaussianRBFKernel.hpp:168-188, 222-254.
//
// What changed is the mapping onto the SM (the lessons of kernels_phi_tc.cuh):
//   * the row operand -2 [hi | lo] of both i-tiles lives in TMEM (written once per segment): every big MMA is TS mode and
//     reads only the column operand from shared memory (an SS MMA at M = N = 128 needs the SM's whole 128 B/clk);
//   * units of 128 x 64 distances with THREE accumulator buffers per i-tile (TMEM [192 w + 64 (c % 3), +64)): the MMA warp
//     of a tile runs up to two units ahead of its counting warpgroup, nobody waits on a round trip;
//   * one MMA-issuing warp per i-tile; persistent CTAs walk contiguous ranges of the (i-pair, j-tile) upper triangle;
//   * the pass is bound by the counting epilogue's instruction issue (~6 instructions per distance, ncu: 48 % issue
//     utilisation with two counting warps per scheduler), so SIXTEEN counting warps (four per scheduler) share a unit:
//     warp = (i-tile, row quadrant, 32-column half).
#pragma once
#include "kernels_phi_tc.cuh"

namespace svgdb {
namespace tc {

constexpr int D2_STAGES = 3;
constexpr int D2_CWARPS = 16;                         // counting warps: 8 per i-tile = 4 row quadrants x 2 column halves of a unit
constexpr int D2_THREADS = (D2_CWARPS + 3) * 32;      // + TMA producer (warp 16) + one MMA issuer per i-tile (warps 17, 18)
constexpr int D2_PRIV = 36;                           // per-thread staging slots (a 32-column chunk may add 32)
constexpr uint32_t D2_XB_BYTES = 32768;               // 128 particles x [hi | lo] bf16 (two 128 B-row SWIZZLE_128B boxes)
constexpr uint32_t D2_STAGE = D2_XB_BYTES + P2_W_BYTES; // 36 KB
constexpr uint32_t D2_TX = D2_STAGE;
constexpr uint32_t D2_SMEM = D2_STAGES * D2_STAGE + 2 * P2_AEX_BYTES + D2_CWARPS * TC_WBUF * 4 + D2_CWARPS * 32 * D2_PRIV * 4 + 256 + 1024;
constexpr uint32_t D2_COL_A = 384;
static_assert(HIST_BINS * 8 <= D2_CWARPS * TC_WBUF * 4, "histogram must fit the warp staging area");

// Operand rows for the distance pass, one warp per particle (scale 1):
//   XA2[row] = -2 [hi | lo] (row operand -> TMEM),  XBD[row] = [hi | lo] (column operand, TMA),
//   UA[row] = [r0 r1 r2 1 1 1 0..],  WB (core-matrix order) = [1 1 1 r0 r1 r2 1 1 1 0..],  r = |x~|^2 as a 3-term bf16 split
//   (columns 6..8 of the row chunk are filled by a FOLD pass with the split of -lo, see dist2_tc32_kernel);
//   padding rows carry r = +inf: their distances never count.  rt[row] = r (FP64) for the pair kernel's later use.
// Rows [row_begin, n_rows_a) of the row-side arrays (and of the column-side ones below n_rows_b).
__global__ void split_dist2_kernel(const double *__restrict__ X, const double *__restrict__ colsum, int64_t n, int64_t row_begin, int64_t n_rows_a,
                                   int64_t n_rows_b, int d, __nv_bfloat16 *__restrict__ XA2, __nv_bfloat16 *__restrict__ XBD,
                                   __nv_bfloat16 *__restrict__ UA, __nv_bfloat16 *__restrict__ WB, double *__restrict__ rt)
{
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows_a) return;
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        double xc = 0.0;
        if (row < n && k < d) xc = X[row * d + k] - colsum[k] / (double)n;
        s += xc * xc;
        const __nv_bfloat16 hi = __float2bfloat16_rn((float)xc);
        const __nv_bfloat16 lo = __float2bfloat16_rn((float)(xc - (double)__bfloat162float(hi)));
        XA2[row * P2_A_LD + k] = __float2bfloat16_rn(-2.0f * __bfloat162float(hi)); // exact: a power-of-two multiple
        XA2[row * P2_A_LD + 64 + k] = __float2bfloat16_rn(-2.0f * __bfloat162float(lo));
        if (row < n_rows_b) {
            XBD[row * 128 + k] = hi;
            XBD[row * 128 + 64 + k] = lo;
        }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && row < n_rows_b) rt[row] = s;
    if (lane < 16) {
        __nv_bfloat16 r0, r1, r2;
        split3_bf16((row < n) ? s : (double)INFINITY, r0, r1, r2);
        const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
        UA[row * 16 + lane] = lane == 0 ? r0 : lane == 1 ? r1 : lane == 2 ? r2 : lane < 6 ? one : zero;
        if (row < n_rows_b)
            *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<uint8_t *>(WB) + (row >> 7) * P2_W_BYTES + p2_ex_offset((uint32_t)(row & 127), (uint32_t)lane)) =
                lane < 3 ? one : lane == 3 ? r0 : lane == 4 ? r1 : lane == 5 ? r2 : lane < 9 ? one : zero; // 6..8: multiply the folded -lo of the row chunk
    }
}

// The operands of the F16 variant of a folded pass (dist2_tc32_kernel<..., F16 = true>): y = s (x - mean) with s a power of two
// chosen by the host so that the median of |y_i - y_j|^2 is ~512 (fp16 has the dynamic range of the particle cloud, not more),
//   XA2[row] = -2 [hi | lo] fp16 (row particle: two terms, 22 bits),   XB[row] = hi fp16 (column particle: its fp16 image),
//   UA[row] = [r r r 1 1 1 0..] with r = |hi + lo|^2,  WB = [1 1 1 r^ r^ r^ 1 1 1 0..] with r^ = |hi|^2 (three-term bf16 splits):
// the accumulator is the exact squared distance between particle i and the fp16 image of particle j, times s^2.
__global__ void split_dist2h_kernel(const double *__restrict__ X, const double *__restrict__ colsum, int64_t n, int64_t row_begin, int64_t n_rows_a,
                                    int64_t n_rows_b, int d, double s, __half *__restrict__ XA2, __half *__restrict__ XB,
                                    __nv_bfloat16 *__restrict__ UA, __nv_bfloat16 *__restrict__ WB, double *__restrict__ rt)
{
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows_a) return;
    double r_full = 0.0, r_hi = 0.0, r_x = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        double xc = 0.0;
        if (row < n && k < d) xc = X[row * d + k] - colsum[k] / (double)n;
        r_x += xc * xc;
        const double y = s * xc;
        const __half hi = __double2half(y);
        const double hid = (double)__half2float(hi);
        const __half lo = __double2half(y - hid);
        const double full = hid + (double)__half2float(lo);
        r_full += full * full;
        r_hi += hid * hid;
        XA2[row * P2_A_LD + k] = __float2half_rn(-2.0f * __half2float(hi)); // exact: a power-of-two multiple
        XA2[row * P2_A_LD + 64 + k] = __float2half_rn(-2.0f * __half2float(lo));
        if (row < n_rows_b) XB[row * 64 + k] = hi;
    }
    for (int o = 16; o; o >>= 1) {
        r_full += __shfl_xor_sync(0xffffffffu, r_full, o);
        r_hi += __shfl_xor_sync(0xffffffffu, r_hi, o);
        r_x += __shfl_xor_sync(0xffffffffu, r_x, o);
    }
    if (lane == 0 && row < n_rows_b) rt[row] = r_x;
    if (lane < 16) {
        __nv_bfloat16 a0, a1, a2, b0, b1, b2;
        split3_bf16((row < n) ? r_full : (double)INFINITY, a0, a1, a2);
        split3_bf16((row < n) ? r_hi : (double)INFINITY, b0, b1, b2);
        const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
        UA[row * 16 + lane] = lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : lane < 6 ? one : zero;
        if (row < n_rows_b)
            *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<uint8_t *>(WB) + (row >> 7) * P2_W_BYTES + p2_ex_offset((uint32_t)(row & 127), (uint32_t)lane)) =
                lane < 3 ? one : lane == 3 ? b0 : lane == 4 ? b1 : lane == 5 ? b2 : lane < 9 ? one : zero;
    }
}

// warp-collective: move `count` staged distances to the global candidate list as keys
__device__ __noinline__ void dist_flush2(const float *mybuf, unsigned int count, unsigned long long *cand, unsigned long long *cand_count,
                                         unsigned long long capacity)
{
    __syncwarp();
    const unsigned int lane = threadIdx.x & 31;
    unsigned long long base = 0ull;
    if (lane == 0 && count) base = atomicAdd(cand_count, (unsigned long long)count);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned int q = lane; q < count; q += 32)
        if (base + q < capacity) cand[base + q] = dist_key(mybuf[q]);
    __syncwarp();
}
// staging bypass for a chunk that overflows the warp buffer (very wide bracket): reserve straight in the global list
__device__ __noinline__ void dist_append_global2(float d2, unsigned int wgt, unsigned long long *cand, unsigned long long *cand_count,
                                                 unsigned long long capacity)
{
    const unsigned long long key = dist_key(d2);
    const unsigned long long g = atomicAdd(cand_count, (unsigned long long)wgt);
    if (g < capacity) cand[g] = key;
    if (wgt == 2u && g + 1 < capacity) cand[g + 1] = key;
}

struct Dist2Args {
    const __nv_bfloat16 *XA2; // [n_pad128 + 256][128]
    const __nv_bfloat16 *UA;  // [n_pad128 + 256][16]
    const __nv_bfloat16 *WB;  // [n_pad128 / 128][4 KB]
    int64_t n_total;
    int n_jtiles;
    int jt_begin;                           // column tiles below this one are left out (they belong to an earlier launch of the same pass)
    int pair_offset, pair_stride, n_ipairs; // this rank owns i-pairs offset, offset + stride, ... (n_ipairs of them)
    float lo_f, hi_f;
    // FOLD instantiations: -lo_f rides in the norm K chunk as a three-term bf16 split (exact: 24 bits), so the accumulator IS
    // t = d2 - lo_f and the epilogue needs no subtraction; fold_l01 = bf16 pair (L0, L1), fold_l2 = (L2, 0)
    unsigned int fold_l01, fold_l2;
    float out_scale;         // FOLD: collected values are (t + lo_f) * out_scale (1 / s^2 of the F16 variant, else 1)
    unsigned int width_bits; // IEEE bits of a float > fl(hi_f - lo_f): d2 is collected iff 0 <= fl(d2 - lo_f) < width
    int open_low;            // lo_f = -inf (nothing lies below): collect d2 < hi_f by a plain compare
    unsigned long long lo_key;
    int shift;
    unsigned long long *below, *hist, *cand, *cand_count;
    unsigned long long capacity;
    int *err;
    int dbg; // measurement aid: 1 = counting warps only hand the buffers back, 2 = count without collecting, 3 = 1 + no TMA after priming
};

// Work list: for this rank's l-th i-pair ip = offset + stride * l, the j-tiles jt in [max(2 ip, jt_begin), n_jtiles) (tile-level upper
// triangle; the two diagonal-adjacent tiles are sorted out by the per-tile weights).  Linearised in that order.
struct D2Seg { int ip, jb, je; };
struct D2Cursor { int l; long long base; }; // base = linear position of the first unit of i-pair l
__device__ __forceinline__ bool d2_segment(const Dist2Args &p, D2Cursor &cur, long long &pos, long long end, D2Seg &s)
{
    if (pos >= end) return false;
    for (;;) { // advance to the i-pair containing pos
        // the rank's i-pairs are visited in folded order 0, n-1, 1, n-2, ...: a long row of the triangle is followed by a short one,
        // so every CTA's contiguous range holds about the same number of segments (each segment start costs a pipeline refill)
        const int lf = (cur.l & 1) ? p.n_ipairs - 1 - (cur.l >> 1) : (cur.l >> 1);
        const int ip = p.pair_offset + p.pair_stride * lf;
        const int j0 = max(2 * ip, p.jt_begin);
        const long long len = max(0, p.n_jtiles - j0);
        if (pos < cur.base + len) {
            s.ip = ip;
            s.jb = j0 + (int)(pos - cur.base);
            const long long seg_end = min(end, cur.base + len);
            s.je = s.jb + (int)(seg_end - pos);
            pos = seg_end;
            return true;
        }
        cur.base += len;
        ++cur.l;
        if (cur.l >= p.n_ipairs) return false;
    }
}
__device__ __forceinline__ long long d2_total_units(const Dist2Args &p)
{
    long long tot = 0;
    for (int l = 0; l < p.n_ipairs; ++l) tot += max(0, p.n_jtiles - max(2 * (p.pair_offset + p.pair_stride * l), p.jt_begin));
    return tot;
}

// MODE_HIST, GATED = true selects 32-bit shared counters (native ATOMS.ADD; the 64-bit shared add is a compare-and-swap loop,
// 4x slower under the contention of a pass whose distances share a few bins): valid while a CTA sees fewer than 2^32 weighted
// pairs, which the launcher checks.
// GATED: the bracket is expected to hold so few distances that most 32 x 32 chunks contain none -- count and test with 3
// instructions per distance and take the collecting path only for chunks where some lane saw a hit.
// FOLD (collecting passes over a predicted bracket): the accumulator is t = d2 - lo (see Dist2Args); counting is the sign bit of t,
// the bracket test is bits(t) <u bits(width).  GATED + FOLD: 1.5 instructions per distance (LEA.HI per value, one three-input
// FMNMX3 over |t| per two values; a chunk whose smallest |t| is below the width takes the exact collecting path).  The values of
// a pass are the roundings of ITS OWN accumulation, so passes that must agree with each other on every single distance
// (histogram narrowing followed by a collecting pass) all run unfolded.
// F16 (folded collecting passes only): two products on fp16 operands instead of three on bf16 -- the row particle keeps two terms, the
// column particle is its fp16 image (split_dist2h_kernel): 8 + 1 MMAs per unit instead of 12 + 1.  The values of such a pass are
// exact distances to the rounded column particles: a zero-mean perturbation of ~1e-5 relative per pair, which moves the median of
// n^2 of them by far less than the 1e-5 tolerance of the scale.  Lower bracket end, width and the folded -lo are given in the
// scaled units of the operands (powers of two: every comparison is unchanged); candidates are scaled back (out_scale).
template <int MODE, bool GATED, bool FOLD, bool F16 = false>
__global__ void __launch_bounds__(D2_THREADS, 1)
dist2_tc32_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ Dist2Args p)
{
    const long long units = d2_total_units(p);
    const long long u_beg = units * blockIdx.x / gridDim.x, u_end = units * (blockIdx.x + 1) / gridDim.x;
    if (u_beg >= u_end) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sAex = smem + D2_STAGES * D2_STAGE;      // [2][P2_AEX_BYTES]
    float *wbuf = (float *)(sAex + 2 * P2_AEX_BYTES); // [D2_CWARPS][TC_WBUF]      per-warp staging
    float *priv = wbuf + D2_CWARPS * TC_WBUF;         // [D2_CWARPS][D2_PRIV][32] per-thread staging, lane-interleaved
    uint64_t *bars = (uint64_t *)(priv + D2_CWARPS * 32 * D2_PRIV);
    uint64_t *full = bars;                  // [D2_STAGES]
    uint64_t *empty = full + D2_STAGES;     // [D2_STAGES] both tiles' MMAs reading the stage complete (2 commits)
    uint64_t *s_full = empty + D2_STAGES;   // [2][3] accumulator buffer complete (commit)
    uint64_t *s_free = s_full + 6;          // [2][3] accumulator buffer is in the counting warps' registers (8 warp arrivals)
    uint64_t *a_ready = s_free + 6;         // [2] row operand of tile w in TMEM / shared memory (8 warp arrivals)
    uint64_t *seg_done = a_ready + 2;       // [2] every MMA of the segment on tile w complete (commit)
    uint32_t *tmem_holder = (uint32_t *)(seg_done + 2);
    unsigned long long *shist = (unsigned long long *)wbuf; // [HIST_BINS] 64-bit (a CTA sees > 2^32 pairs at N = 1M); MODE_HIST only: aliases the warp staging buffers
    unsigned int *shist32 = (unsigned int *)wbuf;           // [HIST_BINS] the 32-bit variant
    constexpr bool HIST32 = MODE == MODE_HIST && GATED;
    static_assert(!(FOLD && MODE == MODE_HIST), "histogram passes run unfolded");
    static_assert(!F16 || FOLD, "the fp16 two-product operands serve folded collecting passes only");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < D2_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 2); }
        for (int s = 0; s < 6; ++s) { mbar_init(s_full + s, 1); mbar_init(s_free + s, 8); }
        for (int s = 0; s < 2; ++s) { mbar_init(a_ready + s, 8); mbar_init(seg_done + s, 1); }
        fence_barrier_init();
    }
    if (MODE == MODE_HIST)
        for (int b = threadIdx.x; b < HIST_BINS; b += blockDim.x) {
            if (HIST32) shist32[b] = 0u;
            else shist[b] = 0ull;
        }
    if (warp == D2_CWARPS) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == D2_CWARPS) { // ---- TMA producer
        long long pos = u_beg;
        D2Cursor cur{0, 0};
        D2Seg sg;
        uint32_t g = 0;
        bool ok = true;
        while (ok && d2_segment(p, cur, pos, u_end, sg)) {
            for (int jt = sg.jb; ok && jt < sg.je; ++jt, ++g) {
                const uint32_t slot = g % D2_STAGES, use = g / D2_STAGES;
                if (!mbar_wait(empty + slot, (use & 1) ^ 1, p.err, 50)) { ok = false; break; }
                if (p.dbg == 3 && g >= D2_STAGES) { // measurement aid: stale operands, no TMA / L2 traffic
                    if (elect_one()) mbar_arrive(full + slot);
                    __syncwarp();
                    continue;
                }
                if (elect_one()) {
                    uint8_t *st = smem + slot * D2_STAGE;
                    mbar_arrive_expect_tx(full + slot, F16 ? 16384u + P2_W_BYTES : D2_TX);
                    tma_load_2d(st, &mapB, 0, jt * TC_TILE, full + slot);
                    if (!F16) tma_load_2d(st + 16384, &mapB, 64, jt * TC_TILE, full + slot);
                    bulk_load_1d(st + D2_XB_BYTES, reinterpret_cast<const uint8_t *>(p.WB) + (size_t)jt * P2_W_BYTES, P2_W_BYTES, full + slot);
                }
                __syncwarp();
            }
        }
    } else if (warp > D2_CWARPS) { // ---- MMA issuer of i-tile w
        const int w = warp - D2_CWARPS - 1;
        const uint32_t idesc_c = make_idesc_bf16(TC_TILE, 64);                      // norm chunk: always bf16
        const uint32_t idesc = F16 ? make_idesc_f16(TC_TILE, 64) : idesc_c;          // products
        const uint32_t st_lo0 = desc_lo_k_sw128(smem_u32(smem));
        const uint32_t aex_lo = (desc_lo_k_sw128(smem_u32(sAex + w * P2_AEX_BYTES))) | DESC_LO_K_NOSW_LBO;
        const uint32_t wb_lo0 = desc_lo_k_sw128(smem_u32(smem + D2_XB_BYTES)) | DESC_LO_K_NOSW_LBO;
        const uint32_t aT = tmem + D2_COL_A + w * 64;
        long long pos = u_beg;
        D2Cursor cur{0, 0};
        D2Seg sg;
        uint32_t g = 0, c = 0; // j-tiles / units issued so far
        bool ok = true;
        for (uint32_t seg = 0; ok && d2_segment(p, cur, pos, u_end, sg); ++seg) {
            if (!mbar_wait(a_ready + w, seg & 1, p.err, 60)) { ok = false; break; }
            for (int jt = sg.jb; ok && jt < sg.je; ++jt, ++g) {
                const uint32_t slot = g % D2_STAGES, use = g / D2_STAGES;
                if (!mbar_wait(full + slot, use & 1, p.err, 61)) { ok = false; break; }
#pragma unroll
                for (int k = 0; k < 2; ++k, ++c) {
                    const uint32_t buf = c % 3u, bu = c / 3u;
                    if (!mbar_wait(s_free + 3 * w + buf, (bu & 1) ^ 1, p.err, 62 + w)) { ok = false; break; }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t dS = tmem + w * 192 + buf * 64;
                        const uint32_t bh = st_lo0 + slot * (D2_STAGE >> 4) + k * (8192 >> 4), bl = bh + (16384 >> 4);
                        umma_f16_ts2<false>(dS, aT, bh, idesc);          // (-2 hi_i) . hi_j
                        umma_f16_ts2<true>(dS, aT + 8, bh + 2, idesc);
                        umma_f16_ts2<true>(dS, aT + 16, bh + 4, idesc);
                        umma_f16_ts2<true>(dS, aT + 24, bh + 6, idesc);
                        umma_f16_ts2<true>(dS, aT + 32, bh, idesc);      // (-2 lo_i) . hi_j
                        umma_f16_ts2<true>(dS, aT + 40, bh + 2, idesc);
                        umma_f16_ts2<true>(dS, aT + 48, bh + 4, idesc);
                        umma_f16_ts2<true>(dS, aT + 56, bh + 6, idesc);
                        if (!F16) {
                            umma_f16_ts2<true>(dS, aT, bl, idesc);       // (-2 hi_i) . lo_j
                            umma_f16_ts2<true>(dS, aT + 8, bl + 2, idesc);
                            umma_f16_ts2<true>(dS, aT + 16, bl + 4, idesc);
                            umma_f16_ts2<true>(dS, aT + 24, bl + 6, idesc);
                        }
                        umma_f16_ss_desc(dS, aex_lo, DESC_HI_K_NOSW, wb_lo0 + slot * (D2_STAGE >> 4) + k * (2048 >> 4), DESC_HI_K_NOSW, idesc_c); // + r_i + r_j
                        umma_commit(s_full + 3 * w + buf);
                        if (k == 1) umma_commit(empty + slot);
                        if (k == 1 && jt + 1 == sg.je) umma_commit(seg_done + w);
                    }
                    __syncwarp();
                }
            }
        }
    } else { // ---- counting warps: thread = row i of tile w, 32-column half h of every unit
        const int w = warp >> 3, h = (warp >> 2) & 1;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tA = tmem + D2_COL_A + w * 64 + lane_base;
        float *mybuf = wbuf + warp * TC_WBUF;
        const uint32_t priv_base = smem_u32(priv + warp * (D2_PRIV * 32) + lane); // this thread's lane-interleaved slots
        const uint32_t wbuf_base = smem_u32(mybuf);
        unsigned int count = 0; // warp-uniform fill level of mybuf
        unsigned long long below = 0ull;
        uint32_t paddr = priv_base; // next free private slot (slot e of this thread lives at priv_base + 128 e)
        unsigned int cur_wgt = 1u;  // weight of the entries currently staged in the private slots
        // warp-collective: move the private entries (all of weight cur_wgt) into the warp buffer / histogram
        auto compact = [&]() {
            const uint32_t mine = (paddr - priv_base) >> 7;
            const unsigned int tot = __reduce_add_sync(0xffffffffu, mine);
            if (tot) { // warp-uniform
                if (MODE == MODE_HIST) {
                    // consecutive entries of a thread mostly share a bin while the range is still wide: one add per run
                    unsigned int run_bin = 0xffffffffu, run_cnt = 0u;
                    auto flush_run = [&]() {
                        if (run_cnt) {
                            if (HIST32) atomicAdd(&shist32[run_bin], run_cnt);
                            else atomicAdd(&shist[run_bin], (unsigned long long)run_cnt);
                        }
                    };
                    for (uint32_t e = 0; e < mine; ++e) {
                        const unsigned long long bin = (dist_key(lds_f32(priv_base + 128u * e)) - p.lo_key) >> p.shift;
                        if (bin < (unsigned long long)HIST_BINS) { // the collected range may end just past hi
                            if ((unsigned int)bin == run_bin) run_cnt += cur_wgt;
                            else { flush_run(); run_bin = (unsigned int)bin; run_cnt = cur_wgt; }
                        }
                    }
                    flush_run();
                } else {
                    if (count + tot * cur_wgt > (unsigned int)TC_WBUF) { dist_flush2(mybuf, count, p.cand, p.cand_count, p.capacity); count = 0; }
                    if (tot * cur_wgt > (unsigned int)TC_WBUF) { // more than an empty buffer holds: straight to global
                        for (uint32_t e = 0; e < mine; ++e)
                            dist_append_global2(FOLD ? (lds_f32(priv_base + 128u * e) + p.lo_f) * p.out_scale : lds_f32(priv_base + 128u * e), cur_wgt, p.cand,
                                                p.cand_count, p.capacity);
                    } else {
                        uint32_t incl = mine; // inclusive scan over lanes
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        uint32_t dst = wbuf_base + 4u * (count + (incl - mine) * cur_wgt);
                        for (uint32_t e = 0; e < mine; ++e) {
                            const float v = FOLD ? (lds_f32(priv_base + 128u * e) + p.lo_f) * p.out_scale : lds_f32(priv_base + 128u * e); // staged as t = d2 - lo when folded
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
        long long pos = u_beg;
        D2Cursor cur{0, 0};
        D2Seg sg;
        uint32_t c = 0;
        bool ok = true;
        for (uint32_t seg = 0; ok && d2_segment(p, cur, pos, u_end, sg); ++seg) {
            const int itile = 2 * sg.ip + w;
            const int64_t iw0 = (int64_t)itile * TC_TILE;
            const int64_t i = iw0 + row;
            const bool row_valid = i < p.n_total;
            // rows outside the particle set never count: give them an empty bracket with nothing below it
            const float lo = row_valid ? p.lo_f : -INFINITY, hi = row_valid ? p.hi_f : -INFINITY;
            const unsigned int wbits = row_valid ? p.width_bits : 0u;       // IEEE bits of (hi - lo) rounded up; 0: never in the bracket
            const bool open_low = p.open_low != 0; // lo = -inf: the difference trick does not apply
            if (seg > 0) { // the previous segment's MMAs on this tile still read the row operand
                if (!mbar_wait(seg_done + w, (seg - 1) & 1, p.err, 70 + w)) { ok = false; break; }
                tc_fence_after();
            }
            { // row operand -2 [hi | lo] -> TMEM, norm chunk -> shared memory (core-matrix order): the h = 0 warps of the tile
                if (h == 0) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(p.XA2 + i * P2_A_LD);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t v[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 x = __ldg(src + 4 * k + q);
                            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                        }
                        tmem_st16(tA + 16 * k, v);
                    }
                    const uint4 *usrc = reinterpret_cast<const uint4 *>(p.UA + i * 16);
                    const uint32_t aex = smem_u32(sAex + w * P2_AEX_BYTES) + p2_ex_offset((uint32_t)row, 0);
                    uint4 ua0 = __ldg(usrc), ua1 = __ldg(usrc + 1);
                    if (FOLD) { ua0.w = p.fold_l01; ua1.x = p.fold_l2; } // columns 6, 7, 8 = -lo (the column chunk carries ones there)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(aex), "r"(ua0.x), "r"(ua0.y), "r"(ua0.z), "r"(ua0.w) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(aex + 128u), "r"(ua1.x), "r"(ua1.y), "r"(ua1.z), "r"(ua1.w) : "memory");
                    fence_proxy_async();
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready + w);
            }
            for (int jt = sg.jb; ok && jt < sg.je; ++jt) {
                // symmetric enumeration: tiles below the diagonal are covered by their transposes (weight 2)
                const unsigned int wgt = jt < itile ? 0u : (jt == itile ? 1u : 2u);
                // (not unrolled, rare paths marked unlikely: the loop body with its in-line compactions is ~15 KB of code per copy, and the
                // counting warps lost a sixth of their time to instruction fetch -- stall_no_inst + branch_resolving, profiles/r02)
#pragma unroll 1
                for (int k = 0; k < 2; ++k, ++c) {
                    const uint32_t buf = c % 3u, bu = c / 3u;
                    const uint32_t tS = tmem + w * 192 + buf * 64 + 32 * h + lane_base;
                    const int64_t j0 = (int64_t)jt * TC_TILE + 64 * k + 32 * h;
                    const int dcol = (int)(i - j0);
                    const bool has_diag = jt == itile;
                    if (!mbar_wait(s_full + 3 * w + buf, bu & 1, p.err, 72 + w)) { ok = false; break; }
                    tc_fence_after();
                    if (wgt == 0u) { // nothing to count: hand the buffer straight back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_free + 3 * w + buf);
                        continue;
                    }
                    uint32_t r0[32];
                    tmem_ld32(tS, r0);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free + 3 * w + buf); // this warp's part of the accumulator is in registers
                    if (p.dbg == 1 || p.dbg == 3) continue;
                    if (__builtin_expect(wgt != cur_wgt, 0)) { compact(); cur_wgt = wgt; }
                    if (__builtin_expect(has_diag, 0)) { // |x_i - x_i|^2 = 0 exactly (diagonal tiles only); folded: t = 0 - lo
                        const uint32_t dz = FOLD ? __float_as_uint(-p.lo_f) : 0u;
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (dcol == q) r0[q] = dz;
                    }
                    // Five instructions per distance, no branch: t = d2 - lo; count += sign bit of t (d2 < lo exactly: a negative
                    // difference never rounds to +0); in = bits(t) < bits(width) as UNSIGNED integers (0 <= t < width; negative
                    // t has the top bit set); @in: store d2 into this thread's staging column and bump its pointer.
                    // width is rounded up from hi - lo, so the collected set is the contiguous range [lo, hi') with
                    // hi' >= hi: a superset of the bracket, which the selection on the host side allows for.
                    unsigned int cnt4[4] = {0u, 0u, 0u, 0u}; // four independent counters: no long dependent add chain
                    auto visit = [&](float d2, unsigned int &cnt) {
                        asm volatile("{\n\t.reg .pred pi;\n\t.reg .f32 t;\n\t.reg .b32 u, s;\n\t"
                                     "sub.rn.f32 t, %2, %3;\n\t"
                                     "mov.b32 u, t;\n\t"
                                     "shr.u32 s, u, 31;\n\t"
                                     "add.u32 %0, %0, s;\n\t"
                                     "setp.lt.u32 pi, u, %4;\n\t"
                                     "@pi st.shared.f32 [%1], %2;\n\t"
                                     "@pi add.u32 %1, %1, 128;\n\t}"
                                     : "+r"(cnt), "+r"(paddr)
                                     : "f"(d2), "f"(lo), "r"(wbits)
                                     : "memory");
                    };
                    if (p.dbg == 2) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) cnt4[q & 3] += FOLD ? r0[q] >> 31 : __float_as_uint(__uint_as_float(r0[q]) - lo) >> 31;
                    } else if (FOLD && MODE == MODE_COLLECT) {
                        // the accumulator is t = d2 - lo: count += sign bit (LEA.HI); in the bracket iff bits(t) <u bits(width)
                        if (GATED) {
                            // ... which needs |t| < width: one FMNMX3 over |t| per two distances decides whether the chunk holds a candidate
                            float m0 = INFINITY, m1 = INFINITY;
#pragma unroll
                            for (int q = 0; q < 32; q += 4) {
                                cnt4[0] += r0[q] >> 31;
                                cnt4[1] += r0[q + 1] >> 31;
                                cnt4[2] += r0[q + 2] >> 31;
                                cnt4[3] += r0[q + 3] >> 31;
                                m0 = fminf(fminf(m0, fabsf(__uint_as_float(r0[q]))), fabsf(__uint_as_float(r0[q + 1])));
                                m1 = fminf(fminf(m1, fabsf(__uint_as_float(r0[q + 2]))), fabsf(__uint_as_float(r0[q + 3])));
                            }
                            if (__builtin_expect(__any_sync(0xffffffffu, fminf(m0, m1) < __uint_as_float(wbits)), 0)) { // rare: collect with the exact test
#pragma unroll
                                for (int q = 0; q < 32; ++q)
                                    asm volatile("{\n\t.reg .pred pi;\n\t"
                                                 "setp.lt.u32 pi, %1, %2;\n\t"
                                                 "@pi st.shared.b32 [%0], %1;\n\t"
                                                 "@pi add.u32 %0, %0, 128;\n\t}"
                                                 : "+r"(paddr)
                                                 : "r"(r0[q]), "r"(wbits)
                                                 : "memory");
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q) {
                                cnt4[q & 3] += r0[q] >> 31;
                                asm volatile("{\n\t.reg .pred pi;\n\t"
                                             "setp.lt.u32 pi, %1, %2;\n\t"
                                             "@pi st.shared.b32 [%0], %1;\n\t"
                                             "@pi add.u32 %0, %0, 128;\n\t}"
                                             : "+r"(paddr)
                                             : "r"(r0[q]), "r"(wbits)
                                             : "memory");
                            }
                        }
                    } else if (GATED && MODE == MODE_COLLECT && !open_low) {
                        // t = d2 - lo; count += sign(t); flag |= bits(t) <u bits(width): FADD, LEA.HI, ISETP.OR per distance
                        unsigned int flag;
#define D2_F4(a, b, c, d)                                                                                                          \
    "sub.rn.f32 t, %" #a ", %37; mov.b32 u, t; shr.u32 s, u, 31; add.u32 %0, %0, s; setp.lt.or.u32 pf, u, %38, pf;\n\t"                  \
    "sub.rn.f32 t, %" #b ", %37; mov.b32 u, t; shr.u32 s, u, 31; add.u32 %1, %1, s; setp.lt.or.u32 pf, u, %38, pf;\n\t"                  \
    "sub.rn.f32 t, %" #c ", %37; mov.b32 u, t; shr.u32 s, u, 31; add.u32 %2, %2, s; setp.lt.or.u32 pf, u, %38, pf;\n\t"                  \
    "sub.rn.f32 t, %" #d ", %37; mov.b32 u, t; shr.u32 s, u, 31; add.u32 %3, %3, s; setp.lt.or.u32 pf, u, %38, pf;\n\t"
                        asm volatile("{\n\t.reg .pred pf;\n\t.reg .f32 t;\n\t.reg .b32 u, s;\n\t"
                                     "setp.ne.u32 pf, 0, 0;\n\t"
                                     D2_F4(5, 6, 7, 8) D2_F4(9, 10, 11, 12) D2_F4(13, 14, 15, 16) D2_F4(17, 18, 19, 20)
                                     D2_F4(21, 22, 23, 24) D2_F4(25, 26, 27, 28) D2_F4(29, 30, 31, 32) D2_F4(33, 34, 35, 36)
                                     "selp.u32 %4, 1, 0, pf;\n\t}"
                                     : "+r"(cnt4[0]), "+r"(cnt4[1]), "+r"(cnt4[2]), "+r"(cnt4[3]), "=r"(flag)
                                     : "f"(__uint_as_float(r0[0])), "f"(__uint_as_float(r0[1])), "f"(__uint_as_float(r0[2])), "f"(__uint_as_float(r0[3])),
                                       "f"(__uint_as_float(r0[4])), "f"(__uint_as_float(r0[5])), "f"(__uint_as_float(r0[6])), "f"(__uint_as_float(r0[7])),
                                       "f"(__uint_as_float(r0[8])), "f"(__uint_as_float(r0[9])), "f"(__uint_as_float(r0[10])), "f"(__uint_as_float(r0[11])),
                                       "f"(__uint_as_float(r0[12])), "f"(__uint_as_float(r0[13])), "f"(__uint_as_float(r0[14])), "f"(__uint_as_float(r0[15])),
                                       "f"(__uint_as_float(r0[16])), "f"(__uint_as_float(r0[17])), "f"(__uint_as_float(r0[18])), "f"(__uint_as_float(r0[19])),
                                       "f"(__uint_as_float(r0[20])), "f"(__uint_as_float(r0[21])), "f"(__uint_as_float(r0[22])), "f"(__uint_as_float(r0[23])),
                                       "f"(__uint_as_float(r0[24])), "f"(__uint_as_float(r0[25])), "f"(__uint_as_float(r0[26])), "f"(__uint_as_float(r0[27])),
                                       "f"(__uint_as_float(r0[28])), "f"(__uint_as_float(r0[29])), "f"(__uint_as_float(r0[30])), "f"(__uint_as_float(r0[31])),
                                       "f"(lo), "r"(wbits));
#undef D2_F4
                        if (__any_sync(0xffffffffu, flag != 0u)) { // some distance of this warp's chunk lies in [lo, hi'): collect with the same test
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                asm volatile("{\n\t.reg .pred pi;\n\t.reg .f32 t;\n\t.reg .b32 u;\n\t"
                                             "sub.rn.f32 t, %1, %2;\n\t"
                                             "mov.b32 u, t;\n\t"
                                             "setp.lt.u32 pi, u, %3;\n\t"
                                             "@pi st.shared.f32 [%0], %1;\n\t"
                                             "@pi add.u32 %0, %0, 128;\n\t}"
                                             : "+r"(paddr)
                                             : "f"(__uint_as_float(r0[q])), "f"(lo), "r"(wbits)
                                             : "memory");
                        }
                    } else if (!open_low) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) visit(__uint_as_float(r0[q]), cnt4[q & 3]);
                    } else { // lo = -inf (nothing lies below, cold start): plain compare against hi
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            asm volatile("{\n\t.reg .pred pi;\n\t"
                                         "setp.lt.f32 pi, %1, %2;\n\t"
                                         "@pi st.shared.f32 [%0], %1;\n\t"
                                         "@pi add.u32 %0, %0, 128;\n\t}"
                                         : "+r"(paddr)
                                         : "f"(__uint_as_float(r0[q])), "f"(hi)
                                         : "memory");
                    }
                    // the next chunk may add up to 32 entries per thread: compact when any thread could overflow
                    if (__builtin_expect(__any_sync(0xffffffffu, paddr - priv_base > (uint32_t)(D2_PRIV - 32) * 128u), 0)) compact();
                    below += (unsigned long long)(cnt4[0] + cnt4[1] + cnt4[2] + cnt4[3]) * wgt;
                }
            }
        }
        compact();
        if (MODE == MODE_COLLECT && count) dist_flush2(mybuf, count, p.cand, p.cand_count, p.capacity);
        for (int o = 16; o; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below) atomicAdd(p.below, below);
    }
    tc_fence_before();
    __syncthreads();
    if (MODE == MODE_HIST)
        for (int b = threadIdx.x; b < HIST_BINS; b += blockDim.x) {
            const unsigned long long cc = HIST32 ? (unsigned long long)shist32[b] : shist[b];
            if (cc) atomicAdd(&p.hist[b], cc);
        }
    if (warp == D2_CWARPS) tmem_dealloc(tmem, 512);
}

} // namespace tc
} // namespace svgdb
