// kernels_dist_wide.cuh — the tensor-core distance pass of the exact median bandwidth (SVGDB_PRECISION_TC32) for 64 < d <= 256.
//
// Same counting / collecting contract and the same arithmetic as kernels_dist_tc.cuh (centred particles in two bf16 terms,
// D2 from hi.hi + lo.hi + hi.lo with fp32 accumulation, norms as three-term bf16 splits in a 16-column K chunk, keys = IEEE bits of
// (double)D2; FOLD passes carry -lo in that chunk).  Reference semantics: Kernel/GaussianRBFKernel.hpp:168-188, 222-254.
//
// Mapping (TMEM: 512 columns): the row operand -2 [hi | lo] of ONE 128-particle i-tile takes DP columns (DP = d padded to 64),
// four accumulator buffers of 128 x 64 distances take 256.  A pipeline stage holds the column operand of one 64-particle unit
// ([hi | lo], 2 DP bf16 per particle, 2 KC boxes of 64 x 64 with SWIZZLE_128B) and its 2 KB norm chunk.  Warps: 8 counting warps
// (row quadrant x 32-column half), one TMA producer, one MMA issuer.  Per unit the tensor pipe needs (12 KC + 1) x 32 cycles
// (1568 at DP = 256) against ~100 issue cycles per counting warp: the pass is tensor-bound by a wide margin.
// Work: i-tiles are dealt cyclically to the ranks and visited in folded order (long row, short row, ...); for i-tile `it` the
// column units 2 it .. n_junits - 1 (upper triangle at tile level: the diagonal tile counts once, the others twice).
#pragma once
#include "kernels_dist_tc.cuh"

namespace svgdb {
namespace tc {

template <int DP>
struct DWCfg {
    static_assert(DP == 128 || DP == 192 || DP == 256, "padded dimension");
    static constexpr int KC = DP / 64;
    static constexpr int CWARPS = 8;
    static constexpr int NB = 4; // accumulator buffers
    static constexpr uint32_t XB_BYTES = 8192u * KC * 2u;
    static constexpr uint32_t STAGE = XB_BYTES + 2048u;
    static constexpr int WBUF_STRIDE = 1024; // floats between the warps' staging buffers (TC_WBUF are used): the 32 KB shared histogram of MODE_HIST aliases exactly this area
    static constexpr uint32_t FIXED = P2_AEX_BYTES + CWARPS * WBUF_STRIDE * 4 + CWARPS * 32 * D2_PRIV * 4 + 512 + 1024;
    static constexpr int STAGES_FIT = (int)((227u * 1024u - FIXED) / STAGE);
    static constexpr int STAGES = STAGES_FIT > 4 ? 4 : STAGES_FIT;
    static constexpr uint32_t SMEM = STAGES * STAGE + FIXED;
    static constexpr uint32_t COL_A = 256;
    static_assert(STAGES >= 2, "two pipeline stages must fit");
    static_assert(HIST_BINS * 8 <= CWARPS * WBUF_STRIDE * 4, "the shared histogram must not reach into the per-thread staging slots");
};
constexpr int DW_THREADS = (8 + 2) * 32;

// Operand rows of the wide distance pass, one warp per particle:
//   XA[row] = -2 [hi(DP) | lo(DP)] (row operand -> TMEM),  XB[row] = [hi(DP) | lo(DP)] (column operand, TMA),  UA / WB / rt as in
//   split_dist2_kernel.
__global__ void split_distw_kernel(const double *__restrict__ X, const double *__restrict__ colsum, int64_t n, int64_t n_rows_a, int64_t n_rows_b,
                                   int d, int dp, __nv_bfloat16 *__restrict__ XA, __nv_bfloat16 *__restrict__ XB, __nv_bfloat16 *__restrict__ UA,
                                   __nv_bfloat16 *__restrict__ WB, double *__restrict__ rt)
{
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows_a) return;
    double s = 0.0;
    for (int k = lane; k < dp; k += 32) {
        double xc = 0.0;
        if (row < n && k < d) xc = X[row * d + k] - colsum[k] / (double)n;
        s += xc * xc;
        const __nv_bfloat16 hi = __float2bfloat16_rn((float)xc);
        const __nv_bfloat16 lo = __float2bfloat16_rn((float)(xc - (double)__bfloat162float(hi)));
        XA[row * (2 * dp) + k] = __float2bfloat16_rn(-2.0f * __bfloat162float(hi)); // exact: a power-of-two multiple
        XA[row * (2 * dp) + dp + k] = __float2bfloat16_rn(-2.0f * __bfloat162float(lo));
        if (row < n_rows_b) {
            XB[row * (2 * dp) + k] = hi;
            XB[row * (2 * dp) + dp + k] = lo;
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
                lane < 3 ? one : lane == 3 ? r0 : lane == 4 ? r1 : lane == 5 ? r2 : lane < 9 ? one : zero;
    }
}

// Column sums for centring, any d <= 256 (blockDim = 256: thread k accumulates column k over a grid-stride range of rows).
__global__ void colsum_wide_kernel(const double *__restrict__ X, int64_t n, int d, double *__restrict__ sum)
{
    const int k = threadIdx.x;
    if (k >= d) return;
    double acc = 0.0;
    for (int64_t row = blockIdx.x; row < n; row += gridDim.x) acc += X[row * d + k];
    atomicAdd(&sum[k], acc);
}

struct DistWArgs {
    const __nv_bfloat16 *XA; // [n_pad128 + 128][2 DP]
    const __nv_bfloat16 *UA; // [n_pad128 + 128][16]
    const __nv_bfloat16 *WB; // [n_pad128 / 128][4 KB]
    int64_t n_total;
    int n_junits;                            // 64-particle column units
    int tile_offset, tile_stride, n_itiles;  // this rank owns i-tiles offset, offset + stride, ... (n_itiles of them)
    float lo_f, hi_f;
    unsigned int fold_l01, fold_l2;
    unsigned int width_bits;
    int open_low;
    unsigned long long lo_key;
    int shift;
    unsigned long long *below, *hist, *cand, *cand_count;
    unsigned long long capacity;
    int *err;
};

struct DWSeg { int it, jb, je; };
struct DWCursor { int l; long long base; };
__device__ __forceinline__ bool dw_segment(const DistWArgs &p, DWCursor &cur, long long &pos, long long end, DWSeg &s)
{
    if (pos >= end) return false;
    for (;;) {
        const int lf = (cur.l & 1) ? p.n_itiles - 1 - (cur.l >> 1) : (cur.l >> 1); // folded order
        const int it = p.tile_offset + p.tile_stride * lf;
        const int j0 = 2 * it;
        const long long len = max(0, p.n_junits - j0);
        if (pos < cur.base + len) {
            s.it = it;
            s.jb = j0 + (int)(pos - cur.base);
            const long long seg_end = min(end, cur.base + len);
            s.je = s.jb + (int)(seg_end - pos);
            pos = seg_end;
            return true;
        }
        cur.base += len;
        ++cur.l;
        if (cur.l >= p.n_itiles) return false;
    }
}
__device__ __forceinline__ long long dw_total_units(const DistWArgs &p)
{
    long long tot = 0;
    for (int l = 0; l < p.n_itiles; ++l) tot += max(0, p.n_junits - 2 * (p.tile_offset + p.tile_stride * l));
    return tot;
}

template <int DP, int MODE, bool GATED, bool FOLD>
__global__ void __launch_bounds__(DW_THREADS, 1)
distw_tc32_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ DistWArgs p)
{
    using Cfg = DWCfg<DP>;
    constexpr int KC = Cfg::KC, STAGES = Cfg::STAGES, NB = Cfg::NB, CW = Cfg::CWARPS;
    constexpr uint32_t STAGE = Cfg::STAGE, XB_BYTES = Cfg::XB_BYTES, COL_A = Cfg::COL_A;
    static_assert(!(FOLD && MODE == MODE_HIST), "histogram passes run unfolded");

    const long long units = dw_total_units(p);
    const long long u_beg = units * blockIdx.x / gridDim.x, u_end = units * (blockIdx.x + 1) / gridDim.x;
    if (u_beg >= u_end) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sAex = smem + STAGES * STAGE;
    float *wbuf = (float *)(sAex + P2_AEX_BYTES);  // [CW][WBUF_STRIDE]  per-warp staging (TC_WBUF entries used)
    float *priv = wbuf + CW * Cfg::WBUF_STRIDE;    // [CW][D2_PRIV][32]  per-thread staging, lane-interleaved
    uint64_t *bars = (uint64_t *)(priv + CW * 32 * D2_PRIV);
    uint64_t *full = bars;                 // [STAGES]
    uint64_t *empty = full + STAGES;       // [STAGES]
    uint64_t *s_full = empty + STAGES;     // [NB] accumulator buffer complete (commit)
    uint64_t *s_free = s_full + NB;        // [NB] accumulator buffer is in the counting warps' registers (8 warp arrivals)
    uint64_t *a_ready = s_free + NB;       // row operand in TMEM / shared memory (8 warp arrivals)
    uint64_t *seg_done = a_ready + 1;      // every MMA of the segment complete (commit)
    uint32_t *tmem_holder = (uint32_t *)(seg_done + 1);
    unsigned long long *shist = (unsigned long long *)wbuf; // [HIST_BINS] 64-bit, MODE_HIST only: aliases the staging areas

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(s_full + s, 1); mbar_init(s_free + s, CW); }
        mbar_init(a_ready, CW);
        mbar_init(seg_done, 1);
        fence_barrier_init();
    }
    if (MODE == MODE_HIST)
        for (int b = threadIdx.x; b < HIST_BINS; b += blockDim.x) shist[b] = 0ull;
    if (warp == CW) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == CW) { // ---- TMA producer
        long long pos = u_beg;
        DWCursor cur{0, 0};
        DWSeg sg;
        uint32_t g = 0;
        bool ok = true;
        while (ok && dw_segment(p, cur, pos, u_end, sg)) {
            for (int ju = sg.jb; ok && ju < sg.je; ++ju, ++g) {
                const uint32_t slot = g % STAGES, use = g / STAGES;
                if (!mbar_wait(empty + slot, (use & 1) ^ 1, p.err, 150)) { ok = false; break; }
                if (elect_one()) {
                    uint8_t *st = smem + slot * STAGE;
                    mbar_arrive_expect_tx(full + slot, STAGE);
#pragma unroll
                    for (int c = 0; c < 2 * KC; ++c) tma_load_2d(st + c * 8192, &mapB, c * 64, ju * 64, full + slot); // hi chunks, then lo chunks
                    bulk_load_1d(st + XB_BYTES, reinterpret_cast<const uint8_t *>(p.WB) + (size_t)ju * 2048, 2048, full + slot);
                }
                __syncwarp();
            }
        }
    } else if (warp == CW + 1) { // ---- MMA issuer
        const uint32_t idesc = make_idesc_bf16(TC_TILE, 64);
        const uint32_t st_lo0 = desc_lo_k_sw128(smem_u32(smem));
        const uint32_t aex_lo = desc_lo_k_sw128(smem_u32(sAex)) | DESC_LO_K_NOSW_LBO;
        const uint32_t wb_lo0 = desc_lo_k_sw128(smem_u32(smem + XB_BYTES)) | DESC_LO_K_NOSW_LBO;
        const uint32_t aT = tmem + COL_A;
        long long pos = u_beg;
        DWCursor cur{0, 0};
        DWSeg sg;
        uint32_t g = 0;
        bool ok = true;
        for (uint32_t seg = 0; ok && dw_segment(p, cur, pos, u_end, sg); ++seg) {
            if (!mbar_wait(a_ready, seg & 1, p.err, 160)) { ok = false; break; }
            for (int ju = sg.jb; ok && ju < sg.je; ++ju, ++g) {
                const uint32_t slot = g % STAGES, use = g / STAGES, buf = g % NB, bu = g / NB;
                if (!mbar_wait(full + slot, use & 1, p.err, 161)) { ok = false; break; }
                if (!mbar_wait(s_free + buf, (bu & 1) ^ 1, p.err, 162)) { ok = false; break; }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dS = tmem + buf * 64;
                    const uint32_t bh = st_lo0 + slot * (STAGE >> 4), bl = bh + ((8192u * KC) >> 4);
#pragma unroll
                    for (int c = 0; c < KC; ++c)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if (c == 0 && ks == 0) umma_f16_ts2<false>(dS, aT, bh, idesc);
                            else umma_f16_ts2<true>(dS, aT + c * 32 + ks * 8, bh + c * (8192 >> 4) + ks * 2, idesc);        // (-2 hi_i) . hi_j
                        }
#pragma unroll
                    for (int c = 0; c < KC; ++c)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16_ts2<true>(dS, aT + DP / 2 + c * 32 + ks * 8, bh + c * (8192 >> 4) + ks * 2, idesc);    // (-2 lo_i) . hi_j
#pragma unroll
                    for (int c = 0; c < KC; ++c)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16_ts2<true>(dS, aT + c * 32 + ks * 8, bl + c * (8192 >> 4) + ks * 2, idesc);             // (-2 hi_i) . lo_j
                    umma_f16_ss_desc(dS, aex_lo, DESC_HI_K_NOSW, wb_lo0 + slot * (STAGE >> 4), DESC_HI_K_NOSW, idesc);      // + r_i + r_j (- lo)
                    umma_commit(s_full + buf);
                    umma_commit(empty + slot);
                    if (ju + 1 == sg.je) umma_commit(seg_done);
                }
                __syncwarp();
            }
        }
    } else { // ---- counting warps: thread = row i of the tile, 32-column half h of every unit
        const int h = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float *mybuf = wbuf + warp * Cfg::WBUF_STRIDE;
        const uint32_t priv_base = smem_u32(priv + warp * (D2_PRIV * 32) + lane);
        const uint32_t wbuf_base = smem_u32(mybuf);
        unsigned int count = 0;
        unsigned long long below = 0ull;
        uint32_t paddr = priv_base;
        unsigned int cur_wgt = 1u;
        auto compact = [&]() {
            const uint32_t mine = (paddr - priv_base) >> 7;
            const unsigned int tot = __reduce_add_sync(0xffffffffu, mine);
            if (tot) {
                if (MODE == MODE_HIST) {
                    unsigned int run_bin = 0xffffffffu, run_cnt = 0u;
                    auto flush_run = [&]() {
                        if (run_cnt) atomicAdd(&shist[run_bin], (unsigned long long)run_cnt);
                    };
                    for (uint32_t e = 0; e < mine; ++e) {
                        const unsigned long long bin = (dist_key(lds_f32(priv_base + 128u * e)) - p.lo_key) >> p.shift;
                        if (bin < (unsigned long long)HIST_BINS) {
                            if ((unsigned int)bin == run_bin) run_cnt += cur_wgt;
                            else { flush_run(); run_bin = (unsigned int)bin; run_cnt = cur_wgt; }
                        }
                    }
                    flush_run();
                } else {
                    if (count + tot * cur_wgt > (unsigned int)TC_WBUF) { dist_flush2(mybuf, count, p.cand, p.cand_count, p.capacity); count = 0; }
                    if (tot * cur_wgt > (unsigned int)TC_WBUF) {
                        for (uint32_t e = 0; e < mine; ++e)
                            dist_append_global2(lds_f32(priv_base + 128u * e) + (FOLD ? p.lo_f : 0.0f), cur_wgt, p.cand, p.cand_count, p.capacity);
                    } else {
                        uint32_t incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        uint32_t dst = wbuf_base + 4u * (count + (incl - mine) * cur_wgt);
                        for (uint32_t e = 0; e < mine; ++e) {
                            const float v = lds_f32(priv_base + 128u * e) + (FOLD ? p.lo_f : 0.0f);
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
        DWCursor cur{0, 0};
        DWSeg sg;
        uint32_t g = 0;
        bool ok = true;
        for (uint32_t seg = 0; ok && dw_segment(p, cur, pos, u_end, sg); ++seg) {
            const int64_t iw0 = (int64_t)sg.it * TC_TILE;
            const int64_t i = iw0 + row;
            const bool row_valid = i < p.n_total;
            const float lo = row_valid ? p.lo_f : -INFINITY, hi = row_valid ? p.hi_f : -INFINITY;
            const unsigned int wbits = row_valid ? p.width_bits : 0u;
            const bool open_low = p.open_low != 0;
            if (seg > 0) { // the previous segment's MMAs still read the row operand
                if (!mbar_wait(seg_done, (seg - 1) & 1, p.err, 170)) { ok = false; break; }
                tc_fence_after();
            }
            { // row operand -2 [hi | lo] -> TMEM: hi half by the h = 0 warp, lo half by the h = 1 warp; norm chunk -> shared memory
                const uint4 *src = reinterpret_cast<const uint4 *>(p.XA + i * (2 * DP) + DP * h);
                const uint32_t tA = tmem + COL_A + (DP / 2) * h + lane_base;
#pragma unroll 1
                for (int k = 0; k < DP / 32; ++k) {
                    uint32_t v[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 x = __ldg(src + 4 * k + q);
                        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                    }
                    tmem_st16(tA + 16 * k, v);
                }
                if (h == 0) {
                    const uint4 *usrc = reinterpret_cast<const uint4 *>(p.UA + i * 16);
                    const uint32_t aex = smem_u32(sAex) + p2_ex_offset((uint32_t)row, 0);
                    uint4 ua0 = __ldg(usrc), ua1 = __ldg(usrc + 1);
                    if (FOLD) { ua0.w = p.fold_l01; ua1.x = p.fold_l2; }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(aex), "r"(ua0.x), "r"(ua0.y), "r"(ua0.z), "r"(ua0.w) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(aex + 128u), "r"(ua1.x), "r"(ua1.y), "r"(ua1.z), "r"(ua1.w) : "memory");
                    fence_proxy_async();
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
            }
            for (int ju = sg.jb; ok && ju < sg.je; ++ju, ++g) {
                const bool has_diag = (ju >> 1) == sg.it;
                const unsigned int wgt = has_diag ? 1u : 2u; // (units left of the diagonal tile are not enumerated)
                const uint32_t buf = g % NB, bu = g / NB;
                const uint32_t tS = tmem + buf * 64 + 32 * h + lane_base;
                const int64_t j0 = (int64_t)ju * 64 + 32 * h;
                const int dcol = (int)(i - j0);
                if (!mbar_wait(s_full + buf, bu & 1, p.err, 172)) { ok = false; break; }
                tc_fence_after();
                uint32_t r0[32];
                tmem_ld32(tS, r0);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free + buf);
                if (wgt != cur_wgt) { compact(); cur_wgt = wgt; }
                if (has_diag) {
                    const uint32_t dz = FOLD ? __float_as_uint(-p.lo_f) : 0u;
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (dcol == q) r0[q] = dz;
                }
                unsigned int cnt4[4] = {0u, 0u, 0u, 0u};
                if (FOLD && MODE == MODE_COLLECT) {
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
                    if (!GATED || __any_sync(0xffffffffu, fminf(m0, m1) < __uint_as_float(wbits))) {
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
                } else if (!open_low) {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        asm volatile("{\n\t.reg .pred pi;\n\t.reg .f32 t;\n\t.reg .b32 u, s;\n\t"
                                     "sub.rn.f32 t, %2, %3;\n\t"
                                     "mov.b32 u, t;\n\t"
                                     "shr.u32 s, u, 31;\n\t"
                                     "add.u32 %0, %0, s;\n\t"
                                     "setp.lt.u32 pi, u, %4;\n\t"
                                     "@pi st.shared.f32 [%1], %2;\n\t"
                                     "@pi add.u32 %1, %1, 128;\n\t}"
                                     : "+r"(cnt4[q & 3]), "+r"(paddr)
                                     : "f"(__uint_as_float(r0[q])), "f"(lo), "r"(wbits)
                                     : "memory");
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
                if (__any_sync(0xffffffffu, paddr - priv_base > (uint32_t)(D2_PRIV - 32) * 128u)) compact();
                below += (unsigned long long)(cnt4[0] + cnt4[1] + cnt4[2] + cnt4[3]) * wgt;
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
            const unsigned long long cc = shist[b];
            if (cc) atomicAdd(&p.hist[b], cc);
        }
    if (warp == CW) tmem_dealloc(tmem, 512);
}

} // namespace tc
} // namespace svgdb
