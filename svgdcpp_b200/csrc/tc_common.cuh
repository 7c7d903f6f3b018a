// tc_common.cuh — hand-written sm_100a plumbing shared by the tensor-core kernels: mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / commit / ld / st) and the shared-memory /
// instruction descriptors.  Inline PTX only; no CUTLASS.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (also mirrored by cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace svgdb {
namespace tc {

// Development safety net: every spin on an mbarrier is bounded; on expiry the kernel records where it
// was stuck and bails out instead of hanging the GPU.
#ifndef SVGDB_SPIN_LIMIT
#define SVGDB_SPIN_LIMIT (1u << 22)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp.  Code that feeds tcgen05.mma / TMA must stay warp-uniform up to this
// point so that descriptors and addresses live in uniform registers (a lane-0 branch forces slow R2UR moves
// in front of every instruction: ~100 cycles per MMA measured).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// Returns false on timeout (and records `tag` in *err, first writer wins).
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *err, int tag)
{
    for (uint32_t spin = 0; spin < SVGDB_SPIN_LIMIT; ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        // somebody else already gave up: do not serialise one timeout after another
        if ((spin & 1023u) == 1023u && err && *reinterpret_cast<volatile int *>(err) != 0) return false;
    }
    if (err) atomicCAS(err, 0, tag);
    return false;
}

// ---- TMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *m, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// ---- thread-block clusters: TMA multicast and the barrier traffic that goes with it ---------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// every thread of every CTA of the cluster
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The box lands at the same shared-memory offset in every CTA of `mask` and completes bytes on the mbarrier at `bar`'s offset in each.
__device__ __forceinline__ void tma_load_2d_mc(void *dst_smem, const CUtensorMap *m, int c0, int c1, uint64_t *bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d_mc(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *holder_smem, uint32_t ncols) // whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) // whole warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------
// K-major operand tile stored as rows of exactly 128 bytes (64 bf16) with the 128B swizzle TMA writes:
// 8-row groups are 1024 B apart (SBO), LBO unused, version 1, layout SWIZZLE_128B (= 2).
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 instruction descriptor: f16 x f16 -> f32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA (issued by ONE thread) ---------------------------------------------------------------------
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// Variants taking the 64-bit shared-memory descriptors as (lo, hi) register pairs: the issuing thread keeps
// the constant hi word and advances only the 14-bit start-address field in lo (one integer add per MMA).
__device__ __forceinline__ uint32_t desc_lo_k_sw128(uint32_t smem_addr) { return (smem_addr >> 4) & 0x3FFFu; }
constexpr uint32_t DESC_HI_K_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29); // SBO = 1024 B, version 1, SWIZZLE_128B
template <bool ACC>
__device__ __forceinline__ void umma_f16_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc)
{
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
                 "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_K_SW128), "r"(idesc), "r"(ACC ? 1u : 0u)
                 : "memory");
}
template <bool ACC>
__device__ __forceinline__ void umma_f16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc)
{
    asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(DESC_HI_K_SW128), "r"(idesc), "r"(ACC ? 1u : 0u)
                 : "memory");
}
// same with a run-time accumulate flag
__device__ __forceinline__ void umma_f16_ts2r(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(DESC_HI_K_SW128), "r"(idesc), "r"(acc)
                 : "memory");
}

// kind::f8f6f4 (here: e5m2 x e5m2 -> f32, K = 32 per instruction), A from TMEM, B a K-major tile of 64-byte rows written by TMA with
// SWIZZLE_64B (8-row groups 512 B apart).  The instruction descriptor has the bit pattern of make_idesc_bf16 (format field 1 = E5M2).
constexpr uint32_t DESC_HI_K_SW64 = (512u >> 4) | (1u << 14) | (4u << 29); // SBO = 512 B, version 1, SWIZZLE_64B
template <bool ACC>
__device__ __forceinline__ void umma_f8_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc)
{
    asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(DESC_HI_K_SW64), "r"(idesc), "r"(ACC ? 1u : 0u)
                 : "memory");
}

// Arrives on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ... and on the mbarrier at `bar`'s offset in every CTA of `mask` (a pipeline stage filled by multicast is free when all readers are done)
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// ---- TMEM <-> registers: warp w touches lanes 32*(w%4) .. +31, thread = one lane (matrix row) ---------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) // lo -> bits [0,16), hi -> bits [16,32)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Explicit shared-space accesses through 32-bit addresses (pointers derived from the re-aligned dynamic smem
// base are generic to the compiler, which then emits 64-bit generic ST/LD with several address instructions).
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) // lo -> bits [0,16), hi -> bits [16,32)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

} // namespace tc
} // namespace svgdb
