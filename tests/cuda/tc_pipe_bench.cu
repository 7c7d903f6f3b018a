// tc_pipe_bench.cu — micro-benchmark of the pair-interaction kernel's MMA pattern on one SM (development aid).
// Warp 0 issues, per step,  2 x [ NPV x PV (TS, M=128, N=NV, K=16) ; NS x S (TS or SS, M=128, N=128, K=16) ]
// exactly like phi2_tc32_kernel's steady state; optional "noise" warps run tcgen05.ld / tcgen05.st / MUFU loops
// next to it.  Prints cycles per step against the tensor-pipe ideal.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_pipe_bench tests/cuda/tc_pipe_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../../svgdcpp_b200/csrc/tc_common.cuh"

using namespace svgdb::tc;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
    } while (0)

__device__ __forceinline__ float ex2a(float x)
{
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// NOISE bit 0: tcgen05.ld x32 loops, bit 1: tcgen05.st x16 loops, bit 2: 32 MUFU per iteration
template <int NV, int NPV, int NS, bool S_SS, int NOISE>
__global__ void __launch_bounds__(320) pipe_kernel(int steps, long long *out, int *err, float *sink)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;            // 128 x 128 B
    uint8_t *sB = smem + 16384;    // 128 x 128 B
    uint8_t *sV = smem + 32768;    // 4 x (up to 128 x 128 B)
    uint64_t *bar = (uint64_t *)(smem + 98304);
    uint32_t *holder = (uint32_t *)(bar + 2);
    volatile int *stop = (volatile int *)(holder + 1);
    for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u; // fp16 1.0 pairs
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); *stop = 0; }
    fence_proxy_async();
    const int warp = threadIdx.x >> 5;
    if (warp == 9) tmem_alloc(holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (warp == 9) {
        const uint32_t idesc_s = make_idesc_f16(128, 128), idesc_v = make_idesc_f16(128, NV);
        const uint32_t al = desc_lo_k_sw128(smem_u32(sA)), bl = desc_lo_k_sw128(smem_u32(sB)), vl = desc_lo_k_sw128(smem_u32(sV));
        long long t0 = clock64();
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                if (elect_one()) {
                    const uint32_t dP = tmem + 256 + w * 64, e = tmem + w * 128, dS = tmem + w * 128, aT = tmem + 384 + w * 64;
#pragma unroll
                    for (int k = 0; k < NPV; ++k) umma_f16_ts2<true>(dP, e + (k & 3) * 8 + ((k >> 2) & 1) * 64, vl + (k >> 2) * 1024 + (k & 3) * 2, idesc_v);
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
                        if (S_SS) umma_f16_ss2<true>(dS, al + (k & 3) * 2, bl + (k & 3) * 2, idesc_s);
                        else umma_f16_ts2<true>(dS, aT + (k & 7) * 8, bl + (k & 3) * 2, idesc_s);
                    }
                }
                __syncwarp();
            }
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0, err, 1);
        long long t2 = clock64();
        *stop = 1;
        if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (warp < 8 && NOISE != 0) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + (warp >> 2) * 128 + 96 + lane_base; // columns the MMAs also touch (values are irrelevant)
        float acc = 0.f;
        long long it = 0;
        while (!*stop) {
            uint32_t r[32];
            if (NOISE & 1) { tmem_ld32(tS, r); tmem_ld_wait(); } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) r[k] = (uint32_t)(k + it);
            }
            if (NOISE & 4) {
#pragma unroll
                for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(ex2a(__uint_as_float(r[k]) * 1e-30f));
            }
            if (NOISE & 2) {
                uint32_t pk[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) pk[k] = r[2 * k] ^ r[2 * k + 1];
                tmem_st16(tS, pk);
                tmem_st_wait();
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) acc += __uint_as_float(r[k]);
            }
            ++it;
        }
        if (acc == 123.456f) sink[threadIdx.x] = acc;
        if (threadIdx.x == 0) out[2] = it;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, 512);
}

template <int NV, int NPV, int NS, bool S_SS, int NOISE>
static void run(long long *dout, int *derr, float *sink)
{
    const int steps = 64;
    const size_t smem = 98304 + 64 + 1024;
    CK(cudaFuncSetAttribute(pipe_kernel<NV, NPV, NS, S_SS, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pipe_kernel<NV, NPV, NS, S_SS, NOISE><<<1, 320, smem>>>(steps, dout, derr, sink);
    pipe_kernel<NV, NPV, NS, S_SS, NOISE><<<1, 320, smem>>>(steps, dout, derr, sink);
    CK(cudaDeviceSynchronize());
    long long h[3];
    CK(cudaMemcpy(h, dout, 24, cudaMemcpyDeviceToHost));
    const double ideal = 2.0 * (NPV * NV / 2.0 + NS * 64.0);
    printf("PV %2d x N=%3d | S %2d x %s | noise %d : issue %8.1f  total %8.1f cyc/step  (ideal %6.0f, x%.2f)  noise iters/step %.1f\n", NPV, NV, NS,
           S_SS ? "SS" : "TS", NOISE, (double)h[0] / steps, (double)h[1] / steps, ideal, (double)h[1] / steps / ideal, NOISE ? (double)h[2] / steps : 0.0);
}

// The kernel's issue structure: per unit  [elect: 8 x PV (N=64) + commit]  [elect: 8 x S (N=64) + commit], 4 units per step,
// operand addresses derived from a per-thread (non-uniform) slot counter when NONUNIFORM = 1.
template <int NONUNIFORM>
__global__ void __launch_bounds__(320) block_kernel(int steps, long long *out, int *err, volatile int *slotp)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + 98304);
    uint32_t *holder = (uint32_t *)(bar + 16);
    for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int k = 0; k < 16; ++k) mbar_init(bar + k, 1); fence_barrier_init(); }
    fence_proxy_async();
    const int warp = threadIdx.x >> 5;
    if (warp == 9) tmem_alloc(holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (warp == 9) {
        const uint32_t idesc = make_idesc_f16(128, 64);
        const uint32_t base = desc_lo_k_sw128(smem_u32(smem));
        long long t0 = clock64();
        for (int s = 0; s < steps; ++s) {
            const uint32_t slot = NONUNIFORM ? (uint32_t)((s + *slotp) & 3) : 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dP = tmem + 256 + (b >> 1) * 64, e = tmem + b * 64, vl = base + slot * 1024 + 2048 + (b & 1) * 512;
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_f16_ts2<true>(dP, e + (k & 3) * 8, vl + (k >> 2) * 1024 + (k & 3) * 2, idesc);
                    umma_commit(bar + 1 + b);
                }
                __syncwarp();
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dS = tmem + b * 64, aT = tmem + 384 + (b >> 1) * 64, bl = base + slot * 1024 + (b & 1) * 512;
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_f16_ts2<true>(dS, aT + k * 8, bl + (k & 3) * 2, idesc);
                    umma_commit(bar + 5 + b);
                }
                __syncwarp();
            }
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0, err, 1);
        long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, 512);
}

template <int NONUNIFORM>
static void runb(long long *dout, int *derr, float *sink)
{
    const int steps = 64;
    const size_t smem = 98304 + 256 + 1024;
    CK(cudaFuncSetAttribute(block_kernel<NONUNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(sink, 0, 4));
    block_kernel<NONUNIFORM><<<1, 320, smem>>>(steps, dout, derr, (volatile int *)sink);
    block_kernel<NONUNIFORM><<<1, 320, smem>>>(steps, dout, derr, (volatile int *)sink);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
    printf("blocks of 8 + commits, 64 MMAs (N=64) per step, nonuniform %d : issue %8.1f  total %8.1f cyc/step  (ideal 2048)\n", NONUNIFORM,
           (double)h[0] / steps, (double)h[1] / steps);
}

// tcgen05.ld bandwidth: NW warps (lane quadrant = warp % 4) stream 32x32b.x32 loads (4 KB each) with no math in between.
template <int NW, int PER_WAIT>
__global__ void __launch_bounds__(NW * 32 + 32) ldtm_kernel(int iters, long long *out, unsigned *sink)
{
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (warp == NW) tmem_alloc(&holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (warp < NW) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
        unsigned acc = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t r[PER_WAIT][32];
#pragma unroll
            for (int k = 0; k < PER_WAIT; ++k) tmem_ld32(base + ((it + k) & 3) * 128, r[k]);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < PER_WAIT; ++k) acc ^= r[k][0] ^ r[k][31];
        }
        long long t1 = clock64();
        if (acc == 0x12345u) sink[threadIdx.x] = acc;
        if (threadIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NW) tmem_dealloc(tmem, 512);
}
template <int NW, int PER_WAIT>
static void runl(long long *dout, float *sink)
{
    const int iters = 256;
    ldtm_kernel<NW, PER_WAIT><<<1, NW * 32 + 32>>>(iters, dout, (unsigned *)sink);
    ldtm_kernel<NW, PER_WAIT><<<1, NW * 32 + 32>>>(iters, dout, (unsigned *)sink);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
    printf("tcgen05.ld x32: %2d warps, %d loads per wait: %7.1f B/clk/SM  (%.0f cycles per 4 KB load per warp)\n", NW, PER_WAIT,
           (double)NW * iters * PER_WAIT * 4096.0 / (double)h, (double)h / (iters * PER_WAIT));
}

int main()
{
    long long *dout;
    int *derr;
    float *sink;
    CK(cudaMalloc(&dout, 24));
    CK(cudaMalloc(&derr, 4));
    CK(cudaMalloc(&sink, 4096));
    CK(cudaMemset(derr, 0, 4));
    CK(cudaMemset(dout, 0, 24));
    runl<4, 1>(dout, sink); runl<4, 2>(dout, sink); runl<8, 1>(dout, sink); runl<8, 2>(dout, sink); runl<16, 1>(dout, sink); runl<16, 2>(dout, sink);
    // issue-rate floor: tiny MMAs
    run<8, 16, 0, false, 0>(dout, derr, sink);
    run<16, 16, 0, false, 0>(dout, derr, sink);
    run<32, 16, 0, false, 0>(dout, derr, sink);
    // single-kind streams
    run<64, 16, 0, false, 0>(dout, derr, sink);
    run<80, 16, 0, false, 0>(dout, derr, sink);
    run<128, 16, 0, false, 0>(dout, derr, sink);
    run<64, 0, 8, false, 0>(dout, derr, sink);
    run<64, 0, 8, true, 0>(dout, derr, sink);
    // the kernel's pattern
    run<64, 16, 8, false, 0>(dout, derr, sink);
    run<64, 16, 8, true, 0>(dout, derr, sink);
    run<64, 8, 8, false, 0>(dout, derr, sink);
    run<128, 8, 8, false, 0>(dout, derr, sink);
    // ... with TMEM / MUFU traffic from the 8 exp warps
    run<64, 16, 8, false, 1>(dout, derr, sink);
    run<64, 16, 8, false, 2>(dout, derr, sink);
    run<64, 16, 8, false, 3>(dout, derr, sink);
    run<64, 16, 8, false, 4>(dout, derr, sink);
    run<64, 16, 8, false, 7>(dout, derr, sink);
    run<64, 16, 8, true, 7>(dout, derr, sink);
    // the same pattern through the kernel's blocks of 8 (elect + commit per block)
    runb<0>(dout, derr, sink);
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    printf("timeout tag %d\n", herr);
    return 0;
}
