// tc_mma_bench.cu — micro-benchmark of tcgen05.mma issue/latency behaviour on one SM (development aid).
// Measures cycles per MMA (kind::f16, M=128, K=16, SS and TS forms) for different N, with the MMAs
// accumulating into 1, 2 or 4 independent TMEM tiles, to decide how the SVGD kernels must order their MMAs.
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

// MODE 0: SS, MODE 1: TS (A from TMEM columns 384..).  Warp 0 issues (warp-uniform code, one elected lane).
template <int N, int NACC, int MODE>
__global__ void __launch_bounds__(128) mma_bench_kernel(int reps, long long *out, int *err)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;            // 128 x 128 B
    uint8_t *sB = smem + 16384;    // 256 x 128 B
    uint64_t *bar = (uint64_t *)(smem + 16384 + 32768);
    uint32_t *holder = (uint32_t *)(bar + 2);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u; // fp16 1.0 pairs
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint32_t al = desc_lo_k_sw128(smem_u32(sA)), bl = desc_lo_k_sw128(smem_u32(sB));
        constexpr int stride = NACC > 1 ? 384 / NACC : 0; // accumulator tiles spread over columns [0,384)
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 4 * NACC) {
#pragma unroll
            for (int u = 0; u < 4 * NACC; ++u) {
                const uint32_t d = tmem + (u % NACC) * stride;
                if (elect_one()) {
                    if (MODE == 0) umma_f16_ss2<true>(d, al + (u & 3) * 2, bl + (u & 3) * 2, idesc);
                    else umma_f16_ts2<true>(d, tmem + 384 + (u & 7) * 8, bl + (u & 3) * 2, idesc);
                }
            }
        }
        __syncwarp();
        long long t1 = clock64();
        if (elect_one()) umma_commit(bar);
        mbar_wait(bar, 0, err, 1);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int MODE>
static void run(long long *dout, int *derr, size_t smem)
{
    const int reps = 768;
    CK(cudaFuncSetAttribute(mma_bench_kernel<N, NACC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_bench_kernel<N, NACC, MODE><<<1, 128, smem>>>(reps, dout, derr); // warm-up
    mma_bench_kernel<N, NACC, MODE><<<1, 128, smem>>>(reps, dout, derr);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
    printf("%-4s %5d %5d %12.1f %12.1f   (ideal %d)\n", MODE ? "TS" : "SS", N, NACC, (double)h[0] / reps, (double)h[1] / reps, N / 2);
}

int main()
{
    long long *dout;
    int *derr;
    CK(cudaMalloc(&dout, 16));
    CK(cudaMalloc(&derr, 4));
    CK(cudaMemset(derr, 0, 4));
    size_t smem = 16384 + 32768 + 64 + 1024;
    printf("%-4s %5s %5s %12s %12s\n", "mode", "N", "nacc", "issue cyc/MMA", "total cyc/MMA");
    run<64, 1, 0>(dout, derr, smem);  run<64, 2, 0>(dout, derr, smem);  run<64, 4, 0>(dout, derr, smem);
    run<80, 1, 0>(dout, derr, smem);  run<80, 2, 0>(dout, derr, smem);  run<80, 4, 0>(dout, derr, smem);
    run<128, 1, 0>(dout, derr, smem); run<128, 2, 0>(dout, derr, smem); run<128, 3, 0>(dout, derr, smem);
    run<256, 1, 0>(dout, derr, smem);
    run<64, 1, 1>(dout, derr, smem);  run<64, 2, 1>(dout, derr, smem);  run<64, 4, 1>(dout, derr, smem);
    run<80, 1, 1>(dout, derr, smem);  run<80, 2, 1>(dout, derr, smem);  run<80, 4, 1>(dout, derr, smem);
    run<128, 1, 1>(dout, derr, smem); run<128, 2, 1>(dout, derr, smem); run<128, 3, 1>(dout, derr, smem);
    run<160, 1, 1>(dout, derr, smem); run<160, 2, 1>(dout, derr, smem);
    run<256, 1, 1>(dout, derr, smem);
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    printf("timeout tag %d\n", herr);
    return 0;
}
