// tc_selftest.cu — bring-up test of the hand-written tcgen05 / TMA / TMEM plumbing (tc_common.cuh).
//
//   stage 1 (SS): S[128x128] = A[128xK] * B[128xK]^T   bf16 operands via TMA (SW128), fp32 in TMEM
//   stage 2 (TS): E = bf16(S) written back to TMEM with tcgen05.st, Phi[128xNV] = E * VT^T  (A from TMEM)
// Both are compared with a host reference.  Every wait is bounded, so a plumbing bug reports an error
// tag instead of hanging the GPU.   Build: nvcc -gencode arch=compute_100a,code=sm_100a tc_selftest.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../svgdcpp_b200/csrc/tc_common.cuh"

using namespace svgdb::tc;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
    } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return (EncodeFn)fn;
}

// rows x cols bf16 row-major (cols contiguous); box = 64 cols x box_rows, 128B swizzle
static CUtensorMap make_map(EncodeFn enc, void *base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
    return m;
}

template <int KCH, int NV> // KCH = number of 64-wide K chunks of stage 1; NV = N of stage 2 (multiple of 16)
__global__ void __launch_bounds__(192) selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                       const __grid_constant__ CUtensorMap mapV, float *S_out, float *Phi_out, int *err)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                               // KCH x [128 x 128 B]
    uint8_t *sB = sA + KCH * 16384;                   // KCH x [128 x 128 B]
    uint8_t *sV = sB + KCH * 16384;                   // 2 x [NV x 128 B]
    uint64_t *bars = (uint64_t *)(sV + 2 * NV * 128); // 0: operands landed, 1: S done, 2: E stored, 3: Phi done
    uint32_t *tmem_holder = (uint32_t *)(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 128);
        mbar_init(&bars[3], 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t tS = tmem;         // 128 fp32 columns
    const uint32_t tE = tmem + 128;   // 64 columns of packed bf16 pairs
    const uint32_t tP = tmem + 256;   // NV fp32 columns

    if (warp == 4 && lane == 0) {
        // TMA producer
        mbar_arrive_expect_tx(&bars[0], KCH * 2 * 16384 + 2 * NV * 128);
        for (int c = 0; c < KCH; ++c) {
            tma_load_2d(sA + c * 16384, &mapA, c * 64, 0, &bars[0]);
            tma_load_2d(sB + c * 16384, &mapB, c * 64, 0, &bars[0]);
        }
        for (int c = 0; c < 2; ++c) tma_load_2d(sV + c * NV * 128, &mapV, c * 64, 0, &bars[0]);
    } else if (warp == 5 && lane == 0) {
        // MMA issuer
        if (!mbar_wait(&bars[0], 0, err, 1)) return;
        tc_fence_after();
        const uint32_t idesc1 = make_idesc_bf16(128, 128);
        for (int c = 0; c < KCH; ++c)
            for (int k = 0; k < 4; ++k) {
                uint64_t da = make_desc_k_sw128(smem_u32(sA + c * 16384) + k * 32);
                uint64_t db = make_desc_k_sw128(smem_u32(sB + c * 16384) + k * 32);
                umma_bf16_ss(tS, da, db, idesc1, (c | k) ? 1u : 0u);
            }
        umma_commit(&bars[1]);
        // stage 2 after the epilogue warps stored E
        if (!mbar_wait(&bars[2], 0, err, 2)) return;
        tc_fence_after();
        const uint32_t idesc2 = make_idesc_bf16(128, NV);
        for (int c = 0; c < 2; ++c)
            for (int k = 0; k < 4; ++k) {
                uint64_t db = make_desc_k_sw128(smem_u32(sV + c * NV * 128) + k * 32);
                umma_bf16_ts(tP, tE + (c * 4 + k) * 8, db, idesc2, (c | k) ? 1u : 0u);
            }
        umma_commit(&bars[3]);
    } else if (warp < 4) {
        // epilogue warps: thread = TMEM lane = matrix row
        const int row = warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        if (mbar_wait(&bars[1], 0, err, 3)) {
            tc_fence_after();
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tS + lane_base + c0, r);
                tmem_ld_wait();
                uint32_t packed[16];
                for (int q = 0; q < 32; ++q) S_out[row * 128 + c0 + q] = __uint_as_float(r[q]);
                for (int q = 0; q < 16; ++q) packed[q] = pack_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
                tmem_st16(tE + lane_base + c0 / 2, packed);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[2]);
            if (mbar_wait(&bars[3], 0, err, 4)) {
                tc_fence_after();
                for (int c0 = 0; c0 < NV; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tP + lane_base + c0, r);
                    tmem_ld_wait();
                    for (int q = 0; q < 16; ++q) Phi_out[row * NV + c0 + q] = __uint_as_float(r[q]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int KCH, int NV>
static int run_case()
{
    const int K = KCH * 64;
    EncodeFn enc = get_encode();
    std::vector<__nv_bfloat16> hA(128 * K), hB(128 * K), hV((size_t)NV * 128);
    std::vector<float> fA(128 * K), fB(128 * K), fV((size_t)NV * 128);
    uint32_t s = 12345u + KCH * 7 + NV;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (int i = 0; i < 128 * K; ++i) { fA[i] = bf16_round(rnd()); hA[i] = __float2bfloat16(fA[i]); fB[i] = bf16_round(rnd()); hB[i] = __float2bfloat16(fB[i]); }
    for (size_t i = 0; i < hV.size(); ++i) { fV[i] = bf16_round(rnd()); hV[i] = __float2bfloat16(fV[i]); }
    __nv_bfloat16 *dA, *dB, *dV;
    float *dS, *dP;
    int *derr;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dV, hV.size() * 2));
    CK(cudaMalloc(&dS, 128 * 128 * 4)); CK(cudaMalloc(&dP, 128 * NV * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dV, hV.data(), hV.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dS, 0, 128 * 128 * 4)); CK(cudaMemset(dP, 0, 128 * NV * 4)); CK(cudaMemset(derr, 0, 4));
    CUtensorMap mA = make_map(enc, dA, 128, K, 128), mB = make_map(enc, dB, 128, K, 128), mV = make_map(enc, dV, NV, 128, NV);
    size_t smem = (size_t)KCH * 2 * 16384 + 2 * NV * 128 + 256 + 1024;
    CK(cudaFuncSetAttribute(selftest_kernel<KCH, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_kernel<KCH, NV><<<1, 192, smem>>>(mA, mB, mV, dS, dP, derr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    std::vector<float> S(128 * 128), P((size_t)128 * NV);
    CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost));
    double errS = 0, errP = 0, magS = 0, magP = 0;
    std::vector<float> E(128 * 128);
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)fA[i * K + k] * fB[j * K + k];
            errS = fmax(errS, fabs(acc - S[i * 128 + j]));
            magS = fmax(magS, fabs(acc));
            E[i * 128 + j] = bf16_round(S[i * 128 + j]);
        }
    for (int i = 0; i < 128; ++i)
        for (int c = 0; c < NV; ++c) {
            double acc = 0;
            for (int j = 0; j < 128; ++j) acc += (double)E[i * 128 + j] * fV[(size_t)c * 128 + j];
            errP = fmax(errP, fabs(acc - P[(size_t)i * NV + c]));
            magP = fmax(magP, fabs(acc));
        }
    bool ok = herr == 0 && errS < 1e-4 * fmax(magS, 1.0) && errP < 1e-4 * fmax(magP, 1.0);
    printf("case K=%d NV=%d: timeout_tag=%d  S max|err|=%.3g (max|S|=%.3g)  Phi max|err|=%.3g (max|Phi|=%.3g)  %s\n", K, NV, herr, errS,
           magS, errP, magP, ok ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dV); cudaFree(dS); cudaFree(dP); cudaFree(derr);
    return ok ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += run_case<1, 16>();
    bad += run_case<1, 128>();
    bad += run_case<3, 144>();
    bad += run_case<3, 80>();
    printf(bad ? "SELFTEST FAILED\n" : "SELFTEST PASSED\n");
    return bad ? 1 : 0;
}
