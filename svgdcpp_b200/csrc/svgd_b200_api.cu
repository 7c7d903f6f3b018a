// svgd_b200_api.cu — context, host orchestration and the C ABI of include/svgd_b200.h.
//
// One SVGD::Step (reference SVGD.hpp:373-400) is, on the device:
//   r = |x|^2  ->  [median bandwidth a]  ->  G = grad log p(X_local)  ->  V = G - 2aX  [all-gather V]
//   ->  fused pair interaction + optimizer + clamp -> X_next(local rows)  [all-gather X_next]  -> swap
// There is no CPU fallback anywhere in this file: every numerical result comes from a kernel.
#include "../../include/svgd_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types only: the library is resolved lazily with dlopen (see NcclApi)
#include <cublas_v2.h> // types only: resolved lazily with dlopen as well (see CublasApi)

#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_math.hpp"
#include "kernels_f64.cuh"
#include "kernels_hessian.cuh"
#include "select.cuh"
#ifdef SVGDB_WITH_TC32
#include "kernels_tc32.cuh"
#include "kernels_phi_tc.cuh"
#include "kernels_dist_tc.cuh"
#include "kernels_phi_wide.cuh"
#include "kernels_dist_wide.cuh"
#endif

using namespace svgdb;
using namespace svgdb::host;

namespace {

enum { MODEL_UNSET = 0, MODEL_MVN_SUM = 1, MODEL_HOOK = 2 };
constexpr uint64_t KEY_END = 0x7FF0000000000001ull; // one past the bit pattern of +inf

// NCCL is bound at run time, on first multi-GPU use, instead of through DT_NEEDED: a process that
// already carries an NCCL (PyTorch bundles its own libnccl.so.2) must keep exactly that one, and a
// single-GPU process must not drag one in at all.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
    auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) api.why = std::string("missing NCCL symbol ") + name; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GetErrorString;
    return api;
}

// cuBLAS, bound at run time like NCCL.  Used for ONE plain library GEMM: Y = (X - mu_c) Sigma_c^-1 of the mixture gradient
// (launch_grad_gemm); everything around it -- and every other kernel of the path -- is this library's own code.
struct CublasApi {
    cublasStatus_t (*Create)(cublasHandle_t *) = nullptr;
    cublasStatus_t (*Destroy)(cublasHandle_t) = nullptr;
    cublasStatus_t (*SetStream)(cublasHandle_t, cudaStream_t) = nullptr;
    cublasStatus_t (*Dgemm)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const double *, const double *, int,
                            const double *, int, const double *, double *, int) = nullptr;
    bool ok = false;
    std::string why;
};

CublasApi &cublas()
{
    static CublasApi api = [] { // (function-local static: initialised once, also when several device threads of the facade get here together)
        CublasApi a;
        void *h = nullptr;
        for (const char *name : {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (h) break;
        }
        if (!h) { a.why = std::string("cannot load libcublas.so.12: ") + dlerror(); return a; }
        auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) a.why = std::string("missing cuBLAS symbol ") + name; return p; };
        a.Create = reinterpret_cast<decltype(a.Create)>(sym("cublasCreate_v2"));
        a.Destroy = reinterpret_cast<decltype(a.Destroy)>(sym("cublasDestroy_v2"));
        a.SetStream = reinterpret_cast<decltype(a.SetStream)>(sym("cublasSetStream_v2"));
        a.Dgemm = reinterpret_cast<decltype(a.Dgemm)>(sym("cublasDgemm_v2"));
        a.ok = a.Create && a.Destroy && a.SetStream && a.Dgemm;
        return a;
    }();
    return api;
}

struct HostScratch { // pinned
    unsigned long long below, cand_total, max_below, cand_count; // the first three mirror the device's pass_words
    unsigned long long hist[HIST_BINS];
    MedianResult med;
    StepDecision dec; // mirror of the device-side verdict of an optimistic step
};

} // namespace

struct svgdb_ctx {
    int device = 0;
    int64_t N = 0;
    int d = 0;
    int precision = SVGDB_PRECISION_F64;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;                 // svgdb_step_host: finished rows leave for the host while the pair kernel works on the rest
    cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_up[4] = {nullptr, nullptr, nullptr, nullptr};
    int up_chunks = 0;                                  // svgdb_step_host: the particles are arriving in this many row chunks (copy stream) ...
    int up_ipairs[4] = {0, 0, 0, 0};                    // ... chunk k ends at i-pair up_ipairs[k] (256 rows each); consumed by the first distance pass
    double *stream_out = nullptr;                       // host destination of this rank's updated rows during svgdb_step_host
    bool streamed = false;                              // the last step delivered its rows to stream_out
    cudaStream_t side_stream = nullptr;                 // grad log p runs here, next to the median pass (it only needs X)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_med = nullptr;
    bool med_pending = false; // the last median's result is still on its way to the pinned mirror
    bool grad_pending = false;
    int sm_count = 148;

    // sharding
    int world = 1, rank = 0;
    ncclComm_t comm = nullptr;
    int64_t rows_per_rank = 0, row0 = 0, n_rows = 0, n_pad = 0;

    // device state
    double *X[2] = {nullptr, nullptr};
    int cur = 0;
    double *V = nullptr, *G = nullptr, *r = nullptr, *s1 = nullptr, *s2 = nullptr, *a_dev = nullptr;
    double *lb = nullptr, *ub = nullptr;
    double *phi_dbg = nullptr;

    // model
    int model_kind = MODEL_UNSET;
    int C = 0;
    double *means_dev = nullptr, *prec_dev = nullptr;
    // mixture gradient through a library DGEMM (launch_grad_gemm): SVGDB_GRAD_GEMM = -1 automatic, 0 never, 1 whenever the model is a mixture
    int grad_gemm = -1;
    cublasHandle_t blas = nullptr;
    double *gg_D = nullptr, *gg_Y = nullptr, *gg_ms = nullptr; // [n_rows][d] differences and products, [2][n_rows] running maximum and sum
    int64_t gg_rows = 0;
    svgdb_grad_fn hook = nullptr;
    void *hook_user = nullptr;

    std::vector<double> prec_host, means_host; // host copies of the model parameters (Hessian scale)

    // kernel
    bool kernel_set = false;
    int scale_method = SVGDB_SCALE_MEDIAN;
    double fixed_a = 0.0;
    // ScaleMethod::Hessian (kernels_hessian.cuh): scale matrix A = R^T R, transformed particles / gradients
    std::vector<double> A_host;
    bool hess_const_valid = false; // one Gaussian: the Hessian does not depend on the particles, A and its factors are kept
    double *Hsum_dev = nullptr, *Wsum_dev = nullptr, *R_dev = nullptr, *Rt_dev = nullptr, *Rinv_dev = nullptr, *Y_dev = nullptr, *GH_dev = nullptr;

    // optimizer
    bool opt_set = false;
    OptParams opt{};
    uint64_t counter = 0;
    bool initialized = false;

    // median machinery
    unsigned long long *pass_words = nullptr; // [below, cand_total, max_below] adjacent: one all-reduce, one read-back
    unsigned long long *below = nullptr, *cand_total = nullptr, *max_below = nullptr, *hist = nullptr, *cand = nullptr, *cand_count = nullptr;
    uint64_t capacity = 1ull << 25;
    SelectState *sel = nullptr;
    MedianResult *medres = nullptr;
    HostScratch *hs = nullptr;
    // bracket prediction for the next median: (up to) cubic extrapolation of the last medians of D2
    int n_hist = 0;              // valid entries of med_hist (most recent first)
    double med_hist[MEDIAN_HISTORY] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double pred_err = 0.0;       // worst back-test error of the extrapolation chosen for the current step
    double delta = 0.0;          // relative half-width of the predicted bracket
    double density = 2.0;        // candidates per (pair x unit relative width of D2) seen by the last predicted pass
    double resid[2] = {0.0, 0.0}; // recent relative prediction errors
    double last_pred = 0.0;
    bool have_pred = false;

    // tensor-core path (SVGDB_PRECISION_TC32)
    int64_t n_pad128 = 0;
    int dp = 64;       // particle dimension padded to a multiple of 64 (d <= 64: the two-tile kernels; 64 < d <= 256: the wide kernels)
    bool wide = false; // d > 64
    CUtensorMap mapBW{}, mapBWP{}, mapVW{}; // wide kernels: column operand hi (XB2, [np][dp]) / [hi | lo] (XBD, [np][2 dp]), V^T ([2 dp][np]); boxes of 64 x 64
    float *phi_buf = nullptr;
    double *rt = nullptr, *colsum = nullptr;
    int *tc_err = nullptr;
    long long *tc_trace = nullptr; // SVGDB_TC_TRACE=<file>: timeline of CTA 0 of the pair-interaction kernel
    // persistent pair-interaction kernel (kernels_phi_tc.cuh): fp16 row / column operands, exponent offsets, V^T
    __half *XA2 = nullptr, *XB2 = nullptr, *VT2 = nullptr;
    float *V32 = nullptr; // v~ = g - 2 a (x - mean) in fp32, [n_pad][d]: what the ranks all-gather in this mode
    __half *UA2 = nullptr, *WB2 = nullptr; // exponent-offset K chunks (row / column side)
    CUtensorMap mapB2{}, mapV2{}, mapB8{}, mapV8{};
    uint8_t *VT8 = nullptr; // e5m2 copy of v_lo^T (pair kernel, F8LO variant)
    uint8_t *XB8 = nullptr; // e5m2 copy of the column operand (pair kernel, F8LO variant)
    int phi_f8 = -1;        // SVGDB_PHI_F8: the FAST pair kernel computes its two correction terms (lo_i . y^_j and E . v_lo) in e5m2:
                            // 13 MMAs per unit instead of 17.  -1 = automatic (phi_use_f8), 0 = never, 1 = whenever the FAST variant runs
    __nv_bfloat16 *XBD = nullptr; // column operand [hi | lo] of the persistent distance pass (kernels_dist_tc.cuh)
    CUtensorMap mapBD{};
    int dist_dbg_mode = 0; // svgdb_time_kernel measurement aid
    int dist_gated = -1;   // SVGDB_DIST_GATED=0/1 forces the flat / gated counting epilogue (default: chosen per pass)
    // optimistic steps (tensor-core path, median history available): the bracket verdict is taken on the device and read by the host
    // after the step has been enqueued; a miss is repaired by repeating the step synchronously
    int optimistic = 0;            // SVGDB_OPTIMISTIC=1: on.  Off by default: measured on B200 it does not pay -- on one GPU the
                                   // host round trip hides behind grad log p on the side stream (2.67 vs 2.63 ms per step with it), on eight the
                                   // step is bound by the collectives and the replicated operand preparation (0.727 vs 0.713 ms)
    bool opt_suspended = false;    // a repair or an inspection call is running: synchronous
    bool pending_verify = false;   // the last step's verdict has not been read yet
    double pending_dl = 0.0;       // half-width of that step's bracket (density bookkeeping)
    double miss_boost = 1.0;       // widening factor of the predicted bracket: x4 on a miss, halved on every hit down to 1
    StepDecision *dec_dev = nullptr;
    int *miss_dev = nullptr;
    bool dist_ops_f16 = false;   // the distance operands in memory are the scaled fp16 ones of the two-product variant (split_dist2h_kernel)
    double dist_s2 = 1.0;        // ... scaled by s with s^2 = dist_s2, a power of two
    int dist_f16 = 1;            // SVGDB_DIST_F16=0 (measurement aid): always the three-product bf16 operands
    bool dist_fold_next = false; // the next collecting pass runs over a predicted bracket (set by median_scale): it may fold -lo into the operands
    int dist_fold = 1;           // SVGDB_DIST_FOLD=0 (measurement aid) disables that
    uint64_t collect_hi_ext = 0; // exclusive key bound of what the last persistent distance pass may have collected (>= its hi)
    int phi_tcsum = 1;   // SVGDB_PHI_TCSUM: the e5m2 variant takes the row sums from the tensor core (four N = 16 MMAs per unit) instead of the exp warps
    int phi_no_vlo = -1; // SVGDB_PHI_NO_VLO: the e5m2 variant of the FAST pair kernel leaves E . v_lo out (v carries one fp16 term, like E and the
                         // column particle: 11 MMAs per unit).  -1 = automatic: from 32,768 particles (measured error)
    int phi_poly = 1;    // SVGDB_PHI_POLY=k (0, 1, 2, 4): k of the 16 exponential pairs of a 32-column chunk on the FMA pipe (measured optimum 1)
    int phi_cluster = -1; // 2-CTA clusters with TMA multicast of the column tiles in the pair kernels.  Default (-1): on for the wide kernel (d > 64:
                          // -10 % at d = 256), off for d <= 64 (measured: no gain, 1.83 ms either way -- that kernel is not bound by L2 traffic);
                          // SVGDB_PHI_CLUSTER=0 / 1 forces it
    int phi_max_seg = 0; // SVGDB_PHI_MAX_SEG (measurement aid): column tiles accumulated in TMEM between flushes (0: default per variant)
    int tc32_variant = SVGDB_TC32_AUTO; // svgdb_set_tc32_variant / SVGDB_TC32_VARIANT: arithmetic of the tensor-core pair kernel
    int phi_dbg_mode = 0; // SVGDB_PHI_DBG (development): see Phi2Args::dbg
    int host_chunks = 1;  // SVGDB_HOST_CHUNKS=0 (measurement aid): svgdb_step_host moves the particles in one piece each way

    // measurement
    svgdb_stats stats{};
    bool profiling = false;
    cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // [5], [6]: around the pair-interaction kernel; [7], [8]: around grad log p (side stream)

    std::string err;
};

namespace {

int fail(svgdb_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, SVGDB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

#define NC(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess)                                                                           \
            return fail(ctx, SVGDB_ERR_NCCL, std::string(#call) + ": " + nccl().GetErrorString(r_));        \
    } while (0)

#define TRY(call)                                                                                        \
    do {                                                                                                 \
        int rc_ = (call);                                                                                \
        if (rc_ != SVGDB_OK) return rc_;                                                                 \
    } while (0)

#define KERNEL_CHECK()                                                                                   \
    do {                                                                                                 \
        ++ctx->stats.kernel_launches;                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, SVGDB_ERR_CUDA, std::string("kernel launch at svgd_b200_api.cu:") + std::to_string(__LINE__) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// Gauss-Jordan inverse with partial pivoting (double).  A is d x d; returns false if singular.
bool invert_matrix(const double *A, int d, std::vector<double> &inv)
{
    std::vector<double> w((size_t)d * 2 * d, 0.0);
    for (int r = 0; r < d; ++r) {
        for (int c = 0; c < d; ++c) w[(size_t)r * 2 * d + c] = A[(size_t)r * d + c];
        w[(size_t)r * 2 * d + d + r] = 1.0;
    }
    for (int col = 0; col < d; ++col) {
        int piv = col;
        double best = std::fabs(w[(size_t)col * 2 * d + col]);
        for (int r = col + 1; r < d; ++r) {
            double v = std::fabs(w[(size_t)r * 2 * d + col]);
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > 0.0) || !std::isfinite(best)) return false;
        if (piv != col)
            for (int c = 0; c < 2 * d; ++c) std::swap(w[(size_t)col * 2 * d + c], w[(size_t)piv * 2 * d + c]);
        double pv = w[(size_t)col * 2 * d + col];
        for (int c = 0; c < 2 * d; ++c) w[(size_t)col * 2 * d + c] /= pv;
        for (int r = 0; r < d; ++r) {
            if (r == col) continue;
            double f = w[(size_t)r * 2 * d + col];
            if (f == 0.0) continue;
            for (int c = 0; c < 2 * d; ++c) w[(size_t)r * 2 * d + c] -= f * w[(size_t)col * 2 * d + c];
        }
    }
    inv.assign((size_t)d * d, 0.0);
    for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) inv[(size_t)r * d + c] = w[(size_t)r * 2 * d + d + c];
    // d/dx of -1/2 x^T P x is -1/2 (P + P^T) x: the symmetric part is what the reference's AD sees
    for (int r = 0; r < d; ++r)
        for (int c = r + 1; c < d; ++c) {
            double s = 0.5 * (inv[(size_t)r * d + c] + inv[(size_t)c * d + r]);
            inv[(size_t)r * d + c] = inv[(size_t)c * d + r] = s;
        }
    return true;
}

int free_sharded(svgdb_ctx *ctx)
{
    cudaFree(ctx->X[0]); cudaFree(ctx->X[1]); cudaFree(ctx->V); cudaFree(ctx->G); cudaFree(ctx->r);
    cudaFree(ctx->s1); cudaFree(ctx->s2); cudaFree(ctx->phi_dbg);
    ctx->X[0] = ctx->X[1] = ctx->V = ctx->G = ctx->r = ctx->s1 = ctx->s2 = ctx->phi_dbg = nullptr;
    cudaFree(ctx->phi_buf);
    cudaFree(ctx->rt); cudaFree(ctx->colsum); cudaFree(ctx->tc_err); cudaFree(ctx->tc_trace);
    cudaFree(ctx->XBD);
    ctx->XBD = nullptr;
    cudaFree(ctx->XB8);
    ctx->XB8 = nullptr;
    cudaFree(ctx->VT8);
    ctx->VT8 = nullptr;
    cudaFree(ctx->V32);
    ctx->V32 = nullptr;
    cudaFree(ctx->XA2); cudaFree(ctx->XB2); cudaFree(ctx->VT2); cudaFree(ctx->UA2); cudaFree(ctx->WB2);
    ctx->XA2 = ctx->XB2 = ctx->VT2 = nullptr;
    ctx->UA2 = ctx->WB2 = nullptr;
    ctx->tc_trace = nullptr;
    ctx->phi_buf = nullptr;
    ctx->rt = ctx->colsum = nullptr;
    ctx->tc_err = nullptr;
    return SVGDB_OK;
}

#ifdef SVGDB_WITH_TC32
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// rows x cols bf16, cols contiguous; box = 64 cols (128 B, SWIZZLE_128B) x box_rows; out-of-bounds reads give zeros
int make_bf16_map(svgdb_ctx *ctx, CUtensorMap *m, void *base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    static TmapEncodeFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn) return fail(ctx, SVGDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        enc = (TmapEncodeFn)fn;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SVGDB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return SVGDB_OK;
}

// rows x 64 bytes, contiguous; box = 64 bytes (SWIZZLE_64B) x box_rows
int make_u8_map_sw64(svgdb_ctx *ctx, CUtensorMap *m, void *base, uint64_t rows, uint32_t box_rows)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) return fail(ctx, SVGDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((TmapEncodeFn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SVGDB_ERR_CUDA, "cuTensorMapEncodeTiled (u8, 64B swizzle) failed with CUresult " + std::to_string((int)r));
    return SVGDB_OK;
}

// rows x cols bytes, cols contiguous; box = 64 bytes (SWIZZLE_64B) x box_rows
int make_u8_map_rows(svgdb_ctx *ctx, CUtensorMap *m, void *base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) return fail(ctx, SVGDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((TmapEncodeFn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SVGDB_ERR_CUDA, "cuTensorMapEncodeTiled (u8 rows, 64B swizzle) failed with CUresult " + std::to_string((int)r));
    return SVGDB_OK;
}

// e5m2 correction terms in the FAST pair kernel.  The rounding error of lo_i (2^-3 of a 2^-11 term) is the same for every pair of
// row i, so it does not average out over j; it does average over the d coordinates of lo_i . y^_j.  Automatic: only where that
// averaging keeps the row error where it was (measured at N = 65,536 on sampled rows incl. the outermost particle: 5.1e-4 against 2.2e-4
// at d = 2, 3.5e-5 / 3.1e-5 at d = 8, no difference from d = 16 up) AND where it pays: the pair pass gains 4 - 6 % at d = 64 and loses
// 2 % at d <= 32, where the zero-padded operands leave the tensor pipe under its power limit and the exp warps' extra work shows.
bool tc32_precise(const svgdb_ctx *ctx);
bool phi_use_f8(const svgdb_ctx *ctx)
{
    return svgdb::host::rule_phi_lean(tc32_precise(ctx), ctx->wide, ctx->phi_f8, ctx->N, ctx->d);
}

int alloc_tc32(svgdb_ctx *ctx)
{
    using namespace svgdb::tc;
    ctx->n_pad128 = (ctx->N + 127) / 128 * 128;
    ctx->wide = ctx->d > TC_D;
    ctx->dp = ctx->wide ? (ctx->d + 63) / 64 * 64 : 64;
    const size_t np = (size_t)ctx->n_pad128, dp = (size_t)ctx->dp;
    const size_t phi_ld = ctx->wide ? dp + 16 : (size_t)TC_PHI_LD;
    CU(cudaMalloc(&ctx->rt, np * 8));
    CU(cudaMalloc(&ctx->colsum, 256 * 8));
    CU(cudaMalloc(&ctx->tc_err, 4));
    if (std::getenv("SVGDB_TC_TRACE")) {
        CU(cudaMalloc(&ctx->tc_trace, 3 * 64 * 8 * 8));
        CU(cudaMemsetAsync(ctx->tc_trace, 0, 3 * 64 * 8 * 8, ctx->stream));
    }
    // rows of a tile may reach past the last rank-local row: keep a tile of slack
    CU(cudaMalloc(&ctx->phi_buf, (np + 256) * phi_ld * 4));
    CU(cudaMemsetAsync(ctx->tc_err, 0, 4, ctx->stream));
    // operands (16-bit elements: the bf16 tensor-map type moves fp16 bit patterns unchanged)
    CU(cudaMalloc(&ctx->XA2, (np + 256) * 2 * dp * 2));             // row operand [hi | lo]
    CU(cudaMalloc(&ctx->XB2, np * dp * 2));                         // column operand hi (pair kernel, fast variant)
    CU(cudaMalloc(&ctx->VT2, (size_t)2 * dp * np * 2));             // [v_hi ; v_lo]^T
    CU(cudaMalloc(&ctx->UA2, (np + 256) * 16 * 2));
    CU(cudaMalloc(&ctx->WB2, np / 128 * P2_W_BYTES));
    CU(cudaMemsetAsync(ctx->XA2, 0, (np + 256) * 2 * dp * 2, ctx->stream));
    CU(cudaMemsetAsync(ctx->UA2, 0, (np + 256) * 16 * 2, ctx->stream));
    CU(cudaMalloc(&ctx->V32, (size_t)ctx->n_pad * ctx->d * sizeof(float)));
    CU(cudaMemsetAsync(ctx->V32, 0, (size_t)ctx->n_pad * ctx->d * sizeof(float), ctx->stream));
    CU(cudaMalloc(&ctx->XBD, np * 2 * dp * 2));                     // column operand [hi | lo] (distance pass; pair kernel, precise variant)
    if (ctx->wide) {
        TRY(make_bf16_map(ctx, &ctx->mapBW, ctx->XB2, np, dp, 64));
        TRY(make_bf16_map(ctx, &ctx->mapBWP, ctx->XBD, np, 2 * dp, 64));
        TRY(make_bf16_map(ctx, &ctx->mapVW, ctx->VT2, 2 * dp, np, 64));
    } else {
        TRY(make_bf16_map(ctx, &ctx->mapBD, ctx->XBD, np, 128, 128));
        TRY(make_bf16_map(ctx, &ctx->mapB2, ctx->XB2, np, 64, 128));
        TRY(make_bf16_map(ctx, &ctx->mapV2, ctx->VT2, 128, np, 64));
        CU(cudaMalloc(&ctx->XB8, np * 64));
        TRY(make_u8_map_sw64(ctx, &ctx->mapB8, ctx->XB8, np, 128));
        CU(cudaMalloc(&ctx->VT8, np * 64));
        TRY(make_u8_map_rows(ctx, &ctx->mapV8, ctx->VT8, 64, np, 64));
    }
    if (const char *e = std::getenv("SVGDB_DIST_GATED")) ctx->dist_gated = std::atoi(e) != 0;
    if (const char *e = std::getenv("SVGDB_DIST_FOLD")) ctx->dist_fold = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_DIST_F16")) ctx->dist_f16 = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_OPTIMISTIC")) ctx->optimistic = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_POLY")) ctx->phi_poly = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_NO_VLO")) ctx->phi_no_vlo = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_GRAD_GEMM")) ctx->grad_gemm = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_TCSUM")) ctx->phi_tcsum = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_TC32_VARIANT")) ctx->tc32_variant = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_DBG")) ctx->phi_dbg_mode = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_MAX_SEG")) ctx->phi_max_seg = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_CLUSTER")) ctx->phi_cluster = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_PHI_F8")) ctx->phi_f8 = std::atoi(e);
    if (const char *e = std::getenv("SVGDB_HOST_CHUNKS")) ctx->host_chunks = std::atoi(e);
    return SVGDB_OK;
}
#endif

int alloc_sharded(svgdb_ctx *ctx)
{
    free_sharded(ctx);
    ctx->rows_per_rank = (ctx->N + ctx->world - 1) / ctx->world;
    ctx->n_pad = ctx->rows_per_rank * ctx->world;
    ctx->row0 = ctx->rows_per_rank * ctx->rank;
    ctx->n_rows = std::max<int64_t>(0, std::min<int64_t>(ctx->rows_per_rank, ctx->N - ctx->row0));
    size_t full = (size_t)ctx->n_pad * ctx->d * sizeof(double);
    size_t local = (size_t)ctx->rows_per_rank * ctx->d * sizeof(double);
    CU(cudaMalloc(&ctx->X[0], full));
    CU(cudaMalloc(&ctx->X[1], full));
    CU(cudaMalloc(&ctx->V, full));
    CU(cudaMalloc(&ctx->r, (size_t)ctx->n_pad * sizeof(double)));
    CU(cudaMalloc(&ctx->G, local));
    CU(cudaMalloc(&ctx->s1, local));
    CU(cudaMalloc(&ctx->s2, local));
    CU(cudaMalloc(&ctx->phi_dbg, local));
    CU(cudaMemsetAsync(ctx->X[0], 0, full, ctx->stream));
    CU(cudaMemsetAsync(ctx->X[1], 0, full, ctx->stream));
    CU(cudaMemsetAsync(ctx->V, 0, full, ctx->stream));
    CU(cudaMemsetAsync(ctx->s1, 0, local, ctx->stream));
    CU(cudaMemsetAsync(ctx->s2, 0, local, ctx->stream));
    ctx->cur = 0;
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) TRY(alloc_tc32(ctx));
#endif
    return SVGDB_OK;
}

// ---- collectives (no-ops at world == 1) ------------------------------------------------------
int allreduce_u64(svgdb_ctx *ctx, unsigned long long *buf, size_t count, ncclRedOp_t op)
{
    if (ctx->world == 1) return SVGDB_OK;
    NC(nccl().AllReduce(buf, buf, count, ncclUint64, op, ctx->comm, ctx->stream));
    return SVGDB_OK;
}

int allgather_rows(svgdb_ctx *ctx, double *buf, int64_t elems_per_row)
{
    if (ctx->world == 1) return SVGDB_OK;
    size_t chunk = (size_t)ctx->rows_per_rank * elems_per_row;
    NC(nccl().AllGather(buf + (size_t)ctx->rank * chunk, buf, chunk, ncclDouble, ctx->comm, ctx->stream));
    return SVGDB_OK;
}

// ---- phases ------------------------------------------------------------------------------------
int launch_rownorm(svgdb_ctx *ctx)
{
    const int wpb = 8;
    int64_t blocks = (ctx->N + wpb - 1) / wpb;
    rownorm_f64_kernel<<<(unsigned)blocks, wpb * 32, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->N, ctx->d, ctx->r);
    KERNEL_CHECK();
    return SVGDB_OK;
}

size_t dist_smem_bytes(int d, bool hist)
{
    int kc = d < 64 ? ((d + 3) & ~3) : 64;
    int ldx = ((kc + 15) & ~15) + 4;
    return (size_t)(2 * 64 * ldx + 128) * sizeof(double) + (hist ? HIST_BINS * sizeof(unsigned int) : 0);
}

#ifdef SVGDB_WITH_TC32
int launch_dist_pass_tc32(svgdb_ctx *ctx, int mode, uint64_t lo, uint64_t hi, int shift);
#endif
int reduce_pass_words(svgdb_ctx *ctx, int mode);
int finish_median(svgdb_ctx *ctx);
int settle_step(svgdb_ctx *ctx);
void prof_mark(svgdb_ctx *ctx, int i);
int kick_grad(svgdb_ctx *ctx);

int launch_dist_pass(svgdb_ctx *ctx, int mode, uint64_t lo, uint64_t hi, int shift)
{
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) return launch_dist_pass_tc32(ctx, mode, lo, hi, shift);
#endif
    DistArgs a{};
    a.X = ctx->X[ctx->cur];
    a.r = ctx->r;
    a.n_total = ctx->N;
    a.d = ctx->d;
    // the distance pass only needs X, which every rank holds: keep the symmetric (upper-triangle) enumeration at
    // any world size and deal the tile pairs cyclically to the ranks
    a.row0 = 0;
    a.n_rows = ctx->N;
    a.sym = 1;
    a.n_tiles_j = (ctx->N + 63) / 64;
    a.n_tiles_i = a.n_tiles_j;
    a.n_work = a.n_tiles_j * (a.n_tiles_j + 1) / 2;
    a.work_offset = ctx->rank;
    a.work_stride = ctx->world;
    a.lo = lo;
    a.hi = hi;
    a.shift = shift;
    a.below = ctx->below;
    a.max_below = ctx->max_below;
    a.hist = ctx->hist;
    a.cand = ctx->cand;
    a.cand_count = ctx->cand_count;
    a.capacity = ctx->capacity;
    CU(cudaMemsetAsync(ctx->below, 0, sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(ctx->max_below, 0, sizeof(unsigned long long), ctx->stream));
    if (a.n_work > 0) {
        unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((a.n_work + ctx->world - 1) / ctx->world, (int64_t)ctx->sm_count * 4));
        if (mode == MODE_HIST) {
            CU(cudaMemsetAsync(ctx->hist, 0, HIST_BINS * sizeof(unsigned long long), ctx->stream));
            dist_pass_f64_kernel<MODE_HIST><<<grid, 128, dist_smem_bytes(ctx->d, true), ctx->stream>>>(a);
        } else {
            CU(cudaMemsetAsync(ctx->cand_count, 0, sizeof(unsigned long long), ctx->stream));
            dist_pass_f64_kernel<MODE_COLLECT><<<grid, 128, dist_smem_bytes(ctx->d, false), ctx->stream>>>(a);
        }
        KERNEL_CHECK();
    }
    ++ctx->stats.median_passes;
    TRY(kick_grad(ctx));
    TRY(reduce_pass_words(ctx, mode));
    return SVGDB_OK;
}

// Global counts of a distance pass: below and the in-bracket total are summed over the ranks in ONE all-reduce (the words
// are adjacent), the largest value below the bracket (FP64 path only: the tensor-core passes do not track it) by a max.
int reduce_pass_words(svgdb_ctx *ctx, int mode)
{
    if (mode != MODE_HIST) CU(cudaMemcpyAsync(ctx->cand_total, ctx->cand_count, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    TRY(allreduce_u64(ctx, ctx->below, mode == MODE_HIST ? 1 : 2, ncclSum));
    if (ctx->precision != SVGDB_PRECISION_TC32) TRY(allreduce_u64(ctx, ctx->max_below, 1, ncclMax));
    if (mode == MODE_HIST) TRY(allreduce_u64(ctx, ctx->hist, HIST_BINS, ncclSum));
    return SVGDB_OK;
}

// Reads below / max_below / (hist | cand_count) back to the pinned mirror.  In the multi-rank case
// `cand_count` stays local; hs->below etc. are global.  mid_total gets the global in-bracket count.
int read_pass_results(svgdb_ctx *ctx, int mode, uint64_t *mid_total)
{
    CU(cudaMemcpyAsync(&ctx->hs->below, ctx->pass_words, 3 * 8, cudaMemcpyDeviceToHost, ctx->stream)); // below, cand_total, max_below
    if (mode == MODE_HIST) {
        CU(cudaMemcpyAsync(ctx->hs->hist, ctx->hist, HIST_BINS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    } else {
        CU(cudaMemcpyAsync(&ctx->hs->cand_count, ctx->cand_count, 8, cudaMemcpyDeviceToHost, ctx->stream)); // stays rank-local
        CU(cudaStreamSynchronize(ctx->stream));
        if (mid_total) *mid_total = ctx->hs->cand_total; // global in-bracket count
    }
    return SVGDB_OK;
}

// dec != nullptr: rank and candidate count come from the device-side verdict (optimistic step); the grid covers the whole buffer
// (the kernels stride over whatever the device says the count is; `expected` only sizes the grid)
int run_select(svgdb_ctx *ctx, uint64_t lo, uint64_t hi, uint64_t kk, bool even, double log_n, const StepDecision *dec = nullptr,
               uint64_t expected = 0)
{
    uint64_t m_local = dec ? std::min<uint64_t>(std::max<uint64_t>(expected, 2048), ctx->capacity) : std::min<uint64_t>(ctx->hs->cand_count, ctx->capacity);
    // The candidates are selected on key - base, base = the bracket's lower end: its span, not the raw bit patterns, decides
    // how many bits matter (a bracket of relative width 2e-4 spans ~2^11 fp32 values: one 12-bit pass).  Keys of the
    // tensor-core path are fp32 distances widened to double: their low 29 bits are zero, and so are those of key - base
    // once base is rounded down to a multiple of 2^29.
    const int bottom = ctx->precision == SVGDB_PRECISION_TC32 ? 29 : 0;
    const uint64_t base = bottom ? (lo & ~((1ull << bottom) - 1ull)) : lo;
    const uint64_t span = (hi - 1) - base;
    int top = 0; // number of low bits of key - base that may be set
    while (top < 64 && (span >> top) != 0ull) ++top;
    if (top < bottom + 1) top = bottom + 1;
    const uint64_t low_mask = top >= 64 ? ~0ull : ((1ull << top) - 1ull);
    select_init_kernel<<<1, 256, 0, ctx->stream>>>(ctx->sel, base, 0ull, ~low_mask, kk, dec);
    KERNEL_CHECK();
    unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((m_local + 2047) / 2048, (uint64_t)ctx->sm_count * 8));
    // digits of up to SELECT_MAX_BITS bits from the top differing bit down to `bottom`
    int hi_bit = top;
    while (hi_bit > bottom) {
        const int bits = std::min(SELECT_MAX_BITS, hi_bit - bottom), shift = hi_bit - bits;
        select_hist_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->cand, m_local, shift, bits, ctx->sel, dec);
        KERNEL_CHECK();
        TRY(allreduce_u64(ctx, ctx->sel->hist, (size_t)1 << bits, ncclSum));
        select_pick_kernel<<<1, 256, 0, ctx->stream>>>(ctx->sel, shift, bits, shift == bottom ? 1 : 0);
        KERNEL_CHECK();
        hi_bit = shift;
    }
    if (even) {
        // multi-GPU: each rank decides need_scan from the same all-reduced histogram, the candidates are rank-local
        select_max_less_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->cand, m_local, ctx->sel, dec);
        KERNEL_CHECK();
        TRY(allreduce_u64(ctx, &ctx->sel->max_less, 1, ncclMax));
    }
    median_finalize_kernel<<<1, 32, 0, ctx->stream>>>(ctx->sel, kk, even ? 1 : 0, ctx->max_below, 0, 0ull, log_n,
                                                      ctx->medres, ctx->a_dev, dec);
    KERNEL_CHECK();
    return SVGDB_OK;
}

// Largest relative half-width whose bracket is expected to fit the candidate buffer with a 2.5x margin, from the
// density of distances around the median observed by the last predicted pass (high-dimensional particle sets have
// narrow distance distributions: the same relative width holds many more pairs).
double bracket_delta_max(const svgdb_ctx *ctx, double total)
{
    return std::min(0.0625, 0.4 * (double)ctx->capacity / (2.0 * std::max(ctx->density, 0.25) * total));
}

// GaussianRBFKernel::ComputeScale, Median branch (Kernel/GaussianRBFKernel.hpp:168-188), exact.
int median_scale(svgdb_ctx *ctx)
{
    TRY(finish_median(ctx));
    const long double totald = (long double)ctx->N * (long double)ctx->N;
    if (totald > 1.8e19L) return fail(ctx, SVGDB_ERR_INVALID, "n^2 overflows 64 bits");
    const uint64_t total = (uint64_t)ctx->N * (uint64_t)ctx->N;
    const bool even = (total % 2ull) == 0ull;
    const uint64_t k_hi = total / 2ull; // 0-based rank of the upper middle (the middle when odd)
    const double log_n = std::log((double)ctx->N);

    uint64_t lo = 0, hi = KEY_END, below_known = 0, in_range = total;
    bool collected = false;
    uint64_t mid = 0;

    // 1) predicted bracket (one pass when it holds): the median of D2 moves smoothly from step to step, so it is
    //    extrapolated (up to cubically) from the last medians; the bracket half-width follows the recent
    //    extrapolation error and is capped by what the candidate buffer can hold.  The prediction is only a hint:
    //    the exact counts returned by the pass decide whether it held.
    double delta_max = bracket_delta_max(ctx, (double)total);
    double predicted = 0.0;
    if (ctx->n_hist > 0 && total > 65536ull) { // (tiny problems: one pass collecting everything is cheaper than any logic)
        const double *m = ctx->med_hist;
        // the extrapolation that fits the recent medians best (host_math.hpp): cubic on Adam's smooth trajectories (measured at the
        // headline shape: relative error 9e-6), the parity-aware ones while AdaGrad's early steps overshoot with period two
        int kind = 0;
        double back_err = INFINITY;
        predicted = median_predict_best(m, ctx->n_hist, &kind, &back_err);
        if (!(predicted > 0.0) || !std::isfinite(predicted)) predicted = m[0];
        ctx->pred_err = std::isfinite(back_err) ? back_err : 0.0;
        // half-width: 4x the chosen extrapolation's worst back-test error over the last two medians (before any back-test is
        // possible: the width kept from the previous steps), never below 2e-5
        const double want = std::isfinite(back_err) ? 4.0 * back_err : ctx->delta;
        const double dl = std::min(delta_max, std::max(want, 2e-5) * ctx->miss_boost);
        uint64_t klo = key_of(std::max(predicted * (1.0 - dl), 0.0)), khi = key_of(predicted * (1.0 + dl)) + 1;
        ctx->dist_fold_next = true;
        TRY(launch_dist_pass(ctx, MODE_COLLECT, klo, khi, 0));
        const bool need_both_dev = even && ctx->precision == SVGDB_PRECISION_TC32;
        const bool optimistic = ctx->optimistic != 0;
        if (optimistic && !ctx->opt_suspended && ctx->precision == SVGDB_PRECISION_TC32) {
            // No host round trip: the verdict on the bracket is taken on the device, the select and everything after it are enqueued
            // at once, and the host reads the verdict when the step has been enqueued (settle_step).
            median_decide_kernel<<<1, 32, 0, ctx->stream>>>(ctx->pass_words, ctx->cand_count, k_hi, even ? 1 : 0, need_both_dev ? 1 : 0, ctx->capacity,
                                                            ctx->dec_dev, ctx->miss_dev);
            KERNEL_CHECK();
            uint64_t hi_sel = khi;
#ifdef SVGDB_WITH_TC32
            if (ctx->collect_hi_ext > hi_sel) hi_sel = ctx->collect_hi_ext;
#endif
            const double expect = 4.0 * ctx->density * 2.0 * dl * (double)total / (double)ctx->world; // in-bracket pairs this rank should see, with a margin
            TRY(run_select(ctx, klo, hi_sel, 0, even, log_n, ctx->dec_dev, (uint64_t)std::min(expect, 1e18)));
            CU(cudaMemcpyAsync(&ctx->hs->dec, ctx->dec_dev, sizeof(StepDecision), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(&ctx->hs->med, ctx->medres, sizeof(MedianResult), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaEventRecord(ctx->ev_med, ctx->stream));
            ctx->med_pending = true;
            ctx->pending_verify = true;
            ctx->pending_dl = dl;
            ctx->last_pred = predicted;
            return SVGDB_OK;
        }
        TRY(read_pass_results(ctx, MODE_COLLECT, &mid));
        uint64_t b = ctx->hs->below;
        ctx->density = std::max(0.25, (double)mid / (2.0 * dl * (double)total)); // exact in-bracket count, even on overflow
        delta_max = bracket_delta_max(ctx, (double)total);
        // the tensor-core pass does not track the largest value below the bracket, so for an even count the lower
        // middle element must be a candidate as well (b < k_hi)
        const bool need_both = even && ctx->precision == SVGDB_PRECISION_TC32;
        bool hit = (need_both ? b < k_hi : b <= k_hi) && (k_hi < b + mid) && (mid <= ctx->capacity) && (!even || k_hi >= 1);
        if (hit) {
            lo = klo; hi = khi; below_known = b; collected = true;
            ++ctx->stats.median_bracket_hits;
            ctx->miss_boost = std::max(1.0, 0.5 * ctx->miss_boost);
        } else {
            ctx->miss_boost = std::min(64.0, ctx->miss_boost * 4.0);
            ctx->delta = std::min(delta_max, std::max(ctx->delta, 2e-5) * 4.0);
        }
    }
    ctx->last_pred = predicted;

    // 2) radix narrowing on the key bits until the bracket fits the candidate buffer
    if (!collected) {
        // The collecting pass may see more than the histogram promised: the tensor-core pass collects a superset [lo, hi') of the
        // bracket, and the widening below moves lo downwards.  If that overflows the candidate buffer while the bracket is still
        // wider than one key, narrow again instead of trusting the overflowing list (a few rounds at most, then an error).
        for (int round = 0;; ++round) {
            while (in_range > ctx->capacity && hi - lo > 1) {
                uint64_t span = hi - lo - 1;
                int bits = 0;
                while (bits < 64 && (span >> bits) != 0ull) ++bits;
                int shift = std::max(0, bits - 12);
                TRY(launch_dist_pass(ctx, MODE_HIST, lo, hi, shift));
                TRY(read_pass_results(ctx, MODE_HIST, nullptr));
                uint64_t cum = ctx->hs->below;
                int b = 0;
                for (; b < HIST_BINS; ++b) {
                    if (k_hi < cum + ctx->hs->hist[b]) break;
                    cum += ctx->hs->hist[b];
                }
                if (b == HIST_BINS) return fail(ctx, SVGDB_ERR_NUMERIC, "median select: rank not found (non-finite particles?)");
                uint64_t nlo = lo + ((uint64_t)b << shift);
                uint64_t nhi = std::min<uint64_t>(hi, nlo + (1ull << shift));
                lo = nlo; hi = nhi; below_known = cum; in_range = ctx->hs->hist[b];
            }
            bool pred_outside = false;
            for (int attempt = 0;; ++attempt) {
                TRY(launch_dist_pass(ctx, MODE_COLLECT, lo, hi, 0));
                TRY(read_pass_results(ctx, MODE_COLLECT, &mid));
                below_known = ctx->hs->below;
                if (!(below_known <= k_hi && k_hi < below_known + mid))
                    return fail(ctx, SVGDB_ERR_NUMERIC, "median select: bracket lost the rank (non-finite particles?)");
                // tensor-core pass, even count, lower middle element just below the bracket: widen downwards and recollect
                pred_outside = even && ctx->precision == SVGDB_PRECISION_TC32 && below_known == k_hi && lo > 0 && mid <= ctx->capacity;
                if (!pred_outside || attempt >= 40) break;
                const uint64_t step = std::max<uint64_t>(hi - lo, 1ull) << std::min(attempt, 20);
                lo = lo > step ? lo - step : 0;
            }
            if (pred_outside)
                return fail(ctx, SVGDB_ERR_NUMERIC, "median select: the lower middle distance was not found below the bracket (non-finite particles?)");
            if (mid <= ctx->capacity || hi - lo == 1) break;
            if (round >= 6)
                return fail(ctx, SVGDB_ERR_NUMERIC, "median select: the candidate buffer cannot hold the narrowed bracket (SVGDB_CAND_CAPACITY too small for this particle set)");
            in_range = mid; // > capacity: forces at least one more histogram pass over [lo, hi)
        }
    }

    const uint64_t kk = k_hi - below_known;
    if (mid > ctx->capacity) {
        // here hi - lo == 1 (guarded above; the predicted-bracket path requires mid <= capacity): more than `capacity` identical distances
        median_finalize_kernel<<<1, 32, 0, ctx->stream>>>(ctx->sel, kk, even ? 1 : 0, ctx->max_below, 1, lo, log_n,
                                                          ctx->medres, ctx->a_dev);
        KERNEL_CHECK();
    } else {
#ifdef SVGDB_WITH_TC32
        // the persistent tensor-core pass collects a contiguous range [lo, hi') with hi' slightly past hi
        if (ctx->precision == SVGDB_PRECISION_TC32 && ctx->collect_hi_ext > hi) hi = ctx->collect_hi_ext;
#endif
        TRY(run_select(ctx, lo, hi, kk, even, log_n));
    }
    // The result (needed by the host only for the NEXT bracket prediction and for statistics) is read back without
    // stopping here: finish_median() picks it up before the next prediction, by which time it has long arrived.
    CU(cudaMemcpyAsync(&ctx->hs->med, ctx->medres, sizeof(MedianResult), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev_med, ctx->stream));
    ctx->med_pending = true;
    return SVGDB_OK;
}

// Completes the bookkeeping of the last median_scale(): bracket prediction state and stats.last_scale.
int finish_median(svgdb_ctx *ctx)
{
    if (!ctx->med_pending) return SVGDB_OK;
    ctx->med_pending = false;
    CU(cudaEventSynchronize(ctx->ev_med));
    if (ctx->pending_verify) { // an optimistic step: its verdict arrived with the result
        if (ctx->hs->dec.hit == 0) return SVGDB_OK; // a miss: the result is meaningless; settle_step() repeats the step
        const double total = (double)ctx->N * (double)ctx->N;
        ctx->density = std::max(0.25, (double)ctx->hs->dec.mid / (2.0 * ctx->pending_dl * total));
        ++ctx->stats.median_bracket_hits;
        ctx->miss_boost = std::max(1.0, 0.5 * ctx->miss_boost);
    }
    const double delta_max = bracket_delta_max(ctx, (double)ctx->N * (double)ctx->N);
    ctx->stats.last_scale = ctx->hs->med.scale;
    {
        const double m_now = 0.5 * (ctx->hs->med.d2_lo + ctx->hs->med.d2_hi);
        bool jump = false;
        if (ctx->last_pred > 0.0 && m_now > 0.0) {
            ctx->resid[1] = ctx->resid[0];
            ctx->resid[0] = std::fabs(m_now - ctx->last_pred) / m_now;
            // 4x the worse of the two most recent extrapolation errors, never below 2e-5
            ctx->delta = std::min(delta_max, std::max(2e-5, 4.0 * std::max(ctx->resid[0], ctx->resid[1])));
            // An extrapolation that is off by more than 5 % did not meet a smooth trajectory but a new particle set (the host
            // replaced the particles): the old medians say nothing about the next one, start the history over.
            jump = ctx->n_hist >= 2 && ctx->resid[0] > 0.05;
        } else {
            ctx->delta = std::min(delta_max, 1e-3);
        }
        if (jump) {
            ctx->n_hist = 0;
            ctx->resid[0] = ctx->resid[1] = 0.0;
            ctx->delta = std::min(delta_max, 1e-3);
        }
        for (int k = MEDIAN_HISTORY - 1; k > 0; --k) ctx->med_hist[k] = ctx->med_hist[k - 1];
        ctx->med_hist[0] = m_now;
        ctx->n_hist = std::min(MEDIAN_HISTORY, ctx->n_hist + 1);
        if (!(m_now > 0.0) || !std::isfinite(m_now)) ctx->n_hist = 0;
    }
    return SVGDB_OK;
}

int compute_scale_dev(svgdb_ctx *ctx)
{
    if (ctx->scale_method == SVGDB_SCALE_MEDIAN) return median_scale(ctx);
    if (ctx->scale_method == SVGDB_SCALE_FIXED) {
        CU(cudaMemcpyAsync(ctx->a_dev, &ctx->fixed_a, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.last_scale = ctx->fixed_a;
        return SVGDB_OK;
    }
    return fail(ctx, SVGDB_ERR_INVALID, "internal: the Hessian scale is computed by hessian_scale_dev");
}

// grad log p of a mixture as C library GEMMs + this library's streaming kernels: per component D = X - mu_c (grad_diff_kernel),
// Y = D Sigma_c^-1 (cublasDgemm; Sigma_c^-1 symmetric), then the same online log-sum-exp update as mvn_sum_grad_f64_kernel
// (grad_accumulate_kernel: q = D . Y, h = -q / 2, G <- G e^(m - m') - e^(h - m') Y), finally G /= sum.  Same arithmetic as the one-kernel
// form up to the summation order inside the GEMM.  The one-kernel form re-reads Sigma_c^-1 from L2 for every 16 particles
// (137 GB at config 4) and runs at 18 % of the FP64 FMA rate there: 85 of the 389 ms of a step.
int launch_grad_gemm(svgdb_ctx *ctx, cudaStream_t stream)
{
    CublasApi &api = cublas();
    if (!api.ok) return fail(ctx, SVGDB_ERR_CUDA, "mixture gradient: cuBLAS is not available (" + api.why + "); SVGDB_GRAD_GEMM=0 selects the one-kernel form");
    const int d = ctx->d;
    const int64_t rows = ctx->n_rows;
    if (rows > 0x7fffffff) return fail(ctx, SVGDB_ERR_INVALID, "mixture gradient: more than 2^31 rows per rank");
    if (!ctx->blas && api.Create(&ctx->blas) != CUBLAS_STATUS_SUCCESS) return fail(ctx, SVGDB_ERR_CUDA, "cublasCreate failed");
    if (ctx->gg_rows < rows) {
        cudaFree(ctx->gg_D); cudaFree(ctx->gg_Y); cudaFree(ctx->gg_ms);
        ctx->gg_D = ctx->gg_Y = ctx->gg_ms = nullptr;
        ctx->gg_rows = 0;
        CU(cudaMalloc(&ctx->gg_D, (size_t)rows * d * sizeof(double)));
        CU(cudaMalloc(&ctx->gg_Y, (size_t)rows * d * sizeof(double)));
        CU(cudaMalloc(&ctx->gg_ms, (size_t)2 * rows * sizeof(double)));
        ctx->gg_rows = rows;
    }
    if (api.SetStream(ctx->blas, stream) != CUBLAS_STATUS_SUCCESS) return fail(ctx, SVGDB_ERR_CUDA, "cublasSetStream failed");
    const double one = 1.0, zero = 0.0;
    const int64_t cnt = rows * d;
    const unsigned eb = (unsigned)std::min<int64_t>((cnt + 255) / 256, 148 * 16);
    const unsigned wb = (unsigned)((rows + 7) / 8); // one warp per particle, 8 per block
    for (int c = 0; c < ctx->C; ++c) {
        grad_diff_kernel<<<eb, 256, 0, stream>>>(ctx->X[ctx->cur] + ctx->row0 * d, cnt, d, ctx->means_dev + (size_t)c * d, ctx->gg_D);
        KERNEL_CHECK();
        // row-major Y [rows][d] = D [rows][d] . P [d][d]  <=>  column-major Y^T = P^T D^T (P symmetric)
        if (api.Dgemm(ctx->blas, CUBLAS_OP_N, CUBLAS_OP_N, d, (int)rows, d, &one, ctx->prec_dev + (size_t)c * d * d, d, ctx->gg_D, d, &zero,
                      ctx->gg_Y, d) != CUBLAS_STATUS_SUCCESS)
            return fail(ctx, SVGDB_ERR_CUDA, "cublasDgemm failed");
        grad_accumulate_kernel<<<wb, 256, 0, stream>>>(ctx->gg_D, ctx->gg_Y, rows, d, c, c == ctx->C - 1 ? 1 : 0, ctx->gg_ms, ctx->gg_ms + rows, ctx->G);
        KERNEL_CHECK();
    }
    return SVGDB_OK;
}

// Automatic (rule_grad_gemm, host_math.hpp) at sizes where a GEMM is not all launch latency.  Measured: config 4 (d = 256, C = 16) 37.7 against
// 83.8 ms, config 3 (d = 64, one Gaussian, N = 65,536) 0.072 against 0.136 ms -- 2.14 against 2.21 ms per step, the gradient shares the SMs
// with the distance pass.
bool grad_use_gemm(const svgdb_ctx *ctx)
{
    return svgdb::host::rule_grad_gemm(ctx->model_kind == MODEL_MVN_SUM, ctx->grad_gemm, ctx->C, ctx->d, ctx->n_rows);
}

int launch_grad(svgdb_ctx *ctx, cudaStream_t stream)
{
    if (ctx->n_rows <= 0) return SVGDB_OK;
    if (ctx->model_kind == MODEL_HOOK) {
        int rc = ctx->hook(ctx->X[ctx->cur], ctx->G, ctx->N, ctx->d, ctx->row0, ctx->n_rows, (void *)stream, ctx->hook_user);
        if (rc != 0) return fail(ctx, SVGDB_ERR_INVALID, "device gradient hook returned " + std::to_string(rc));
        return SVGDB_OK;
    }
    if (grad_use_gemm(ctx)) return launch_grad_gemm(ctx, stream);
    constexpr int PT = 16;
    size_t smem = ((size_t)3 * PT * ctx->d + 5 * PT) * sizeof(double);
    unsigned blocks = (unsigned)((ctx->n_rows + PT - 1) / PT);
    mvn_sum_grad_f64_kernel<PT><<<blocks, 128, smem, stream>>>(ctx->X[ctx->cur], ctx->N, ctx->d, ctx->row0, ctx->n_rows,
                                                                     ctx->C, ctx->means_dev, ctx->prec_dev, ctx->G);
    KERNEL_CHECK();
    return SVGDB_OK;
}

// Enqueue grad log p on the side stream, ordered after everything the main stream has been given so far.
int kick_grad(svgdb_ctx *ctx)
{
    if (!ctx->grad_pending) return SVGDB_OK;
    ctx->grad_pending = false;
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    if (ctx->profiling) cudaEventRecord(ctx->ev[7], ctx->side_stream);
    TRY(launch_grad(ctx, ctx->side_stream));
    if (ctx->profiling) cudaEventRecord(ctx->ev[8], ctx->side_stream);
    CU(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    return SVGDB_OK;
}

size_t phi_smem_bytes(int d, int dc)
{
    int kc = d < 64 ? ((d + 3) & ~3) : 64;
    int ldx = ((kc + 15) & ~15) + 4;
    return (size_t)(2 * 64 * ldx + 64 * (dc + 2) + 64) * sizeof(double);
}

template <int DC>
int launch_phi_dc(svgdb_ctx *ctx, const PhiArgs &a)
{
    dim3 grid((unsigned)((ctx->n_rows + 63) / 64), (unsigned)((ctx->d + DC - 1) / DC));
    prof_mark(ctx, 5);
    phi_f64_kernel<DC><<<grid, 128, phi_smem_bytes(ctx->d, DC), ctx->stream>>>(a);
    KERNEL_CHECK();
    prof_mark(ctx, 6);
    ++ctx->stats.phi_launches;
    return SVGDB_OK;
}

int launch_phi(svgdb_ctx *ctx, bool debug_phi)
{
    if (ctx->n_rows <= 0) return SVGDB_OK;
    PhiArgs a{};
    a.X = ctx->X[ctx->cur];
    a.V = ctx->V;
    a.r = ctx->r;
    a.a_ptr = ctx->a_dev;
    a.n_total = ctx->N;
    a.d = ctx->d;
    a.row0 = ctx->row0;
    a.n_rows = ctx->n_rows;
    a.opt = ctx->opt;
    a.s1 = ctx->s1;
    a.s2 = ctx->s2;
    a.lb = ctx->lb;
    a.ub = ctx->ub;
    a.X_out = ctx->X[ctx->cur ^ 1];
    a.phi_out = debug_phi ? ctx->phi_dbg : nullptr;
    if (ctx->d <= 8) return launch_phi_dc<8>(ctx, a);
    if (ctx->d <= 16) return launch_phi_dc<16>(ctx, a);
    if (ctx->d <= 32) return launch_phi_dc<32>(ctx, a);
    return launch_phi_dc<64>(ctx, a);
}

int launch_make_v(svgdb_ctx *ctx)
{
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) { // centred, fp32: see make_v32_kernel
        if (ctx->n_rows > 0) {
            int64_t cnt = ctx->n_rows * ctx->d;
            svgdb::tc::make_v32_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->G, ctx->colsum, ctx->a_dev, ctx->N,
                                                                                              ctx->row0, ctx->n_rows, ctx->d, ctx->V32);
            KERNEL_CHECK();
        }
        if (ctx->world > 1) {
            size_t chunk = (size_t)ctx->rows_per_rank * ctx->d;
            NC(nccl().AllGather(ctx->V32 + (size_t)ctx->rank * chunk, ctx->V32, chunk, ncclFloat, ctx->comm, ctx->stream));
        }
        return SVGDB_OK;
    }
#endif
    if (ctx->n_rows > 0) {
        int64_t cnt = ctx->n_rows * ctx->d;
        make_v_f64_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->G, ctx->a_dev, ctx->row0,
                                                                                   ctx->n_rows, ctx->d, ctx->V);
        KERNEL_CHECK();
    }
    return allgather_rows(ctx, ctx->V, ctx->d);
}

#ifdef SVGDB_WITH_TC32
// column sums (for centring) and the centred bf16-split operand rows of the distance pass
// column sums of `X` (rows [0, rows)) into ctx->colsum on the main stream
int launch_colsum(svgdb_ctx *ctx, const double *X, int64_t rows)
{
    using namespace svgdb::tc;
    CU(cudaMemsetAsync(ctx->colsum, 0, 256 * 8, ctx->stream));
    if (ctx->wide) colsum_wide_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(X, rows, ctx->d, ctx->colsum);
    else colsum_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(X, rows, ctx->d, ctx->colsum);
    KERNEL_CHECK();
    return SVGDB_OK;
}

// centred operand rows [r0, r1) of the d <= 64 distance pass, in the variant ctx->dist_ops_f16 says
int launch_split_dist2(svgdb_ctx *ctx, int64_t r0, int64_t r1)
{
    using namespace svgdb::tc;
    if (r1 <= r0) return SVGDB_OK;
    const unsigned blocks = (unsigned)((r1 - r0 + 7) / 8);
    if (ctx->dist_ops_f16)
        split_dist2h_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->colsum, ctx->N, r0, r1, ctx->n_pad128, ctx->d, std::sqrt(ctx->dist_s2),
                                                             ctx->XA2, ctx->XB2, reinterpret_cast<__nv_bfloat16 *>(ctx->UA2),
                                                             reinterpret_cast<__nv_bfloat16 *>(ctx->WB2), ctx->rt);
    else
        split_dist2_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->colsum, ctx->N, r0, r1, ctx->n_pad128, ctx->d,
                                                            reinterpret_cast<__nv_bfloat16 *>(ctx->XA2), ctx->XBD, reinterpret_cast<__nv_bfloat16 *>(ctx->UA2),
                                                            reinterpret_cast<__nv_bfloat16 *>(ctx->WB2), ctx->rt);
    KERNEL_CHECK();
    return SVGDB_OK;
}

int launch_dist_operands(svgdb_ctx *ctx)
{
    using namespace svgdb::tc;
    // The first pass of a step with a median history is a folded collecting pass over a predicted bracket: it runs on the scaled
    // fp16 operands (two products).  Anything else -- no history yet, narrowing passes after a miss -- runs on the bf16 ones.
    TRY(finish_median(ctx));
    ctx->dist_ops_f16 = false;
    if (!ctx->wide && ctx->dist_f16 != 0 && ctx->dist_fold != 0 && ctx->scale_method == SVGDB_SCALE_MEDIAN && ctx->n_hist > 0 &&
        (double)ctx->N * (double)ctx->N > 65536.0 && ctx->med_hist[0] > 0.0 && std::isfinite(ctx->med_hist[0])) {
        const int k = (int)std::lround(std::log2(512.0 / ctx->med_hist[0]));
        if (k > -100 && k < 100) {
            ctx->dist_ops_f16 = true;
            ctx->dist_s2 = std::ldexp(1.0, k); // the median of the scaled squared distances lands in [362, 724)
        }
    }
    if (ctx->wide) { // 64 < d <= 256: kernels_dist_wide.cuh (no chunked upload on this path)
        TRY(launch_colsum(ctx, ctx->X[ctx->cur], ctx->N));
        const int64_t rows_a = ctx->n_pad128 + 256;
        split_distw_kernel<<<(unsigned)((rows_a + 7) / 8), 256, 0, ctx->stream>>>(
            ctx->X[ctx->cur], ctx->colsum, ctx->N, rows_a, ctx->n_pad128, ctx->d, ctx->dp, reinterpret_cast<__nv_bfloat16 *>(ctx->XA2), ctx->XBD,
            reinterpret_cast<__nv_bfloat16 *>(ctx->UA2), reinterpret_cast<__nv_bfloat16 *>(ctx->WB2), ctx->rt);
        KERNEL_CHECK();
        return SVGDB_OK;
    }
    CU(cudaMemsetAsync(ctx->colsum, 0, 256 * 8, ctx->stream));
    const int64_t rows_a = ctx->n_pad128 + 256;
    // particles still arriving in row chunks (svgdb_step_host): centre on the mean of the first chunk and prepare its rows only,
    // the other chunks are prepared by the first distance pass as they land (launch_dist_pass_tc32)
    const int64_t rows_now = ctx->up_chunks > 0 ? (int64_t)ctx->up_ipairs[0] * 256 : ctx->N;
    const int64_t rows_end = ctx->up_chunks > 0 ? rows_now : rows_a;
    if (ctx->up_chunks > 0) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[0], 0));
    colsum_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(ctx->X[ctx->cur], rows_now, ctx->d, ctx->colsum);
    KERNEL_CHECK();
    if (ctx->up_chunks > 0) {
        colsum_scale_kernel<<<1, 64, 0, ctx->stream>>>(ctx->colsum, ctx->d, (double)ctx->N / (double)rows_now);
        KERNEL_CHECK();
    }
    return launch_split_dist2(ctx, 0, rows_end);
}

int launch_dist_pass_tc32(svgdb_ctx *ctx, int mode, uint64_t lo, uint64_t hi, int shift)
{
    using namespace svgdb::tc;
    float lo_f = lo == 0 ? -INFINITY : key_to_float_ceil(lo); // nothing lies below an open lower end (D2 may round slightly negative)
    float hi_f = key_to_float_ceil(hi);
    CU(cudaMemsetAsync(ctx->below, 0, sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(ctx->max_below, 0, sizeof(unsigned long long), ctx->stream));
    // Symmetric enumeration over all rows at any world size; i-pairs are dealt cyclically to the ranks.
    // Particles arriving in row chunks (svgdb_step_host, one rank): the pass is issued in as many launches, launch k covering
    // the pairs whose later row lies in chunk k, i.e. i-pairs below the chunk's end against the column tiles of the chunk.
    // Bracket width as the kernel tests it, and the exclusive key bound of what the pass may collect.  Pure host arithmetic on
    // (lo, hi): computed on EVERY rank, also on ranks that own no i-pair (they take part in the select's collectives and must
    // derive the same digit plan from the same hi).
    uint32_t width_bits = 0;
    {
        float wdt = std::isinf(lo_f) || std::isinf(hi_f) ? INFINITY : (float)((double)hi_f - (double)lo_f);
        if ((double)wdt < (double)hi_f - (double)lo_f) wdt = std::nextafterf(wdt, INFINITY);
        wdt = std::nextafterf(std::nextafterf(wdt, INFINITY), INFINITY); // strictly above the rounded difference of any pair
        if (!(wdt > 0.0f)) wdt = std::numeric_limits<float>::min();
        std::memcpy(&width_bits, &wdt, 4);
        // every collected distance satisfies fl(d2 - lo) < wdt, hence d2 < lo + wdt (1 + 2^-24): an exclusive upper key bound
        float hi_ext = std::isinf(wdt) || std::isinf(lo_f) ? hi_f : (float)((double)lo_f + (double)wdt * (1.0 + 1.0 / 8388608.0));
        hi_ext = std::nextafterf(std::nextafterf(hi_ext, INFINITY), INFINITY);
        ctx->collect_hi_ext = std::isinf(hi_ext) ? hi : std::max<uint64_t>(hi, key_of((double)hi_ext) + 1);
    }
    // A collecting pass over a predicted bracket defines its own values: -lo is folded into the norm K chunk (exact three-term bf16
    // split of the float) and the accumulator is d2 - lo.  Passes that must agree with one another distance by distance (narrowing
    // histograms and the collecting pass that follows them) run unfolded.
    const bool fold = mode == MODE_COLLECT && ctx->dist_fold_next && ctx->dist_fold != 0 && std::isfinite(lo_f) && std::isfinite(hi_f);
    ctx->dist_fold_next = false;
    const bool f16 = fold && ctx->dist_ops_f16 && !ctx->wide;
    if (!f16 && ctx->dist_ops_f16) { // this pass needs the bf16 operands (a narrowing or unfolded pass after a miss): rebuild them
        ctx->dist_ops_f16 = false;
        TRY(launch_split_dist2(ctx, 0, ctx->n_pad128 + 256));
    }
    float out_scale = 1.0f;
    if (f16) { // the operands are scaled by s, the distances by s^2 (a power of two: bracket end, width and every comparison scale exactly)
        const float s2 = (float)ctx->dist_s2;
        float wdt;
        std::memcpy(&wdt, &width_bits, 4);
        wdt *= s2;
        lo_f *= s2;
        hi_f *= s2;
        if (!std::isfinite(wdt) || !std::isfinite(lo_f) || !std::isfinite(hi_f) || !(wdt > 0.0f))
            return fail(ctx, SVGDB_ERR_NUMERIC, "median select: the predicted bracket leaves the fp32 range after scaling");
        std::memcpy(&width_bits, &wdt, 4);
        out_scale = (float)(1.0 / ctx->dist_s2);
    }
    uint32_t fold_l01 = 0, fold_l2 = 0;
    if (fold) {
        float rem = -lo_f;
        uint16_t t[3];
        for (int k = 0; k < 3; ++k) { // bf16 by truncation: the remainder of a 24-bit float is exact after three 8-bit terms
            uint32_t bits;
            std::memcpy(&bits, &rem, 4);
            bits &= 0xFFFF0000u;
            float term;
            std::memcpy(&term, &bits, 4);
            t[k] = (uint16_t)(bits >> 16);
            rem -= term;
        }
        fold_l01 = (uint32_t)t[0] | ((uint32_t)t[1] << 16);
        fold_l2 = (uint32_t)t[2];
    }
    if (mode == MODE_HIST) CU(cudaMemsetAsync(ctx->hist, 0, HIST_BINS * sizeof(unsigned long long), ctx->stream));
    else CU(cudaMemsetAsync(ctx->cand_count, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->wide) { // 64 < d <= 256: one i-tile per CTA (kernels_dist_wide.cuh); i-tiles dealt cyclically to the ranks
        const int n_itiles_all = (int)(ctx->n_pad128 / 128);
        const int n_itiles = (n_itiles_all - ctx->rank + ctx->world - 1) / ctx->world;
        if (n_itiles > 0) {
            DistWArgs b{};
            b.XA = reinterpret_cast<const __nv_bfloat16 *>(ctx->XA2);
            b.UA = reinterpret_cast<const __nv_bfloat16 *>(ctx->UA2);
            b.WB = reinterpret_cast<const __nv_bfloat16 *>(ctx->WB2);
            b.n_total = ctx->N;
            b.n_junits = (int)(ctx->n_pad128 / 64);
            b.tile_offset = ctx->rank;
            b.tile_stride = ctx->world;
            b.n_itiles = n_itiles;
            b.lo_f = lo_f;
            b.hi_f = hi_f;
            b.open_low = std::isinf(lo_f) ? 1 : 0;
            b.width_bits = width_bits;
            b.fold_l01 = fold_l01;
            b.fold_l2 = fold_l2;
            b.lo_key = lo;
            b.shift = shift;
            b.below = ctx->below;
            b.hist = ctx->hist;
            b.cand = ctx->cand;
            b.cand_count = ctx->cand_count;
            b.capacity = ctx->capacity;
            b.err = ctx->tc_err;
            long long units = 0;
            for (int l = 0; l < n_itiles; ++l) units += std::max(0, b.n_junits - 2 * (b.tile_offset + b.tile_stride * l));
            const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ctx->sm_count, units));
            // variants: histogram (unfolded), collecting unfolded (after a histogram, or an open lower end), collecting folded + gated (predicted bracket)
#define SVGDB_DW_LAUNCH(DPV)                                                                                                                   \
    case DPV:                                                                                                                                  \
        if (mode == MODE_HIST) distw_tc32_kernel<DPV, MODE_HIST, false, false><<<grid, DW_THREADS, DWCfg<DPV>::SMEM, ctx->stream>>>(ctx->mapBWP, b); \
        else if (fold) distw_tc32_kernel<DPV, MODE_COLLECT, true, true><<<grid, DW_THREADS, DWCfg<DPV>::SMEM, ctx->stream>>>(ctx->mapBWP, b);  \
        else distw_tc32_kernel<DPV, MODE_COLLECT, false, false><<<grid, DW_THREADS, DWCfg<DPV>::SMEM, ctx->stream>>>(ctx->mapBWP, b);          \
        break;
            if (units > 0) {
                switch (ctx->dp) {
                    SVGDB_DW_LAUNCH(128)
                    SVGDB_DW_LAUNCH(192)
                    SVGDB_DW_LAUNCH(256)
                default: return fail(ctx, SVGDB_ERR_DIMENSION, "internal: no wide distance kernel for this dimension");
                }
                KERNEL_CHECK();
            }
#undef SVGDB_DW_LAUNCH
        }
        ++ctx->stats.median_passes;
        TRY(kick_grad(ctx));
        TRY(reduce_pass_words(ctx, mode));
        return SVGDB_OK;
    }
    const int n_ipairs_all = (int)((ctx->N + 255) / 256);
    const int n_launches = std::max(1, ctx->up_chunks);
    for (int ch = 0; ch < n_launches; ++ch) {
        int n_ipairs = (n_ipairs_all - ctx->rank + ctx->world - 1) / ctx->world; // this rank's share
        int jt_begin = 0, jt_end = (int)(ctx->n_pad128 / 128);
        if (ctx->up_chunks > 0) {
            const bool last = ch == ctx->up_chunks - 1;
            n_ipairs = last ? n_ipairs_all : ctx->up_ipairs[ch];
            jt_begin = ch == 0 ? 0 : 2 * ctx->up_ipairs[ch - 1];
            if (!last) jt_end = 2 * ctx->up_ipairs[ch];
            if (ch > 0) { // this chunk's rows have landed: centre and split them (the last launch also writes the padding rows)
                using namespace svgdb::tc;
                CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[ch], 0));
                const int64_t r0 = (int64_t)ctx->up_ipairs[ch - 1] * 256, r1 = last ? ctx->n_pad128 + 256 : (int64_t)ctx->up_ipairs[ch] * 256;
                TRY(launch_split_dist2(ctx, r0, r1));
            }
        }
        if (n_ipairs > 0) {
            Dist2Args b{};
            b.XA2 = reinterpret_cast<const __nv_bfloat16 *>(ctx->XA2);
            b.UA = reinterpret_cast<const __nv_bfloat16 *>(ctx->UA2);
            b.WB = reinterpret_cast<const __nv_bfloat16 *>(ctx->WB2);
            b.n_total = ctx->N;
            b.n_jtiles = jt_end;
            b.jt_begin = jt_begin;
            b.pair_offset = ctx->rank;
            b.pair_stride = ctx->world;
            b.n_ipairs = n_ipairs;
            b.lo_f = lo_f;
            b.hi_f = hi_f;
            b.open_low = std::isinf(lo_f) ? 1 : 0;
            b.width_bits = width_bits;
            b.fold_l01 = fold_l01;
            b.fold_l2 = fold_l2;
            b.out_scale = out_scale;
            b.lo_key = lo;
            b.shift = shift;
            b.below = ctx->below;
            b.hist = ctx->hist;
            b.cand = ctx->cand;
            b.cand_count = ctx->cand_count;
            b.capacity = ctx->capacity;
            b.err = ctx->tc_err;
            b.dbg = ctx->dist_dbg_mode;
            long long units = 0;
            for (int l = 0; l < n_ipairs; ++l) units += std::max(0, b.n_jtiles - std::max(2 * (b.pair_offset + b.pair_stride * l), b.jt_begin));
            const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ctx->sm_count, units));
            if (units <= 0) continue;
            if (mode == MODE_HIST) {
                // a unit is 256 x 128 pairs of weight <= 2: 32-bit shared counters while a CTA's share stays below 2^32
                const bool hist32 = (units / grid + 1) * 65536ll < 4294967296ll;
                if (hist32) dist2_tc32_kernel<MODE_HIST, true, false><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapBD, b);
                else dist2_tc32_kernel<MODE_HIST, false, false><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapBD, b);
            } else {
                // expected share of pairs inside the bracket (density of the last pass x relative width): with fewer than ~1 hit per
                // two 32 x 32 warp chunks the gated epilogue (1.5 instructions per distance + rare collection) is the cheaper one
                double width_rel = 1.0;
                if (lo > 0 && hi < KEY_END) {
                    double dlo, dhi;
                    std::memcpy(&dlo, &lo, 8);
                    std::memcpy(&dhi, &hi, 8);
                    if (dhi > 0.0) width_rel = (dhi - dlo) / dhi;
                }
                const bool gated = ctx->dist_gated >= 0 ? ctx->dist_gated != 0 : ctx->density * width_rel * 1024.0 < 0.5;
                if (f16) {
                    if (gated) dist2_tc32_kernel<MODE_COLLECT, true, true, true><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapB2, b);
                    else dist2_tc32_kernel<MODE_COLLECT, false, true, true><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapB2, b);
                } else if (fold) {
                    if (gated) dist2_tc32_kernel<MODE_COLLECT, true, true><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapBD, b);
                    else dist2_tc32_kernel<MODE_COLLECT, false, true><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapBD, b);
                } else {
                    if (gated) dist2_tc32_kernel<MODE_COLLECT, true, false><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapBD, b);
                    else dist2_tc32_kernel<MODE_COLLECT, false, false><<<grid, D2_THREADS, D2_SMEM, ctx->stream>>>(ctx->mapBD, b);
                }
            }
            KERNEL_CHECK();
        }
    }
    ctx->up_chunks = 0; // everything has landed and is prepared
    ++ctx->stats.median_passes;
    TRY(kick_grad(ctx));
    TRY(reduce_pass_words(ctx, mode));
    return SVGDB_OK;
}

// Which arithmetic the tensor-core pair kernel runs (kernels_phi_tc.cuh, DESIGN.md "Precision modes").  The fast variant's error
// terms (fp16 rounding of the column particle and of the kernel values) are zero-mean and average over a row's effective neighbours:
// it is the default only where that holds by construction -- one Gaussian target, at least 16,384 particles and d >= 8 (the
// rounding of y^_j also averages over the coordinates: at d = 2, N = 65,536 the outermost particle's row is off by 2.2e-4 of max|phi|,
// 3e-5 at d = 8); mixtures, user models behind the gradient hook, small particle sets and d < 8 get the precise variant (1.5x the MMAs).
bool tc32_precise(const svgdb_ctx *ctx)
{
    return svgdb::host::rule_tc32_precise(ctx->tc32_variant, ctx->model_kind == MODEL_MVN_SUM && ctx->C == 1, ctx->N, ctx->d);
}

// The particle-side operands of the pair kernel (they need X and the bandwidth, not V) and the zeroed accumulator.
int launch_phi_x_operands(svgdb_ctx *ctx, cudaStream_t stream)
{
    using namespace svgdb::tc;
    if (ctx->n_rows <= 0) return SVGDB_OK;
    CU(cudaMemsetAsync(ctx->phi_buf, 0, (size_t)(ctx->n_pad128 + 256) * (ctx->wide ? ctx->dp + 16 : TC_PHI_LD) * 4, stream));
    // the bandwidth is folded into the operands, so the accumulator of the first contraction is the exponent
    const int64_t rows_a = ctx->n_pad128 + 256;
    if (ctx->wide) {
        const bool prec = tc32_precise(ctx);
        split_phiw_kernel<<<(unsigned)((rows_a + 7) / 8), 256, 0, stream>>>(ctx->X[ctx->cur], ctx->colsum, ctx->a_dev, ctx->N, rows_a, ctx->n_pad128, ctx->d,
                                                                             ctx->dp, ctx->XA2, prec ? reinterpret_cast<__half *>(ctx->XBD) : ctx->XB2,
                                                                             ctx->UA2, ctx->WB2, prec ? 1 : 0);
        KERNEL_CHECK();
        return SVGDB_OK;
    }
    // precise variant: the column operand keeps both fp16 terms; it goes into the distance pass's column buffer (free by now)
    const bool precise = tc32_precise(ctx);
    split_phi2_kernel<<<(unsigned)((rows_a + 7) / 8), 256, 0, stream>>>(ctx->X[ctx->cur], ctx->colsum, ctx->a_dev, ctx->N, rows_a, ctx->n_pad128,
                                                                         ctx->d, ctx->XA2, precise ? reinterpret_cast<__half *>(ctx->XBD) : ctx->XB2,
                                                                         ctx->UA2, ctx->WB2, precise ? 1 : 0, phi_use_f8(ctx) ? ctx->XB8 : nullptr);
    KERNEL_CHECK();
    return SVGDB_OK;
}

// 64 < d <= 256: kernels_phi_wide.cuh (one i-tile per CTA, Phi in column groups at d > 192), then the FP64 optimizer kernel
int launch_phi_wide(svgdb_ctx *ctx, bool debug_phi)
{
    using namespace svgdb::tc;
    const int dp = ctx->dp;
    make_vtw_kernel<<<dim3((unsigned)(ctx->n_pad128 / 64), (unsigned)(dp / 64)), 256, 0, ctx->stream>>>(ctx->V32, ctx->N, ctx->n_pad128, ctx->d, dp, ctx->VT2);
    KERNEL_CHECK();
    PhiWArgs a{};
    a.phi_buf = ctx->phi_buf;
    a.XA = ctx->XA2;
    a.UA = ctx->UA2;
    a.WB = ctx->WB2;
    a.row0 = ctx->row0;
    a.n_rows = ctx->n_rows;
    a.n_junits = (int)(ctx->n_pad128 / 64);
    const int cl = (ctx->phi_cluster < 0 ? true : ctx->phi_cluster != 0) ? 2 : 1; // CTAs per cluster sharing the column tiles by TMA multicast
    a.n_itiles = (int)((ctx->n_rows + 128 * cl - 1) / (128 * cl));
    a.max_seg = ctx->phi_max_seg > 0 ? ctx->phi_max_seg : (tc32_precise(ctx) ? 64 : 128); // column units (64 particles) per flush
    a.dbg = 0;
    a.err = ctx->tc_err;
    const bool precise = tc32_precise(ctx);
    const int groups = dp == 256 ? 2 : 1;
    const long long units = (long long)a.n_itiles * groups * a.n_junits;
    const unsigned n_clusters = (unsigned)std::max<long long>(1, std::min<long long>(ctx->sm_count / cl, units));
    prof_mark(ctx, 5);
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(n_clusters * cl);
        cfg.blockDim = dim3(PW_THREADS);
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)cl;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
#define SVGDB_PW_LAUNCH(DPV, PR, CLV)                                                                                   \
    do {                                                                                                                \
        cfg.dynamicSmemBytes = PWCfg<DPV, PR>::SMEM;                                                                    \
        CU(cudaLaunchKernelEx(&cfg, phiw_tc32_kernel<DPV, PR, CLV>, PR ? ctx->mapBWP : ctx->mapBW, ctx->mapVW, a));       \
    } while (0)
#define SVGDB_PW_DP(DPV)                                                                                                \
    case DPV:                                                                                                           \
        if (precise) { if (cl == 2) SVGDB_PW_LAUNCH(DPV, true, 2); else SVGDB_PW_LAUNCH(DPV, true, 1); }                \
        else { if (cl == 2) SVGDB_PW_LAUNCH(DPV, false, 2); else SVGDB_PW_LAUNCH(DPV, false, 1); }                      \
        break;
        switch (dp) {
            SVGDB_PW_DP(128)
            SVGDB_PW_DP(192)
            SVGDB_PW_DP(256)
        default: return fail(ctx, SVGDB_ERR_DIMENSION, "internal: no wide pair kernel for this dimension");
        }
#undef SVGDB_PW_DP
#undef SVGDB_PW_LAUNCH
    }
    KERNEL_CHECK();
    prof_mark(ctx, 6);
    OptWArgs o{};
    o.X = ctx->X[ctx->cur];
    o.colsum = ctx->colsum;
    o.phi_buf = ctx->phi_buf;
    o.a_ptr = ctx->a_dev;
    o.n_total = ctx->N;
    o.row0 = ctx->row0;
    o.n_rows = ctx->n_rows;
    o.state_row0 = ctx->row0;
    o.d = ctx->d;
    o.ld = dp + 16;
    o.ones_col = dp;
    o.opt = ctx->opt;
    o.s1 = ctx->s1;
    o.s2 = ctx->s2;
    o.lb = ctx->lb;
    o.ub = ctx->ub;
    o.X_out = ctx->X[ctx->cur ^ 1];
    o.phi_out = debug_phi ? ctx->phi_dbg : nullptr;
    o.miss = ctx->miss_dev;
    const int64_t cnt = ctx->n_rows * ctx->d;
    opt_update_wide_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(o);
    KERNEL_CHECK();
    ++ctx->stats.phi_launches;
    ctx->streamed = false;
    return SVGDB_OK;
}

int launch_phi_tc32(svgdb_ctx *ctx, bool debug_phi, bool x_operands_done = false)
{
    using namespace svgdb::tc;
    if (ctx->n_rows <= 0) return SVGDB_OK;
    if (!x_operands_done) TRY(launch_phi_x_operands(ctx, ctx->stream));
    if (ctx->wide) return launch_phi_wide(ctx, debug_phi);
    make_vt2_kernel<<<(unsigned)(ctx->n_pad128 / 64), 256, 0, ctx->stream>>>(ctx->V32, ctx->N, ctx->n_pad128, ctx->d, ctx->VT2,
                                                                              phi_use_f8(ctx) ? ctx->VT8 : nullptr);
    KERNEL_CHECK();
    // Row chunks.  Normally one; when the updated rows are wanted on the host (svgdb_step_host) the rows are processed in four
    // launches of 3/8, 3/8, 1/8, 1/8 of the i-pairs, each followed by its optimizer kernel and a device-to-host copy on the
    // copy stream, so that all but the last eighth of the PCIe transfer hides behind the remaining pair interactions.
    const int n_ipairs_all = (int)((ctx->n_rows + 255) / 256);
    const bool stream_rows = ctx->stream_out != nullptr && !debug_phi && ctx->phi_dbg_mode == 0 && n_ipairs_all >= 32;
    const int n_chunks = stream_rows ? 4 : 1;
    int chunk_ipairs[4] = {n_ipairs_all, 0, 0, 0};
    if (stream_rows) download_chunks(n_ipairs_all, chunk_ipairs);
    int ipair0 = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int64_t off = (int64_t)ipair0 * 256;
        const int64_t rows = std::min<int64_t>((int64_t)chunk_ipairs[ch] * 256, ctx->n_rows - off);
        ipair0 += chunk_ipairs[ch];
        if (rows <= 0) continue;
        Phi2Args a{};
        a.phi_buf = ctx->phi_buf;
        a.XA2 = ctx->XA2;
        a.UA = ctx->UA2;
        a.WB = ctx->WB2;
        a.row0 = ctx->row0 + off;
        a.n_rows = rows;
        a.n_jtiles = (int)(ctx->n_pad128 / 128);
        const int cl = (ctx->phi_cluster < 0 ? false : ctx->phi_cluster != 0) ? 2 : 1; // CTAs per cluster sharing the column tiles by TMA multicast
        a.n_ipairs = (chunk_ipairs[ch] + cl - 1) / cl; // i-pair groups
        a.max_seg = ctx->phi_max_seg > 0 ? ctx->phi_max_seg : (tc32_precise(ctx) ? 32 : 128); // j-tiles (128 particles) per flush
        a.poly = ctx->phi_poly;
        a.no_vlo = svgdb::host::rule_phi_one_term_v(ctx->phi_no_vlo, ctx->N) ? 1 : 0;
        a.dbg = ctx->phi_dbg_mode;
        a.err = ctx->tc_err;
        a.trace = ctx->tc_trace;
        const long long units = (long long)a.n_ipairs * a.n_jtiles;
        const unsigned n_clusters = (unsigned)std::max<long long>(1, std::min<long long>(ctx->sm_count / cl, units)); // persistent: one CTA per SM
        if (ch == 0) prof_mark(ctx, 5);
        const bool precise = tc32_precise(ctx);
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(n_clusters * cl);
            cfg.blockDim = dim3(P2_THREADS);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr{};
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = (unsigned)cl;
            attr.val.clusterDim.y = 1;
            attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
#define SVGDB_P2_LAUNCH(P, PR, CLV)                                                                                   \
    do {                                                                                                              \
        cfg.dynamicSmemBytes = P2Cfg<PR>::SMEM;                                                                       \
        CU(cudaLaunchKernelEx(&cfg, phi2_tc32_kernel<P, PR, CLV>, PR ? ctx->mapBD : ctx->mapB2, ctx->mapV2, a, ctx->mapB8, ctx->mapV8)); \
    } while (0)
#define SVGDB_PHI2_CASE(P)                                                                                            \
    case P:                                                                                                           \
        if (precise) { if (cl == 2) SVGDB_P2_LAUNCH(P, true, 2); else SVGDB_P2_LAUNCH(P, true, 1); }                  \
        else if (phi_use_f8(ctx) && cl == 1) {                                                                            \
            cfg.dynamicSmemBytes = P2Cfg<false, true>::SMEM;                                                          \
            if (ctx->phi_tcsum)                                                                                        \
                CU(cudaLaunchKernelEx(&cfg, phi2_tc32_kernel<P, false, 1, true, true>, ctx->mapB2, ctx->mapV2, a, ctx->mapB8, ctx->mapV8)); \
            else                                                                                                       \
                CU(cudaLaunchKernelEx(&cfg, phi2_tc32_kernel<P, false, 1, true, false>, ctx->mapB2, ctx->mapV2, a, ctx->mapB8, ctx->mapV8)); \
        } else { if (cl == 2) SVGDB_P2_LAUNCH(P, false, 2); else SVGDB_P2_LAUNCH(P, false, 1); }                      \
        break;
            switch (ctx->phi_poly) {
                SVGDB_PHI2_CASE(0)
                SVGDB_PHI2_CASE(1)
                SVGDB_PHI2_CASE(2)
                SVGDB_PHI2_CASE(4)
            default: return fail(ctx, SVGDB_ERR_INVALID, "SVGDB_PHI_POLY must be 0, 1, 2 or 4");
            }
#undef SVGDB_PHI2_CASE
#undef SVGDB_P2_LAUNCH
        }
        KERNEL_CHECK();
        if (ch == n_chunks - 1) prof_mark(ctx, 6);
        if (ch == 0 && ctx->tc_trace) {
            std::vector<long long> h(3 * 64 * 8);
            CU(cudaMemcpyAsync(h.data(), ctx->tc_trace, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (FILE *f = std::fopen(std::getenv("SVGDB_TC_TRACE"), "w")) {
                for (size_t k = 0; k < h.size(); ++k) std::fprintf(f, "%lld%c", h[k], (k % 8 == 7) ? '\n' : ' ');
                std::fclose(f);
            }
        }
        const int64_t cnt = rows * ctx->d;
        if (ctx->phi_dbg_mode != 0 && !debug_phi) { // development: the kernel produced garbage on purpose, keep the particles
            CU(cudaMemcpyAsync(ctx->X[ctx->cur ^ 1] + a.row0 * ctx->d, ctx->X[ctx->cur] + a.row0 * ctx->d, (size_t)cnt * sizeof(double),
                               cudaMemcpyDeviceToDevice, ctx->stream));
            continue;
        }
        OptTcArgs o{};
        o.X = ctx->X[ctx->cur];
        o.colsum = ctx->colsum;
        o.phi_buf = ctx->phi_buf;
        o.a_ptr = ctx->a_dev;
        o.n_total = ctx->N;
        o.row0 = a.row0;
        o.n_rows = rows;
        o.state_row0 = ctx->row0;
        o.d = ctx->d;
        o.opt = ctx->opt;
        o.s1 = ctx->s1;
        o.s2 = ctx->s2;
        o.lb = ctx->lb;
        o.ub = ctx->ub;
        o.X_out = ctx->X[ctx->cur ^ 1];
        o.phi_out = debug_phi ? ctx->phi_dbg : nullptr;
        o.miss = ctx->miss_dev;
        opt_update_tc32_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(o);
        KERNEL_CHECK();
        if (stream_rows) {
            CU(cudaEventRecord(ctx->ev_chunk[ch], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[ch], 0));
            CU(cudaMemcpyAsync(ctx->stream_out + (size_t)off * ctx->d, ctx->X[ctx->cur ^ 1] + (size_t)a.row0 * ctx->d, (size_t)cnt * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
    }
    ++ctx->stats.phi_launches;
    ctx->streamed = stream_rows;
    return SVGDB_OK;
}

int check_tc_err(svgdb_ctx *ctx)
{
    int e = 0;
    CU(cudaMemcpyAsync(&e, ctx->tc_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (e != 0) {
        cudaMemsetAsync(ctx->tc_err, 0, 4, ctx->stream);
        return fail(ctx, SVGDB_ERR_CUDA, "tensor-core pipeline timed out waiting on barrier tag " + std::to_string(e));
    }
    return SVGDB_OK;
}
#endif

int check_ready(svgdb_ctx *ctx)
{
    if (ctx->model_kind == MODEL_UNSET) return fail(ctx, SVGDB_ERR_UNSET, "Model function is unset.");
    if (!ctx->kernel_set) return fail(ctx, SVGDB_ERR_UNSET, "Kernel function is unset.");
    return SVGDB_OK;
}

void prof_mark(svgdb_ctx *ctx, int i)
{
    if (ctx->profiling) cudaEventRecord(ctx->ev[i], ctx->stream);
}

// ---- ScaleMethod::Hessian (kernels_hessian.cuh) -----------------------------------------------------------------------
// Sum over all particles of the particle-dependent part of -Hessian(log p) and of the softmax weights (kernels_hessian.cuh).
int hessian_partial_sums(svgdb_ctx *ctx, std::vector<double> &H, std::vector<double> &W)
{
    const int d = ctx->d, C = ctx->C;
    const size_t dd = (size_t)d * d;
    if (!ctx->Wsum_dev) CU(cudaMalloc(&ctx->Wsum_dev, (size_t)C * sizeof(double))); // released when the model changes
    CU(cudaMemsetAsync(ctx->Hsum_dev, 0, dd * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(ctx->Wsum_dev, 0, (size_t)C * sizeof(double), ctx->stream));
    if (ctx->n_rows > 0) {
        const int tiles = (d + 63) / 64;
        int pt = 16; // particles per group: as many as the shared-memory budget allows
        while (pt > 1 && hessian_smem_doubles(pt, C, d) * sizeof(double) > 160 * 1024) pt >>= 1;
        const size_t smem = hessian_smem_doubles(pt, C, d) * sizeof(double);
        if (smem > 200 * 1024) return fail(ctx, SVGDB_ERR_DIMENSION, "ScaleMethod::Hessian: components x dimension too large for the Hessian kernel");
        const int64_t groups = (ctx->n_rows + pt - 1) / pt;
        dim3 grid((unsigned)std::min<int64_t>(groups, (int64_t)ctx->sm_count * 4), (unsigned)(tiles * tiles));
#define SVGDB_HESS_CASE(PT)                                                                                                           \
    case PT:                                                                                                                          \
        CU(cudaFuncSetAttribute(mvn_sum_hessian_f64_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        mvn_sum_hessian_f64_kernel<PT><<<grid, 256, smem, ctx->stream>>>(ctx->X[ctx->cur], d, ctx->row0, ctx->n_rows, C, ctx->means_dev, \
                                                                         ctx->prec_dev, tiles, ctx->Hsum_dev, ctx->Wsum_dev);         \
        break;
        switch (pt) {
            SVGDB_HESS_CASE(16)
            SVGDB_HESS_CASE(8)
            SVGDB_HESS_CASE(4)
            SVGDB_HESS_CASE(2)
            SVGDB_HESS_CASE(1)
        }
#undef SVGDB_HESS_CASE
        KERNEL_CHECK();
    }
    if (ctx->world > 1) {
        NC(nccl().AllReduce(ctx->Hsum_dev, ctx->Hsum_dev, dd, ncclDouble, ncclSum, ctx->comm, ctx->stream));
        NC(nccl().AllReduce(ctx->Wsum_dev, ctx->Wsum_dev, (size_t)C, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    }
    CU(cudaMemcpyAsync(H.data(), ctx->Hsum_dev, dd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(W.data(), ctx->Wsum_dev, (size_t)C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SVGDB_OK;
}

// A = 1/(2 d n) sum_i -Hessian(log p)(x_i) on the host (GaussianRBFKernel.hpp:189-210), its factor R and R^-1 on the device.
int hessian_scale_dev(svgdb_ctx *ctx)
{
    if (ctx->model_kind != MODEL_MVN_SUM)
        return fail(ctx, SVGDB_ERR_UNSET, "ScaleMethod::Hessian needs a model with a device Hessian (MultivariateNormal or a sum of them)");
    const int d = ctx->d, C = ctx->C;
    const size_t dd = (size_t)d * d;
    if (!ctx->Hsum_dev) {
        CU(cudaMalloc(&ctx->Hsum_dev, dd * sizeof(double)));
        CU(cudaMalloc(&ctx->R_dev, dd * sizeof(double)));
        CU(cudaMalloc(&ctx->Rt_dev, dd * sizeof(double)));
        CU(cudaMalloc(&ctx->Rinv_dev, dd * sizeof(double)));
        CU(cudaMalloc(&ctx->Y_dev, (size_t)ctx->n_pad * d * sizeof(double)));
        CU(cudaMalloc(&ctx->GH_dev, (size_t)ctx->rows_per_rank * d * sizeof(double)));
        CU(cudaMemsetAsync(ctx->Y_dev, 0, (size_t)ctx->n_pad * d * sizeof(double), ctx->stream));
    }
    if (ctx->hess_const_valid) return SVGDB_OK;
    std::vector<double> H(dd, 0.0), W((size_t)C, 0.0), R, Rinv;
    if (C == 1) {
        // One Gaussian: w = 1 and ybar = y, so the particle-dependent part of the Hessian vanishes identically (the kernel
        // would add exact zeros) and sum_i w_i = n: A = P / (2 d) for every X.
        W[0] = (double)ctx->N;
    } else {
        TRY(hessian_partial_sums(ctx, H, W));
    }
    ctx->A_host.assign(dd, 0.0);
    const double scale = 1.0 / (2.0 * (double)d * (double)ctx->N);
    for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) {
            double s = 0.5 * (H[(size_t)r * d + c] + H[(size_t)c * d + r]);
            for (int k = 0; k < C; ++k) s += W[(size_t)k] * ctx->prec_host[((size_t)k * d + r) * d + c]; // symmetric up to the rounding of the host inversion
            ctx->A_host[(size_t)r * d + c] = s * scale;
        }
    if (!cholesky_upper(ctx->A_host, d, R, Rinv))
        return fail(ctx, SVGDB_ERR_NUMERIC, "ScaleMethod::Hessian: the scale matrix (mean negative Hessian of log p over the particles) is not positive definite");
    std::vector<double> Rt(dd);
    for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) Rt[(size_t)r * d + c] = R[(size_t)c * d + r];
    CU(cudaMemcpyAsync(ctx->R_dev, R.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->Rt_dev, Rt.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->Rinv_dev, Rinv.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // R, Rt, Rinv are locals
    ctx->stats.last_scale = ctx->A_host[0];
    ctx->hess_const_valid = C == 1;
    return SVGDB_OK;
}

int launch_phi(svgdb_ctx *ctx, bool debug_phi);
int launch_make_v(svgdb_ctx *ctx);

// One Hessian-scaled phi (and update): the scalar-bandwidth pair kernel with a = 1 on y = R x, g^ = R^-T g, then phi = R^T phi^.
int prepare_and_phi_hessian(svgdb_ctx *ctx, bool debug_phi)
{
    const int d = ctx->d;
    prof_mark(ctx, 0);
    TRY(hessian_scale_dev(ctx));
    prof_mark(ctx, 1);
    TRY(launch_grad(ctx, ctx->stream));
    prof_mark(ctx, 2);
    // Y = X R^T (all rows: every rank holds X), G^ = G R^-1 (local rows).  Row-vector form: y_i = x_i R^T, g^_i = g_i R^-1.
    row_times_matrix_f64_kernel<false><<<row_times_matrix_blocks(ctx->N, d), 256, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->Rt_dev, ctx->N, d, ctx->Y_dev, RowApply{});
    KERNEL_CHECK();
    if (ctx->n_rows > 0) {
        row_times_matrix_f64_kernel<false><<<row_times_matrix_blocks(ctx->n_rows, d), 256, 0, ctx->stream>>>(ctx->G, ctx->Rinv_dev, ctx->n_rows, d, ctx->GH_dev, RowApply{});
        KERNEL_CHECK();
    }
    // the scalar-bandwidth machinery on the transformed quantities, a = 1, writing phi^ for the local rows into phi_dbg
    const double one = 1.0;
    CU(cudaMemcpyAsync(ctx->a_dev, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    double *X_keep = ctx->X[ctx->cur], *G_keep = ctx->G;
    ctx->X[ctx->cur] = ctx->Y_dev;
    ctx->G = ctx->GH_dev;
    int rc = SVGDB_OK;
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) { // the tensor-core pair kernel centres its operands: column sums of Y
        rc = launch_colsum(ctx, ctx->Y_dev, ctx->N);
        if (rc == SVGDB_OK) rc = launch_phi_x_operands(ctx, ctx->stream);
        if (rc == SVGDB_OK) rc = launch_make_v(ctx);
        prof_mark(ctx, 3);
        if (rc == SVGDB_OK) rc = launch_phi_tc32(ctx, true, true);
    } else
#endif
    {
        rc = launch_rownorm(ctx);
        if (rc == SVGDB_OK) rc = launch_make_v(ctx);
        prof_mark(ctx, 3);
        if (rc == SVGDB_OK) rc = launch_phi(ctx, true);
    }
    ctx->X[ctx->cur] = X_keep;
    ctx->G = G_keep;
    TRY(rc);
    // phi = phi^ R (rows): for ComputePhi through GH_dev back into phi_dbg, for a step straight into the optimizer update
    if (ctx->n_rows > 0) {
        const unsigned blocks = row_times_matrix_blocks(ctx->n_rows, d);
        if (debug_phi) {
            row_times_matrix_f64_kernel<false><<<blocks, 256, 0, ctx->stream>>>(ctx->phi_dbg, ctx->R_dev, ctx->n_rows, d, ctx->GH_dev, RowApply{});
            KERNEL_CHECK();
            CU(cudaMemcpyAsync(ctx->phi_dbg, ctx->GH_dev, (size_t)ctx->n_rows * d * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            RowApply ap{ctx->X[ctx->cur], ctx->X[ctx->cur ^ 1], ctx->row0, ctx->opt, ctx->s1, ctx->s2, ctx->lb, ctx->ub};
            row_times_matrix_f64_kernel<true><<<blocks, 256, 0, ctx->stream>>>(ctx->phi_dbg, ctx->R_dev, ctx->n_rows, d, nullptr, ap);
            KERNEL_CHECK();
        }
    }
    ++ctx->stats.phi_launches;
    prof_mark(ctx, 4);
    return SVGDB_OK;
}

// phi (and everything it needs) for the current X; leaves V, r, a on the device
int prepare_and_phi(svgdb_ctx *ctx, bool debug_phi)
{
    if (ctx->scale_method == SVGDB_SCALE_HESSIAN) return prepare_and_phi_hessian(ctx, debug_phi);
    prof_mark(ctx, 0);
    // grad log p only needs X: it runs on the side stream next to the bandwidth computation, enqueued right behind the
    // (persistent, GPU-filling) distance pass so that it overlaps the select kernels, which leave most of the GPU idle
    ctx->grad_pending = true; // kicked off behind the first distance pass (kick_grad), or below if there is none
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) TRY(launch_dist_operands(ctx));
#endif
    if (ctx->precision != SVGDB_PRECISION_TC32) TRY(launch_rownorm(ctx));
    TRY(compute_scale_dev(ctx));
    prof_mark(ctx, 1);
    TRY(kick_grad(ctx));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    prof_mark(ctx, 2);
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) {
        // the particle-side operands do not need V: they are built on the side stream while V is formed and all-gathered
        CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        TRY(launch_phi_x_operands(ctx, ctx->side_stream));
        CU(cudaEventRecord(ctx->ev_join, ctx->side_stream));
        TRY(launch_make_v(ctx));
        prof_mark(ctx, 3);
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        TRY(launch_phi_tc32(ctx, debug_phi, true));
        prof_mark(ctx, 4);
        return SVGDB_OK;
    }
#endif
    TRY(launch_make_v(ctx));
    prof_mark(ctx, 3);
    TRY(launch_phi(ctx, debug_phi));
    prof_mark(ctx, 4);
    return SVGDB_OK;
}

int one_step(svgdb_ctx *ctx);

// Reads the verdict of the last optimistic step (by now the step has been enqueued in full; usually it has long finished) and, if
// its predicted bracket missed the median, repeats that step with the synchronous median: nothing persistent was changed by the
// missed attempt (the optimizer update is predicated on the device-side flag), its X_next is simply overwritten.
int settle_step(svgdb_ctx *ctx)
{
    if (!ctx->pending_verify) return SVGDB_OK;
    TRY(finish_median(ctx)); // waits for the verdict; on a hit also books the median into the prediction history
    ctx->pending_verify = false;
    if (ctx->hs->dec.hit != 0) return SVGDB_OK;
    // miss: undo the bookkeeping of the attempt and run the step again, this time waiting for every pass's counts
    CU(cudaMemsetAsync(ctx->miss_dev, 0, sizeof(int), ctx->stream));
    ctx->cur ^= 1;
    --ctx->counter;
    --ctx->stats.iterations;
    ctx->delta = std::min(bracket_delta_max(ctx, (double)ctx->N * (double)ctx->N), std::max(ctx->delta, 2e-5) * 4.0);
    ctx->miss_boost = std::min(64.0, ctx->miss_boost * 4.0);
    const bool keep = ctx->opt_suspended;
    ctx->opt_suspended = true;
    const int rc = one_step(ctx);
    ctx->opt_suspended = keep;
    return rc;
}

int one_step(svgdb_ctx *ctx)
{
    TRY(settle_step(ctx));
    ++ctx->counter; // Adam increments its counter before the bias correction (Adam.hpp:80-82)
    if (ctx->opt.kind == OPT_ADAM) {
        ctx->opt.bias1 = 1.0 - std::pow(ctx->opt.beta1, (double)ctx->counter);
        ctx->opt.bias2 = 1.0 - std::pow(ctx->opt.beta2, (double)ctx->counter);
    }
    {
        int rc = prepare_and_phi(ctx, false);
        if (rc == SVGDB_OK) rc = allgather_rows(ctx, ctx->X[ctx->cur ^ 1], ctx->d);
        if (rc != SVGDB_OK) { // the step was not enqueued: it does not count (Adam's bias correction depends on the counter)
            --ctx->counter;
            return rc;
        }
    }
    ctx->cur ^= 1;
    ++ctx->stats.iterations;
    if (ctx->profiling) {
        cudaEvent_t end;
        CU(cudaEventCreate(&end));
        CU(cudaEventRecord(end, ctx->stream));
        CU(cudaEventSynchronize(end));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->stats.ms_median += ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->stats.ms_grad += ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->stats.ms_comm += ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); ctx->stats.ms_phi += ms;
        if (ctx->n_rows > 0 && cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]) == cudaSuccess) ctx->stats.ms_phi_kernel += ms;
        if (ctx->n_rows > 0 && ctx->scale_method != SVGDB_SCALE_HESSIAN && cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]) == cudaSuccess) ctx->stats.ms_grad_kernel += ms;
        cudaGetLastError();
        cudaEventElapsedTime(&ms, ctx->ev[4], end); ctx->stats.ms_comm += ms;
        cudaEventDestroy(end);
    }
    return SVGDB_OK;
}

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char *svgdb_version(void) { return "svgd_b200 0.1 (sm_100a)"; }

const char *svgdb_last_error(const svgdb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int svgdb_create(svgdb_ctx **out, int device, int64_t n_total, int32_t d, int precision_mode)
{
    if (!out) return SVGDB_ERR_INVALID;
    *out = nullptr;
    if (n_total < 1 || d < 1) return SVGDB_ERR_DIMENSION;
    svgdb_ctx *ctx = new (std::nothrow) svgdb_ctx();
    if (!ctx) return SVGDB_ERR_NOMEM;
    *out = ctx; // returned even on failure so the caller can read svgdb_last_error, then destroy
    ctx->device = device;
    ctx->N = n_total;
    ctx->d = d;
    ctx->precision = precision_mode;
    if (precision_mode != SVGDB_PRECISION_F64 && precision_mode != SVGDB_PRECISION_TC32)
        return fail(ctx, SVGDB_ERR_INVALID, "unknown precision mode");
#ifndef SVGDB_WITH_TC32
    if (precision_mode == SVGDB_PRECISION_TC32)
        return fail(ctx, SVGDB_ERR_INVALID, "this build has no TC32 (tcgen05) path");
#endif
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, SVGDB_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                             " (this library has no CPU fallback)");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ctx, SVGDB_ERR_CUDA, "built for sm_100a (B200); found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
    ctx->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (auto &ev : ctx->ev) CU(cudaEventCreate(&ev));
    CU(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto &e : ctx->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : ctx->ev_up) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_med, cudaEventDisableTiming));
    {
        // candidate buffer of the median select: 2^25 keys up to N = 46K, then N^2 / 64 keys (the bracket that fits it keeps
        // the same relative width: it must absorb the extrapolation error of the median), at most 2^31 keys (16 GiB)
        const long double want = (long double)n_total * (long double)n_total / 64.0L;
        if (want > (long double)ctx->capacity) ctx->capacity = (uint64_t)std::min<long double>(want, 2147483648.0L);
    }
    if (const char *s = std::getenv("SVGDB_CAND_CAPACITY")) {
        long long v = std::atoll(s);
        if (v >= 64) ctx->capacity = (uint64_t)v;
    }
    // opt-in to large dynamic shared memory for every instantiation we launch
    CU(cudaFuncSetAttribute(phi_f64_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)phi_smem_bytes(d, 8)));
    CU(cudaFuncSetAttribute(phi_f64_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)phi_smem_bytes(d, 16)));
    CU(cudaFuncSetAttribute(phi_f64_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)phi_smem_bytes(d, 32)));
    CU(cudaFuncSetAttribute(phi_f64_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)phi_smem_bytes(d, 64)));
    CU(cudaFuncSetAttribute(dist_pass_f64_kernel<MODE_HIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dist_smem_bytes(d, true)));
    CU(cudaFuncSetAttribute(dist_pass_f64_kernel<MODE_COLLECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dist_smem_bytes(d, false)));
    {
        size_t gsm = ((size_t)3 * 16 * d + 5 * 16) * sizeof(double);
        if (gsm > (size_t)prop.sharedMemPerBlockOptin)
            return fail(ctx, SVGDB_ERR_DIMENSION, "dimension too large for the built-in Gaussian gradient kernel (d <= 590)");
        CU(cudaFuncSetAttribute(mvn_sum_grad_f64_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    }
#ifdef SVGDB_WITH_TC32
    if (precision_mode == SVGDB_PRECISION_TC32) {
        if (d > 256)
            return fail(ctx, SVGDB_ERR_DIMENSION, "SVGDB_PRECISION_TC32 supports d <= 256 (the row operand and the accumulator of a 128-particle tile must fit the 512 TMEM columns); use SVGDB_PRECISION_F64");
        if (d > svgdb::tc::TC_D) {
            using namespace svgdb::tc;
#define SVGDB_WIDE_ATTR(DPV)                                                                                                                              \
    CU(cudaFuncSetAttribute(phiw_tc32_kernel<DPV, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWCfg<DPV, false>::SMEM));                 \
    CU(cudaFuncSetAttribute(phiw_tc32_kernel<DPV, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWCfg<DPV, true>::SMEM));                   \
    CU(cudaFuncSetAttribute(phiw_tc32_kernel<DPV, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWCfg<DPV, false>::SMEM));                 \
    CU(cudaFuncSetAttribute(phiw_tc32_kernel<DPV, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWCfg<DPV, true>::SMEM));                   \
    CU(cudaFuncSetAttribute(distw_tc32_kernel<DPV, MODE_HIST, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DWCfg<DPV>::SMEM));        \
    CU(cudaFuncSetAttribute(distw_tc32_kernel<DPV, MODE_COLLECT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DWCfg<DPV>::SMEM));     \
    CU(cudaFuncSetAttribute(distw_tc32_kernel<DPV, MODE_COLLECT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DWCfg<DPV>::SMEM));
            SVGDB_WIDE_ATTR(128)
            SVGDB_WIDE_ATTR(192)
            SVGDB_WIDE_ATTR(256)
#undef SVGDB_WIDE_ATTR
        }
#define SVGDB_PHI2_ATTR(P)                                                                                                                        \
    CU(cudaFuncSetAttribute(svgdb::tc::phi2_tc32_kernel<P, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::P2Cfg<false>::SMEM)); \
    CU(cudaFuncSetAttribute(svgdb::tc::phi2_tc32_kernel<P, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::P2Cfg<true>::SMEM));   \
    CU(cudaFuncSetAttribute(svgdb::tc::phi2_tc32_kernel<P, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::P2Cfg<false>::SMEM)); \
    CU(cudaFuncSetAttribute(svgdb::tc::phi2_tc32_kernel<P, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::P2Cfg<true>::SMEM));   \
    CU(cudaFuncSetAttribute(svgdb::tc::phi2_tc32_kernel<P, false, 1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::P2Cfg<false, true>::SMEM)); \
    CU(cudaFuncSetAttribute(svgdb::tc::phi2_tc32_kernel<P, false, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::P2Cfg<false, true>::SMEM));
        SVGDB_PHI2_ATTR(0)
        SVGDB_PHI2_ATTR(1)
        SVGDB_PHI2_ATTR(2)
        SVGDB_PHI2_ATTR(4)
#undef SVGDB_PHI2_ATTR
#define SVGDB_D2_ATTR(M, G, F) \
    CU(cudaFuncSetAttribute(svgdb::tc::dist2_tc32_kernel<M, G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::D2_SMEM));
        SVGDB_D2_ATTR(MODE_HIST, false, false)
        SVGDB_D2_ATTR(MODE_HIST, true, false)
        SVGDB_D2_ATTR(MODE_COLLECT, false, false)
        SVGDB_D2_ATTR(MODE_COLLECT, true, false)
        SVGDB_D2_ATTR(MODE_COLLECT, false, true)
        SVGDB_D2_ATTR(MODE_COLLECT, true, true)
#undef SVGDB_D2_ATTR
        CU(cudaFuncSetAttribute(svgdb::tc::dist2_tc32_kernel<MODE_COLLECT, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::D2_SMEM));
        CU(cudaFuncSetAttribute(svgdb::tc::dist2_tc32_kernel<MODE_COLLECT, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svgdb::tc::D2_SMEM));
    }
#endif
    CU(cudaMalloc(&ctx->a_dev, sizeof(double)));
    CU(cudaMalloc(&ctx->pass_words, 4 * 8));
    ctx->below = ctx->pass_words;
    ctx->cand_total = ctx->pass_words + 1;
    ctx->max_below = ctx->pass_words + 2;
    CU(cudaMalloc(&ctx->cand_count, 8));
    CU(cudaMalloc(&ctx->hist, HIST_BINS * 8));
    CU(cudaMalloc(&ctx->sel, sizeof(SelectState)));
    CU(cudaMalloc(&ctx->dec_dev, sizeof(StepDecision)));
    CU(cudaMalloc(&ctx->miss_dev, sizeof(int)));
    CU(cudaMemsetAsync(ctx->dec_dev, 0, sizeof(StepDecision), ctx->stream));
    CU(cudaMemsetAsync(ctx->miss_dev, 0, sizeof(int), ctx->stream));
    CU(cudaMalloc(&ctx->medres, sizeof(MedianResult)));
    CU(cudaMallocHost(&ctx->hs, sizeof(HostScratch)));
    {
        // the candidate buffer never needs more than n^2 entries
        long double tot = (long double)n_total * (long double)n_total;
        if (tot < (long double)ctx->capacity) ctx->capacity = std::max<uint64_t>(64, (uint64_t)tot);
        CU(cudaMalloc(&ctx->cand, ctx->capacity * 8));
    }
    TRY(alloc_sharded(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    return SVGDB_OK;
}

void svgdb_destroy(svgdb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->comm) nccl().CommDestroy(ctx->comm);
    free_sharded(ctx);
    cudaFree(ctx->a_dev); cudaFree(ctx->lb); cudaFree(ctx->ub); cudaFree(ctx->means_dev); cudaFree(ctx->prec_dev);
    cudaFree(ctx->gg_D); cudaFree(ctx->gg_Y); cudaFree(ctx->gg_ms);
    if (ctx->blas) cublas().Destroy(ctx->blas);
    cudaFree(ctx->Hsum_dev); cudaFree(ctx->Wsum_dev); cudaFree(ctx->R_dev); cudaFree(ctx->Rt_dev); cudaFree(ctx->Rinv_dev); cudaFree(ctx->Y_dev); cudaFree(ctx->GH_dev);
    cudaFree(ctx->pass_words); cudaFree(ctx->cand_count); cudaFree(ctx->hist); cudaFree(ctx->cand);
    cudaFree(ctx->sel); cudaFree(ctx->medres); cudaFree(ctx->dec_dev); cudaFree(ctx->miss_dev);
    if (ctx->hs) cudaFreeHost(ctx->hs);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_med) cudaEventDestroy(ctx->ev_med);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto &e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_up) if (e) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int svgdb_device_count(int *count)
{
    if (!count) return SVGDB_ERR_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return SVGDB_ERR_CUDA;
    }
    *count = n;
    return SVGDB_OK;
}

int svgdb_set_stream(svgdb_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SVGDB_OK;
}

int svgdb_nccl_unique_id(void *out_id, size_t bytes)
{
    if (!out_id || bytes < sizeof(ncclUniqueId)) return SVGDB_ERR_INVALID;
    ncclUniqueId id;
    if (!nccl().ok || nccl().GetUniqueId(&id) != ncclSuccess) return SVGDB_ERR_NCCL;
    std::memcpy(out_id, &id, sizeof(id));
    return SVGDB_OK;
}

int svgdb_comm_init(svgdb_ctx *ctx, int world, int rank, const void *nccl_unique_id, size_t bytes)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, SVGDB_ERR_INVALID, "bad world/rank");
    if (world > ctx->N) return fail(ctx, SVGDB_ERR_DIMENSION, "more ranks than particles");
    CU(cudaSetDevice(ctx->device));
    if (ctx->comm) { nccl().CommDestroy(ctx->comm); ctx->comm = nullptr; }
    if (world > 1) {
        if (!nccl().ok) return fail(ctx, SVGDB_ERR_NCCL, nccl().why);
        if (!nccl_unique_id || bytes < sizeof(ncclUniqueId)) return fail(ctx, SVGDB_ERR_INVALID, "missing ncclUniqueId");
        ncclUniqueId id;
        std::memcpy(&id, nccl_unique_id, sizeof(id));
        CU(cudaSetDevice(ctx->device));
        NC(nccl().CommInitRank(&ctx->comm, world, id, rank));
    }
    ctx->world = world;
    ctx->rank = rank;
    ctx->n_hist = 0;
    TRY(alloc_sharded(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    return SVGDB_OK;
}

int svgdb_set_particles(svgdb_ctx *ctx, const double *X)
{
    if (!ctx || !X) return fail(ctx, SVGDB_ERR_INVALID, "null particle matrix");
    CU(cudaSetDevice(ctx->device)); // one process may drive several contexts (one per GPU): every entry point selects its own
    CU(cudaMemcpyAsync(ctx->X[ctx->cur], X, (size_t)ctx->N * ctx->d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // the caller may reuse or free X as soon as this returns, also when X is pinned memory (svgdb_host_alloc)
    CU(cudaStreamSynchronize(ctx->stream));
    // the median bracket prediction is only a hint (verified by exact counts), so it survives re-uploads
    return SVGDB_OK;
}

int svgdb_get_particles(svgdb_ctx *ctx, double *X)
{
    if (!ctx || !X) return fail(ctx, SVGDB_ERR_INVALID, "null particle matrix");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(X, ctx->X[ctx->cur], (size_t)ctx->N * ctx->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) TRY(check_tc_err(ctx));
#endif
    return SVGDB_OK;
}

int svgdb_local_rows(svgdb_ctx *ctx, int64_t *row0, int64_t *n_rows)
{
    if (!ctx || !row0 || !n_rows) return SVGDB_ERR_INVALID;
    *row0 = ctx->row0;
    *n_rows = ctx->n_rows;
    return SVGDB_OK;
}

} // extern "C"

namespace {
// enqueue only: svgdb_step_host returns after a synchronize of its own, so the caller's buffer is read before it gets control back
int set_particles_rows_async(svgdb_ctx *ctx, const double *rows_local)
{
    if (ctx->n_rows > 0)
        CU(cudaMemcpyAsync(ctx->X[ctx->cur] + (size_t)ctx->row0 * ctx->d, rows_local, (size_t)ctx->n_rows * ctx->d * sizeof(double),
                           cudaMemcpyHostToDevice, ctx->stream));
    return allgather_rows(ctx, ctx->X[ctx->cur], ctx->d); // every rank needs all particles: NVLink instead of N x PCIe
}
} // namespace

extern "C" {

int svgdb_set_particles_rows(svgdb_ctx *ctx, const double *rows_local)
{
    if (!ctx || (!rows_local && ctx->n_rows > 0)) return fail(ctx, SVGDB_ERR_INVALID, "null particle rows");
    CU(cudaSetDevice(ctx->device));
    TRY(set_particles_rows_async(ctx, rows_local));
    CU(cudaStreamSynchronize(ctx->stream)); // rows_local may be reused or freed on return
    return SVGDB_OK;
}

int svgdb_get_particles_rows(svgdb_ctx *ctx, double *rows_local)
{
    if (!ctx || (!rows_local && ctx->n_rows > 0)) return fail(ctx, SVGDB_ERR_INVALID, "null particle rows");
    CU(cudaSetDevice(ctx->device));
    if (ctx->n_rows > 0)
        CU(cudaMemcpyAsync(rows_local, ctx->X[ctx->cur] + (size_t)ctx->row0 * ctx->d, (size_t)ctx->n_rows * ctx->d * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) TRY(check_tc_err(ctx));
#endif
    return SVGDB_OK;
}

int svgdb_set_model_mvn_sum(svgdb_ctx *ctx, int32_t C, const double *means, const double *covs)
{
    if (!ctx || !means || !covs) return fail(ctx, SVGDB_ERR_INVALID, "null model parameters");
    if (C < 1) return fail(ctx, SVGDB_ERR_INVALID, "a mixture needs at least one component");
    CU(cudaSetDevice(ctx->device));
    const int d = ctx->d;
    std::vector<double> prec((size_t)C * d * d), inv;
    for (int c = 0; c < C; ++c) {
        if (!invert_matrix(covs + (size_t)c * d * d, d, inv))
            return fail(ctx, SVGDB_ERR_NUMERIC, "covariance of component " + std::to_string(c) + " is singular");
        std::copy(inv.begin(), inv.end(), prec.begin() + (size_t)c * d * d);
    }
    cudaFree(ctx->means_dev); cudaFree(ctx->prec_dev); cudaFree(ctx->Wsum_dev);
    ctx->means_dev = ctx->prec_dev = ctx->Wsum_dev = nullptr;
    ctx->hess_const_valid = false;
    CU(cudaMalloc(&ctx->means_dev, (size_t)C * d * sizeof(double)));
    CU(cudaMalloc(&ctx->prec_dev, prec.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->means_dev, means, (size_t)C * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->prec_dev, prec.data(), prec.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // prec is a local
    ctx->C = C;
    ctx->model_kind = MODEL_MVN_SUM;
    ctx->prec_host = prec;
    ctx->means_host.assign(means, means + (size_t)C * d);
    return SVGDB_OK;
}

int svgdb_set_model_mvn(svgdb_ctx *ctx, const double *mean, const double *cov)
{
    return svgdb_set_model_mvn_sum(ctx, 1, mean, cov);
}

int svgdb_set_model_device_hook(svgdb_ctx *ctx, svgdb_grad_fn fn, void *user)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if (!fn) return fail(ctx, SVGDB_ERR_UNSET, "Model function is unset.");
    ctx->hook = fn;
    ctx->hook_user = user;
    ctx->model_kind = MODEL_HOOK;
    return SVGDB_OK;
}

int svgdb_set_kernel_rbf(svgdb_ctx *ctx, int scale_method, double fixed_a)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if (scale_method != SVGDB_SCALE_MEDIAN && scale_method != SVGDB_SCALE_FIXED && scale_method != SVGDB_SCALE_HESSIAN)
        return fail(ctx, SVGDB_ERR_INVALID, "[Argument error] Invalid scale method Enum provided.");
    if (scale_method == SVGDB_SCALE_FIXED && !(fixed_a > 0.0) ) return fail(ctx, SVGDB_ERR_INVALID, "fixed kernel scale must be positive");
    ctx->scale_method = scale_method;
    ctx->fixed_a = fixed_a;
    ctx->kernel_set = true;
    return SVGDB_OK;
}

int svgdb_set_tc32_variant(svgdb_ctx *ctx, int variant)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if (variant != SVGDB_TC32_AUTO && variant != SVGDB_TC32_FAST && variant != SVGDB_TC32_PRECISE)
        return fail(ctx, SVGDB_ERR_INVALID, "unknown TC32 variant");
    ctx->tc32_variant = variant;
    return SVGDB_OK;
}

int svgdb_set_optimizer(svgdb_ctx *ctx, int kind, double lr, double beta1, double beta2, double eps)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if (kind == SVGDB_OPT_ADAM) {
        if (beta1 >= 1.0 || beta1 < 0.0 || beta2 >= 1.0 || beta2 < 0.0)
            return fail(ctx, SVGDB_ERR_INVALID, "[Argument Error] Invalid value for decay parameter beta.");
    } else if (kind == SVGDB_OPT_RMSPROP) {
        if (beta1 > 1.0 || beta1 < 0.0) return fail(ctx, SVGDB_ERR_INVALID, "[Argument Error] Invalid value for decay parameter beta.");
    } else if (kind != SVGDB_OPT_ADAGRAD) {
        return fail(ctx, SVGDB_ERR_INVALID, "unknown optimizer kind");
    }
    ctx->opt.kind = kind;
    ctx->opt.lr = lr;
    ctx->opt.beta1 = beta1;
    ctx->opt.beta2 = beta2;
    ctx->opt.eps = eps;
    ctx->opt.bias1 = ctx->opt.bias2 = 1.0;
    ctx->opt_set = true;
    return SVGDB_OK;
}

int svgdb_set_bounds(svgdb_ctx *ctx, const double *lb, const double *ub, int32_t n_bound)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaFree(ctx->lb); cudaFree(ctx->ub);
    ctx->lb = ctx->ub = nullptr;
    if (!lb && !ub) return SVGDB_OK;
    if (!lb || !ub) return fail(ctx, SVGDB_ERR_INVALID, "both bounds are required");
    if (n_bound != 1 && n_bound != ctx->d) return fail(ctx, SVGDB_ERR_DIMENSION, "The provided bounds have incorrect dimensions.");
    std::vector<double> l(ctx->d), u(ctx->d);
    for (int k = 0; k < ctx->d; ++k) { l[k] = lb[n_bound == 1 ? 0 : k]; u[k] = ub[n_bound == 1 ? 0 : k]; }
    CU(cudaMalloc(&ctx->lb, ctx->d * sizeof(double)));
    CU(cudaMalloc(&ctx->ub, ctx->d * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->lb, l.data(), ctx->d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->ub, u.data(), ctx->d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SVGDB_OK;
}

int svgdb_initialize(svgdb_ctx *ctx)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    TRY(check_ready(ctx));
    if (!ctx->opt_set) return fail(ctx, SVGDB_ERR_UNSET, "Optimizer is unset.");
    size_t local = (size_t)ctx->rows_per_rank * ctx->d * sizeof(double);
    CU(cudaMemsetAsync(ctx->s1, 0, local, ctx->stream));
    CU(cudaMemsetAsync(ctx->s2, 0, local, ctx->stream));
    ctx->counter = 0;
    ctx->initialized = true;
    // a new run: the medians of the previous one do not predict this one's
    TRY(finish_median(ctx));
    ctx->n_hist = 0;
    ctx->resid[0] = ctx->resid[1] = 0.0;
    ctx->last_pred = 0.0;
    return SVGDB_OK;
}

int svgdb_step(svgdb_ctx *ctx, int64_t iters)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    TRY(check_ready(ctx));
    if (!ctx->opt_set) return fail(ctx, SVGDB_ERR_UNSET, "Optimizer is unset.");
    CU(cudaSetDevice(ctx->device));
    for (int64_t it = 0; it < iters; ++it) TRY(one_step(ctx));
    return settle_step(ctx); // the last step's bracket verdict (repairs the step if it missed); everything else stays asynchronous
}

int svgdb_step_host(svgdb_ctx *ctx, const double *rows_in, double *rows_out, int64_t iters)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if ((!rows_in || !rows_out) && ctx->n_rows > 0) return fail(ctx, SVGDB_ERR_INVALID, "null particle rows");
    if (iters < 1) return fail(ctx, SVGDB_ERR_INVALID, "svgdb_step_host needs at least one iteration");
    TRY(check_ready(ctx));
    if (!ctx->opt_set) return fail(ctx, SVGDB_ERR_UNSET, "Optimizer is unset.");
    CU(cudaSetDevice(ctx->device));
    cudaPointerAttributes attr{};
    const bool in_pinned = rows_in && cudaPointerGetAttributes(&attr, rows_in) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const int n_ipairs_all = (int)((ctx->N + 255) / 256);
    bool chunked_upload = false;
#ifdef SVGDB_WITH_TC32
    // One rank, tensor-core path, median scale: the first thing a step does is the distance pass over all pairs, and the pairs
    // among the rows that have already arrived can be counted while the rest is still on the PCIe bus.  The particles go up in
    // four row chunks on the copy stream; the first distance pass of the step is issued chunk by chunk behind them.
    chunked_upload = in_pinned && ctx->world == 1 && ctx->precision == SVGDB_PRECISION_TC32 && ctx->scale_method == SVGDB_SCALE_MEDIAN &&
                     n_ipairs_all >= 32 && ctx->host_chunks != 0 && !ctx->wide;
#endif
    if (chunked_upload) {
        CU(cudaEventRecord(ctx->ev_fork, ctx->stream)); // whatever still reads or writes X[cur] on the main stream comes first
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
        int64_t r0 = 0;
        upload_chunk_ends(n_ipairs_all, ctx->up_ipairs);
        for (int ch = 0; ch < 4; ++ch) {
            const int64_t r1 = std::min<int64_t>((int64_t)ctx->up_ipairs[ch] * 256, ctx->N);
            CU(cudaMemcpyAsync(ctx->X[ctx->cur] + (size_t)r0 * ctx->d, rows_in + (size_t)r0 * ctx->d, (size_t)(r1 - r0) * ctx->d * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaEventRecord(ctx->ev_up[ch], ctx->copy_stream));
            r0 = r1;
        }
        ctx->up_chunks = 4;
    } else {
        TRY(set_particles_rows_async(ctx, rows_in));
    }
    for (int64_t it = 0; it + 1 < iters; ++it) {
        const int rc = one_step(ctx);
        if (rc != SVGDB_OK) {
            ctx->up_chunks = 0;
            cudaStreamSynchronize(ctx->copy_stream);
            return rc;
        }
    }
    // The last iteration hands its rows to the host as they are finished -- from pinned memory only: a copy into pageable
    // memory blocks the launching thread, which would hold back the launches of the remaining row chunks.
    cudaPointerAttributes attr_out{};
    const bool pinned = rows_out && cudaPointerGetAttributes(&attr_out, rows_out) == cudaSuccess && attr_out.type == cudaMemoryTypeHost;
    cudaGetLastError();
    ctx->stream_out = pinned && ctx->host_chunks != 0 ? rows_out : nullptr;
    ctx->streamed = false;
    int rc = one_step(ctx);
    ctx->stream_out = nullptr;
    ctx->up_chunks = 0;
    if (rc == SVGDB_OK && ctx->pending_verify) {
        CU(cudaStreamSynchronize(ctx->copy_stream)); // the streamed rows of a missed attempt must have landed before the repair's rows are copied over them
        rc = settle_step(ctx);
        if (rc == SVGDB_OK && ctx->hs->dec.hit == 0) ctx->streamed = false; // repaired: fetch the rows again below
    }
    if (rc != SVGDB_OK) {
        cudaStreamSynchronize(ctx->copy_stream); // nothing may still be reading or writing the caller's buffers
        return rc;
    }
    if (ctx->streamed) {
        CU(cudaStreamSynchronize(ctx->copy_stream));
        CU(cudaStreamSynchronize(ctx->stream));
#ifdef SVGDB_WITH_TC32
        TRY(check_tc_err(ctx));
#endif
        return SVGDB_OK;
    }
    return svgdb_get_particles_rows(ctx, rows_out);
}

int svgdb_compute_phi(svgdb_ctx *ctx, double *phi, double *scale_out)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    TRY(check_ready(ctx));
    CU(cudaSetDevice(ctx->device));
    {
        const bool keep = ctx->opt_suspended;
        ctx->opt_suspended = true; // inspection: the scale is read back right away
        const int rc = prepare_and_phi(ctx, true);
        ctx->opt_suspended = keep;
        TRY(rc);
    }
    if (phi) {
        // gather the local rows of every rank through V's storage (phi is debug output)
        double *tmp = ctx->V;
        if (ctx->n_rows > 0)
            CU(cudaMemcpyAsync(tmp + (size_t)ctx->row0 * ctx->d, ctx->phi_dbg, (size_t)ctx->n_rows * ctx->d * sizeof(double),
                               cudaMemcpyDeviceToDevice, ctx->stream));
        TRY(allgather_rows(ctx, tmp, ctx->d));
        CU(cudaMemcpyAsync(phi, tmp, (size_t)ctx->N * ctx->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    double a = 0.0;
    CU(cudaMemcpyAsync(&a, ctx->a_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (scale_out) *scale_out = a;
    return SVGDB_OK;
}

int svgdb_compute_scale(svgdb_ctx *ctx, double *scale_out)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    if (!ctx->kernel_set) return fail(ctx, SVGDB_ERR_UNSET, "Kernel function is unset.");
    CU(cudaSetDevice(ctx->device));
    if (ctx->scale_method == SVGDB_SCALE_HESSIAN) { // matrix-valued: A(0,0) here, the whole matrix through svgdb_get_scale_matrix
        TRY(hessian_scale_dev(ctx));
        if (scale_out) *scale_out = ctx->A_host[0];
        return SVGDB_OK;
    }
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) TRY(launch_dist_operands(ctx));
#endif
    if (ctx->precision != SVGDB_PRECISION_TC32) TRY(launch_rownorm(ctx));
    {
        const bool keep = ctx->opt_suspended;
        ctx->opt_suspended = true; // inspection: the scale is read back right away
        const int rc = compute_scale_dev(ctx);
        ctx->opt_suspended = keep;
        TRY(rc);
    }
    double a = 0.0;
    CU(cudaMemcpyAsync(&a, ctx->a_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (scale_out) *scale_out = a;
    return SVGDB_OK;
}

int svgdb_get_scale_matrix(svgdb_ctx *ctx, double *A_dxd)
{
    if (!ctx || !A_dxd) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (ctx->scale_method != SVGDB_SCALE_HESSIAN) {
        TRY(finish_median(ctx));
        const double a = ctx->scale_method == SVGDB_SCALE_FIXED ? ctx->fixed_a : ctx->stats.last_scale;
        for (int r = 0; r < ctx->d; ++r)
            for (int c = 0; c < ctx->d; ++c) A_dxd[(size_t)r * ctx->d + c] = r == c ? a : 0.0;
        return SVGDB_OK;
    }
    if (ctx->A_host.size() != (size_t)ctx->d * ctx->d) return fail(ctx, SVGDB_ERR_UNSET, "the kernel scale has not been computed yet");
    std::copy(ctx->A_host.begin(), ctx->A_host.end(), A_dxd);
    return SVGDB_OK;
}

int svgdb_compute_log_model_grad(svgdb_ctx *ctx, double *G)
{
    if (!ctx || !G) return fail(ctx, SVGDB_ERR_INVALID, "null output");
    if (ctx->model_kind == MODEL_UNSET) return fail(ctx, SVGDB_ERR_UNSET, "Model function is unset.");
    CU(cudaSetDevice(ctx->device));
    TRY(launch_grad(ctx, ctx->stream));
    double *tmp = ctx->V;
    if (ctx->n_rows > 0)
        CU(cudaMemcpyAsync(tmp + (size_t)ctx->row0 * ctx->d, ctx->G, (size_t)ctx->n_rows * ctx->d * sizeof(double),
                           cudaMemcpyDeviceToDevice, ctx->stream));
    TRY(allgather_rows(ctx, tmp, ctx->d));
    CU(cudaMemcpyAsync(G, tmp, (size_t)ctx->N * ctx->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SVGDB_OK;
}

int svgdb_compute_log_model(svgdb_ctx *ctx, double *logp)
{
    if (!ctx || !logp) return fail(ctx, SVGDB_ERR_INVALID, "null output");
    if (ctx->model_kind == MODEL_UNSET) return fail(ctx, SVGDB_ERR_UNSET, "Model function is unset.");
    if (ctx->model_kind != MODEL_MVN_SUM)
        return fail(ctx, SVGDB_ERR_UNSET, "a model behind the device-gradient hook has no value function on the device");
    CU(cudaSetDevice(ctx->device));
    double *tmp = ctx->V; // N x d doubles: room for N values at any row offset
    if (ctx->n_rows > 0) {
        mvn_sum_logp_f64_kernel<<<(unsigned)((ctx->n_rows + 127) / 128), 128, 0, ctx->stream>>>(ctx->X[ctx->cur], ctx->d, ctx->row0, ctx->n_rows, ctx->C,
                                                                                           ctx->means_dev, ctx->prec_dev, tmp + ctx->row0);
        KERNEL_CHECK();
    }
    TRY(allgather_rows(ctx, tmp, 1));
    CU(cudaMemcpyAsync(logp, tmp, (size_t)ctx->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SVGDB_OK;
}

int svgdb_compute_kernel_matrices(svgdb_ctx *ctx, double *K, double *gradK, double *scale_out)
{
    if (!ctx || !K || !gradK) return fail(ctx, SVGDB_ERR_INVALID, "null output");
    if (!ctx->kernel_set) return fail(ctx, SVGDB_ERR_UNSET, "Kernel function is unset.");
    if (ctx->world > 1) return fail(ctx, SVGDB_ERR_INVALID, "the kernel matrices are an inspection aid for one rank");
    const int64_t n = ctx->N;
    const int d = ctx->d;
    const double bytes = (double)n * (double)n * (double)(d + 1) * 8.0;
    if (bytes > 2147483648.0)
        return fail(ctx, SVGDB_ERR_DIMENSION, "kernel matrices of " + std::to_string(n) + " particles in " + std::to_string(d) +
                                                   " dimensions exceed the 2 GiB limit of this inspection path");
    double a = 0.0;
    TRY(svgdb_compute_scale(ctx, &a)); // the scale Kernel::Step would compute from the current particles
    std::vector<double> A((size_t)d * d, 0.0);
    if (ctx->scale_method == SVGDB_SCALE_HESSIAN) A = ctx->A_host;
    else
        for (int r = 0; r < d; ++r) A[(size_t)r * d + r] = a;
    double *A_dev = nullptr, *K_dev = nullptr, *dK_dev = nullptr;
    int rc = SVGDB_OK;
    do {
        if (cudaMalloc(&A_dev, A.size() * 8) != cudaSuccess || cudaMalloc(&K_dev, (size_t)n * n * 8) != cudaSuccess ||
            cudaMalloc(&dK_dev, (size_t)n * n * d * 8) != cudaSuccess) {
            rc = fail(ctx, SVGDB_ERR_NOMEM, "out of device memory for the kernel matrices");
            break;
        }
        if (cudaMemcpyAsync(A_dev, A.data(), A.size() * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { rc = fail(ctx, SVGDB_ERR_CUDA, "copy failed"); break; }
        kernel_matrices_f64_kernel<<<(unsigned)((n * n + 127) / 128), 128, 0, ctx->stream>>>(ctx->X[ctx->cur], n, d, A_dev, K_dev, dK_dev);
        if (cudaGetLastError() != cudaSuccess) { rc = fail(ctx, SVGDB_ERR_CUDA, "kernel_matrices_f64_kernel launch failed"); break; }
        cudaMemcpyAsync(K, K_dev, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(gradK, dK_dev, (size_t)n * n * d * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fail(ctx, SVGDB_ERR_CUDA, "kernel matrices: " + std::string(cudaGetErrorString(cudaGetLastError()))); break; }
    } while (false);
    cudaFree(A_dev); cudaFree(K_dev); cudaFree(dK_dev);
    if (rc == SVGDB_OK && scale_out) *scale_out = a;
    return rc;
}

int svgdb_get_opt_state(svgdb_ctx *ctx, double *s1, double *s2, uint64_t *counter)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    // optimizer state is row-sharded; gather through V's storage one array at a time
    double *src[2] = {ctx->s1, ctx->s2};
    double *dst[2] = {s1, s2};
    for (int k = 0; k < 2; ++k) {
        if (!dst[k]) continue;
        if (ctx->n_rows > 0)
            CU(cudaMemcpyAsync(ctx->V + (size_t)ctx->row0 * ctx->d, src[k], (size_t)ctx->n_rows * ctx->d * sizeof(double),
                               cudaMemcpyDeviceToDevice, ctx->stream));
        TRY(allgather_rows(ctx, ctx->V, ctx->d));
        CU(cudaMemcpyAsync(dst[k], ctx->V, (size_t)ctx->N * ctx->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (counter) *counter = ctx->counter;
    return SVGDB_OK;
}

int svgdb_set_opt_state(svgdb_ctx *ctx, const double *s1, const double *s2, uint64_t counter)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)std::max<int64_t>(ctx->n_rows, 0) * ctx->d * sizeof(double);
    if (s1 && bytes) CU(cudaMemcpyAsync(ctx->s1, s1 + (size_t)ctx->row0 * ctx->d, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (s2 && bytes) CU(cudaMemcpyAsync(ctx->s2, s2 + (size_t)ctx->row0 * ctx->d, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->counter = counter;
    return SVGDB_OK;
}

int svgdb_sync(svgdb_ctx *ctx)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    TRY(finish_median(ctx));
#ifdef SVGDB_WITH_TC32
    if (ctx->precision == SVGDB_PRECISION_TC32) TRY(check_tc_err(ctx));
#endif
    return SVGDB_OK;
}

int svgdb_set_profiling(svgdb_ctx *ctx, int enabled)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    ctx->profiling = enabled != 0;
    return SVGDB_OK;
}

int svgdb_get_stats(svgdb_ctx *ctx, svgdb_stats *out)
{
    if (!ctx || !out) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    TRY(finish_median(ctx)); // stats.last_scale
    *out = ctx->stats;
    return SVGDB_OK;
}

int svgdb_reset_stats(svgdb_ctx *ctx)
{
    if (!ctx) return SVGDB_ERR_INVALID;
    double last = ctx->stats.last_scale;
    ctx->stats = svgdb_stats{};
    ctx->stats.last_scale = last;
    return SVGDB_OK;
}

int svgdb_time_steps(svgdb_ctx *ctx, int64_t iters, float *ms_out)
{
    if (!ctx || !ms_out) return SVGDB_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventRecord(e0, ctx->stream));
    int rc = svgdb_step(ctx, iters);
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaEventSynchronize(e1));
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

int svgdb_host_alloc(void **out, size_t bytes)
{
    if (!out) return SVGDB_ERR_INVALID;
    return cudaMallocHost(out, bytes) == cudaSuccess ? SVGDB_OK : SVGDB_ERR_NOMEM;
}

int svgdb_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SVGDB_OK : SVGDB_ERR_CUDA; }

int svgdb_time_kernel(svgdb_ctx *ctx, int which, int reps, int variant, double rel_halfwidth, float *ms_out)
{
    if (!ctx || !ms_out || reps < 1) return SVGDB_ERR_INVALID;
#ifdef SVGDB_WITH_TC32
    if (ctx->precision != SVGDB_PRECISION_TC32) return fail(ctx, SVGDB_ERR_INVALID, "svgdb_time_kernel needs SVGDB_PRECISION_TC32");
    CU(cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    int rc = SVGDB_OK;
    if (which == 0) { // distance pass, collecting a bracket of the given relative half-width around the last median
        TRY(finish_median(ctx));
        if (ctx->n_hist < 1) return fail(ctx, SVGDB_ERR_UNSET, "svgdb_time_kernel: no median yet (run a step first)");
        const double m = ctx->med_hist[0];
        const uint64_t klo = key_of(std::max(m * (1.0 - rel_halfwidth), 0.0)), khi = key_of(m * (1.0 + rel_halfwidth)) + 1;
        rc = launch_dist_operands(ctx);
        ctx->dist_dbg_mode = variant;
        ctx->dist_fold_next = true;
        if (rc == SVGDB_OK) rc = launch_dist_pass_tc32(ctx, MODE_COLLECT, klo, khi, 0); // warm-up
        cudaEventRecord(e0, ctx->stream);
        for (int r = 0; rc == SVGDB_OK && r < reps; ++r) {
            ctx->dist_fold_next = true;
            rc = launch_dist_pass_tc32(ctx, MODE_COLLECT, klo, khi, 0);
        }
        cudaEventRecord(e1, ctx->stream);
        ctx->dist_dbg_mode = 0;
    } else if (which == 1) { // pair-interaction kernel (with its operand preparation and the optimizer epilogue kernel), no state change
        rc = launch_phi_tc32(ctx, true);
        cudaEventRecord(e0, ctx->stream);
        for (int r = 0; rc == SVGDB_OK && r < reps; ++r) rc = launch_phi_tc32(ctx, true);
        cudaEventRecord(e1, ctx->stream);
    } else {
        rc = fail(ctx, SVGDB_ERR_INVALID, "svgdb_time_kernel: unknown kernel id");
    }
    if (rc == SVGDB_OK && cudaEventSynchronize(e1) != cudaSuccess) rc = fail(ctx, SVGDB_ERR_CUDA, "svgdb_time_kernel: synchronize failed");
    float ms = 0.f;
    if (rc == SVGDB_OK) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc == SVGDB_OK) rc = check_tc_err(ctx);
    *ms_out = ms / (float)reps;
    return rc;
#else
    (void)which; (void)variant; (void)rel_halfwidth;
    return fail(ctx, SVGDB_ERR_INVALID, "this build has no TC32 (tcgen05) path");
#endif
}

int svgdb_probe_peak(int device, int what, double *out)
{
    if (!out || what != 0) return SVGDB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return SVGDB_ERR_CUDA;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SVGDB_ERR_CUDA;
    double *sink = nullptr;
    if (cudaMalloc(&sink, sizeof(double)) != cudaSuccess) return SVGDB_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = prop.multiProcessorCount * 4, threads = 256;
    dmma_probe_kernel<<<blocks, threads>>>(iters, sink); // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dmma_probe_kernel<<<blocks, threads>>>(iters, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return SVGDB_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // per warp and iteration: 16 independent DMMAs of 8*8*4 MACs
        double flops = 2.0 * 256.0 * 16.0 * iters * (double)blocks * (threads / 32);
        best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *out = best;
    return SVGDB_OK;
}

} // extern "C"
