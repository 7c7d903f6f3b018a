// host_math.hpp -- the pure host-side arithmetic of the library (no CUDA): shared by svgd_b200_api.cu and by the CPU unit test
// tests/cpp/host_math_test.cpp, so that this logic is exercised without a GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace svgdb {
namespace host {

// Order-preserving integer key of a non-negative double: its IEEE bit pattern (the median select works on these keys).
inline uint64_t key_of(double v)
{
    uint64_t k;
    std::memcpy(&k, &v, 8);
    return k;
}

// smallest float >= the double whose bits are `key` (+inf for keys past +inf)
inline float key_to_float_ceil(uint64_t key)
{
    if (key >= 0x7FF0000000000000ull) return INFINITY;
    double x;
    std::memcpy(&x, &key, 8);
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}

// ---- prediction of the next median of D2 from the recent ones (bracket of the one-pass median, svgd_b200_api.cu: median_scale) ----
// m[0] is the most recent value, n the number of valid entries.  Six extrapolations; which one fits depends on the optimizer:
// Adam's trajectory is smooth (the cubic wins), AdaGrad's first dozens of steps overshoot with period two (the increments
// alternate around their trend: the predictors that continue the increment of TWO steps ago win).
constexpr int MEDIAN_HISTORY = 8;
constexpr int MEDIAN_PREDICTORS = 6;
inline int median_predictor_needs(int kind)
{
    static const int needs[MEDIAN_PREDICTORS] = {1, 2, 3, 4, 3, 5};
    return needs[kind];
}
inline double median_predict(const double *m, int kind)
{
    switch (kind) {
    case 0: return m[0];                                            // constant
    case 1: return 2.0 * m[0] - m[1];                               // linear
    case 2: return 3.0 * m[0] - 3.0 * m[1] + m[2];                  // quadratic
    case 3: return 4.0 * m[0] - 6.0 * m[1] + 4.0 * m[2] - m[3];     // cubic
    case 4: return m[0] + (m[1] - m[2]);                            // the increment of two steps ago
    default: return m[0] + 2.0 * (m[1] - m[2]) - (m[3] - m[4]);     // ... extrapolated linearly within its parity class
    }
}
// Picks the extrapolation that would have predicted the last (up to) two medians best; returns the prediction of the next one and,
// in *err, that predictor's worst relative back-test error (INFINITY if the history is too short for any back-test).
inline double median_predict_best(const double *m, int n, int *kind_out, double *err)
{
    int best = -1;
    double best_err = INFINITY;
    for (int k = 0; k < MEDIAN_PREDICTORS; ++k) {
        const int need = median_predictor_needs(k);
        double e = -1.0;
        for (int s = 1; s <= 2; ++s) { // predict m[s - 1] from m[s ...]
            if (n < need + s) break;
            const double target = m[s - 1];
            const double rel = std::fabs(median_predict(m + s, k) - target) / std::fabs(target);
            e = std::max(e, rel);
        }
        if (e >= 0.0 && e < best_err) { best_err = e; best = k; }
    }
    if (best < 0) { // no back-test possible yet: the highest order the history allows, as before
        best = n >= 4 ? 3 : n == 3 ? 2 : n == 2 ? 1 : 0;
    }
    if (kind_out) *kind_out = best;
    if (err) *err = best_err;
    return median_predict(m, best);
}

// Cholesky A = R^T R with R upper triangular; returns false if A is not positive definite.  Rinv = R^-1.
inline bool cholesky_upper(const std::vector<double> &A, int d, std::vector<double> &R, std::vector<double> &Rinv)
{
    R.assign((size_t)d * d, 0.0);
    for (int j = 0; j < d; ++j) {
        double s = A[(size_t)j * d + j];
        for (int k = 0; k < j; ++k) s -= R[(size_t)k * d + j] * R[(size_t)k * d + j];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        const double rjj = std::sqrt(s);
        R[(size_t)j * d + j] = rjj;
        for (int c = j + 1; c < d; ++c) {
            double t = A[(size_t)j * d + c];
            for (int k = 0; k < j; ++k) t -= R[(size_t)k * d + j] * R[(size_t)k * d + c];
            R[(size_t)j * d + c] = t / rjj;
        }
    }
    Rinv.assign((size_t)d * d, 0.0); // back substitution, column by column: R Rinv = I
    for (int c = 0; c < d; ++c)
        for (int r = c; r >= 0; --r) {
            double t = (r == c) ? 1.0 : 0.0;
            for (int k = r + 1; k <= c; ++k) t -= R[(size_t)r * d + k] * Rinv[(size_t)k * d + c];
            Rinv[(size_t)r * d + c] = t / R[(size_t)r * d + r];
        }
    return true;
}

// svgdb_step_host, download: the rows leave in four chunks of 3/8, 3/8, 1/8, 1/8 of the i-pairs (256 rows each): every copy but the
// last hides behind the pair interactions of the chunks after it, and the exposed tail is an eighth of the transfer.
inline void download_chunks(int n_ipairs, int chunk_ipairs[4])
{
    chunk_ipairs[0] = chunk_ipairs[1] = (3 * n_ipairs) / 8;
    chunk_ipairs[2] = n_ipairs / 8;
    chunk_ipairs[3] = n_ipairs - chunk_ipairs[0] - chunk_ipairs[1] - chunk_ipairs[2];
}

// svgdb_step_host, upload: four equal chunks; entry k is the i-pair at which chunk k ends.
inline void upload_chunk_ends(int n_ipairs, int ends[4])
{
    for (int ch = 0; ch < 4; ++ch) ends[ch] = ch == 3 ? n_ipairs : (n_ipairs * (ch + 1)) / 4;
}

// ---- which arithmetic / which kernels a context runs: the measured thresholds in one place (DESIGN.md sections 3 and 5) ------------
// `forced` arguments: -1 = automatic, 0 = off, 1 = on (environment knobs); `variant`: 0 AUTO, 1 FAST, 2 PRECISE (svgdb_tc32_variant).

// PRECISE unless FAST's zero-mean error terms average out by construction: one Gaussian target, >= 16,384 particles, d >= 8.
inline bool rule_tc32_precise(int variant, bool one_gaussian_target, int64_t n, int d)
{
    if (variant == 1) return false;
    if (variant == 2) return true;
    return !(one_gaussian_target && n >= 16384 && d >= 8);
}

// lean FAST pair kernel (e5m2 `lo_i . y^_j`, row sums on the tensor core): d <= 64 kernels only; pays and keeps the error from d >= 48.
inline bool rule_phi_lean(bool precise, bool wide, int forced, int64_t n, int d)
{
    if (precise || wide) return false;
    if (forced >= 0) return forced != 0;
    return d >= 48 && n >= 16384;
}

// ... and v = grad - 2 a x~ in one fp16 term (E . v_lo left out) once a row has enough neighbours to average the rounding of v_j.
inline bool rule_phi_one_term_v(int forced, int64_t n) { return forced >= 0 ? forced != 0 : n >= 32768; }

// grad log p through library DGEMMs + streaming kernels instead of the one-kernel form: where a GEMM is not all launch latency.
inline bool rule_grad_gemm(bool gaussian_sum_model, int forced, int components, int d, int64_t rows)
{
    if (!gaussian_sum_model) return false;
    if (forced >= 0) return forced != 0 && components >= 1;
    return d >= 32 && rows * (int64_t)components >= 16384;
}

} // namespace host
} // namespace svgdb
