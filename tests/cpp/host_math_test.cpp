// CPU unit test of svgdcpp_b200/csrc/host_math.hpp (the library's pure host arithmetic): run by tests/test_cabi_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>

#include "host_math.hpp"

using namespace svgdb::host;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

int main()
{
    std::mt19937_64 rng(5);
    std::normal_distribution<double> normal(0.0, 1.0);

    // keys: order preserving on non-negative doubles; key_to_float_ceil = smallest float >= the double
    {
        double prev = 0.0;
        for (int i = 0; i < 2000; ++i) {
            const double v = prev + std::ldexp(std::fabs(normal(rng)), (i % 80) - 40);
            CHECK(key_of(v) >= key_of(prev));
            const float f = key_to_float_ceil(key_of(v));
            CHECK((double)f >= v);
            CHECK(f == 0.0f || (double)std::nextafterf(f, -std::numeric_limits<float>::infinity()) < v);
            prev = v;
        }
        CHECK(key_of(0.0) == 0ull);
        CHECK(std::isinf(key_to_float_ceil(key_of(std::numeric_limits<double>::infinity()))));
        CHECK(std::isinf(key_to_float_ceil(key_of(1e300)))); // beyond the float range
        CHECK(key_to_float_ceil(key_of(1.0)) == 1.0f);
    }

    // Cholesky: A = R^T R with R upper triangular, R Rinv = I; indefinite and non-finite matrices are refused
    for (int d : {1, 2, 7, 64}) {
        std::vector<double> M((size_t)d * d), A((size_t)d * d, 0.0), R, Rinv;
        for (auto &m : M) m = normal(rng);
        for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c) {
                double s = r == c ? 0.5 : 0.0;
                for (int k = 0; k < d; ++k) s += M[(size_t)r * d + k] * M[(size_t)c * d + k] / d;
                A[(size_t)r * d + c] = s;
            }
        CHECK(cholesky_upper(A, d, R, Rinv));
        double err_a = 0.0, err_i = 0.0, below = 0.0;
        for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c) {
                double s = 0.0, t = 0.0;
                for (int k = 0; k < d; ++k) { s += R[(size_t)k * d + r] * R[(size_t)k * d + c]; t += R[(size_t)r * d + k] * Rinv[(size_t)k * d + c]; }
                err_a = std::fmax(err_a, std::fabs(s - A[(size_t)r * d + c]));
                err_i = std::fmax(err_i, std::fabs(t - (r == c ? 1.0 : 0.0)));
                if (r > c) below = std::fmax(below, std::fabs(R[(size_t)r * d + c]) + std::fabs(Rinv[(size_t)r * d + c]));
            }
        CHECK(err_a < 1e-13 * d);
        CHECK(err_i < 1e-12 * d);
        CHECK(below == 0.0);
        std::vector<double> B = A;
        B[(size_t)(d - 1) * d + (d - 1)] = -1.0; // not positive definite
        CHECK(!cholesky_upper(B, d, R, Rinv));
        B = A;
        B[0] = std::numeric_limits<double>::quiet_NaN();
        CHECK(!cholesky_upper(B, d, R, Rinv));
    }

    // row chunks of svgdb_step_host: every i-pair exactly once, in order, nothing negative
    for (int n = 32; n < 5000; n += (n < 300 ? 1 : 37)) {
        int down[4], up[4];
        download_chunks(n, down);
        upload_chunk_ends(n, up);
        CHECK(down[0] >= 0 && down[1] >= 0 && down[2] >= 0 && down[3] >= 0);
        CHECK(down[0] + down[1] + down[2] + down[3] == n);
        CHECK(down[2] + down[3] <= n / 4 + 8); // the exposed tail stays small
        CHECK(up[0] > 0 && up[0] <= up[1] && up[1] <= up[2] && up[2] <= up[3] && up[3] == n);
    }

    // median extrapolation: on a smooth (cubic) sequence the cubic is chosen and exact; on a sequence whose increments alternate
    // around a linear trend (AdaGrad's early overshoot, measured on the config-4 recipe) a parity-aware predictor is chosen and
    // beats the cubic by an order of magnitude
    {
        double m[MEDIAN_HISTORY];
        auto cubic = [](double t) { return 100.0 + 0.3 * t - 0.02 * t * t + 0.001 * t * t * t; };
        for (int k = 0; k < MEDIAN_HISTORY; ++k) m[k] = cubic(10.0 - k); // m[0] most recent (t = 10)
        int kind = -1;
        double err = 1.0;
        double p = median_predict_best(m, MEDIAN_HISTORY, &kind, &err);
        CHECK(kind == 3);
        CHECK(err < 1e-12);
        CHECK(std::fabs(p - cubic(11.0)) < 1e-9);
        // increments d_t = -(0.0025 - 0.00002 t) * (1 + 0.25 (-1)^t): period-two oscillation around a slowly decaying trend
        double seq[16];
        seq[0] = 6000.0;
        for (int t = 1; t < 16; ++t) seq[t] = seq[t - 1] * (1.0 - (0.0025 - 0.00002 * t) * (1.0 + 0.25 * ((t & 1) ? -1.0 : 1.0)));
        for (int k = 0; k < MEDIAN_HISTORY; ++k) m[k] = seq[14 - k]; // history up to t = 14, predict t = 15
        p = median_predict_best(m, MEDIAN_HISTORY, &kind, &err);
        CHECK(kind == 4 || kind == 5);
        const double e_best = std::fabs(p - seq[15]) / seq[15], e_cubic = std::fabs(median_predict(m, 3) - seq[15]) / seq[15];
        CHECK(e_best < 1e-4);
        CHECK(e_best * 10.0 < e_cubic);
        // short histories fall back to what is available
        p = median_predict_best(m, 1, &kind, &err);
        CHECK(kind == 0 && p == m[0] && !std::isfinite(err));
        p = median_predict_best(m, 2, &kind, &err);
        CHECK(kind == 0 || kind == 1);
    }

    if (failures == 0) std::printf("host_math_test: all checks passed\n");
    { // selection rules (thresholds of DESIGN.md sections 3 and 5)
        // AUTO: FAST only for one Gaussian, >= 16,384 particles, d >= 8; explicit variants win
        CHECK(rule_tc32_precise(0, true, 65536, 64) == false);
        CHECK(rule_tc32_precise(0, true, 16384, 8) == false);
        CHECK(rule_tc32_precise(0, true, 16383, 64) == true);
        CHECK(rule_tc32_precise(0, true, 65536, 7) == true);
        CHECK(rule_tc32_precise(0, false, 262144, 64) == true);
        CHECK(rule_tc32_precise(1, false, 100, 2) == false);
        CHECK(rule_tc32_precise(2, true, 65536, 64) == true);
        // lean pair kernel: FAST, d <= 64 kernels, d >= 48, N >= 16,384; environment override; never with PRECISE or the wide kernels
        CHECK(rule_phi_lean(false, false, -1, 65536, 64) == true);
        CHECK(rule_phi_lean(false, false, -1, 16384, 48) == true);
        CHECK(rule_phi_lean(false, false, -1, 65536, 47) == false);
        CHECK(rule_phi_lean(false, false, -1, 16383, 64) == false);
        CHECK(rule_phi_lean(false, false, 1, 128, 2) == true);
        CHECK(rule_phi_lean(false, false, 0, 65536, 64) == false);
        CHECK(rule_phi_lean(true, false, 1, 65536, 64) == false);
        CHECK(rule_phi_lean(false, true, 1, 65536, 128) == false);
        CHECK(rule_phi_one_term_v(-1, 32768) == true);
        CHECK(rule_phi_one_term_v(-1, 32767) == false);
        CHECK(rule_phi_one_term_v(0, 1 << 20) == false);
        CHECK(rule_phi_one_term_v(1, 128) == true);
        // gradient through library DGEMMs: Gaussian-sum models only; d >= 32 and rows x components >= 16,384
        CHECK(rule_grad_gemm(true, -1, 1, 64, 65536) == true);
        CHECK(rule_grad_gemm(true, -1, 16, 256, 1024) == true);
        CHECK(rule_grad_gemm(true, -1, 16, 256, 1023) == false);
        CHECK(rule_grad_gemm(true, -1, 1, 31, 1 << 20) == false);
        CHECK(rule_grad_gemm(true, -1, 1, 64, 8192) == false);   // config 3 on 8 GPUs: 8192 rows per rank stay with the one-kernel form
        CHECK(rule_grad_gemm(false, 1, 1, 64, 65536) == false);  // gradient hooks never
        CHECK(rule_grad_gemm(true, 0, 16, 256, 262144) == false);
        CHECK(rule_grad_gemm(true, 1, 2, 5, 10) == true);
    }
    return failures == 0 ? 0 : 1;
}
