// host_math.hpp -- the pure host-side arithmetic of the library (no CUDA): shared by svgd_b200_api.cu and by the CPU unit test
// tests/cpp/host_math_test.cpp, so that this logic is exercised without a GPU.
#pragma once
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

} // namespace host
} // namespace svgdb
