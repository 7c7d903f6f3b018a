// select.cuh — exact on-device radix select over the candidate keys gathered by the distance
// pass, and the final bandwidth a = log(n) / med^2.
//
// Reference semantics (Kernel/GaussianRBFKernel.hpp:222-254): for an even number of values the
// median is the mean of the two middle order statistics, for an odd number the middle one.  The
// keys are IEEE bit patterns of D2 >= 0; sqrt is monotone and correctly rounded, so selecting on
// D2 and taking the square root of the two selected values equals the reference's
// sqrt-then-nth_element (:185-187).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svgdb {

struct SelectState {
    unsigned long long prefix;    // bits of the answer decided so far
    unsigned long long mask;      // which bits of `prefix` are decided
    unsigned long long rank;      // rank of the answer among the keys matching prefix/mask
    unsigned long long n_less;    // candidates strictly below every key matching prefix/mask
    unsigned long long max_less;  // largest candidate < prefix once all bits are decided
    unsigned long long hist[256];
};

struct MedianResult {
    unsigned long long key_lo, key_hi; // bit patterns of the two middle D2 (equal when n^2 is odd)
    double d2_lo, d2_hi;
    double median;                     // of the distances
    double scale;                      // a
};

__global__ void select_init_kernel(SelectState *st, unsigned long long prefix, unsigned long long mask,
                                   unsigned long long rank)
{
    int t = threadIdx.x;
    if (t < 256) st->hist[t] = 0ull;
    if (t == 0) { st->prefix = prefix; st->mask = mask; st->rank = rank; st->n_less = 0ull; st->max_less = 0ull; }
}

__global__ void __launch_bounds__(256)
select_hist_kernel(const unsigned long long *__restrict__ cand, unsigned long long m, int shift, SelectState *st)
{
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < m;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long key = cand[t];
        if ((key & mask) == prefix) atomicAdd(&sh[(unsigned int)((key >> shift) & 255ull)], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// one thread: pick the digit holding `rank`, extend the prefix, reset the histogram
__global__ void select_pick_kernel(SelectState *st, int shift)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long rank = st->rank, cum = 0ull;
    int digit = 255;
    for (int b = 0; b < 256; ++b) {
        unsigned long long c = st->hist[b];
        if (rank < cum + c) { digit = b; break; }
        cum += c;
    }
    st->rank = rank - cum;
    st->n_less += cum;
    st->prefix |= ((unsigned long long)digit) << shift;
    st->mask |= 255ull << shift;
    for (int b = 0; b < 256; ++b) st->hist[b] = 0ull;
}

__global__ void __launch_bounds__(256)
select_max_less_kernel(const unsigned long long *__restrict__ cand, unsigned long long m, SelectState *st)
{
    const unsigned long long key_hi = st->prefix;
    unsigned long long best = 0ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < m;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long key = cand[t];
        if (key < key_hi && key > best) best = key;
    }
    for (int o = 16; o; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&st->max_less, best);
}

// kk = rank of the upper middle value among the candidates (k_hi - below); even != 0 when n^2 is even.
// direct != 0: the answer key was fixed by the caller (massive tie), st is not consulted.
__global__ void median_finalize_kernel(const SelectState *st, unsigned long long kk, int even,
                                       const unsigned long long *max_below_global, int direct,
                                       unsigned long long direct_key, double log_n, MedianResult *out,
                                       double *a_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long key_hi = direct ? direct_key : st->prefix;
    unsigned long long key_lo = key_hi;
    if (even) {
        if (kk == 0ull) key_lo = *max_below_global;                 // predecessor lies below the bracket
        else if (!direct && st->n_less == kk) key_lo = st->max_less; // predecessor is a smaller candidate
        // otherwise the predecessor ties with key_hi
    }
    double d_lo = __longlong_as_double((long long)key_lo), d_hi = __longlong_as_double((long long)key_hi);
    double med = (sqrt(d_lo) + sqrt(d_hi)) / 2.0;
    out->key_lo = key_lo;
    out->key_hi = key_hi;
    out->d2_lo = d_lo;
    out->d2_hi = d_hi;
    out->median = med;
    out->scale = log_n / (med * med);
    *a_out = out->scale;
}

} // namespace svgdb
