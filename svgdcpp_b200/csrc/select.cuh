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

constexpr int SELECT_MAX_BITS = 12;              // widest digit (tensor-core keys: 24 significant bits = 2 passes)
constexpr int SELECT_MAX_BINS = 1 << SELECT_MAX_BITS;

struct SelectState {
    unsigned long long base;      // every candidate is selected on key - base (the bracket's lower end): a narrow bracket has few
                                  // significant bits whatever power-of-two boundaries its raw bit patterns straddle
    unsigned long long prefix;    // bits of the answer (as key - base) decided so far
    unsigned long long mask;      // which bits of `prefix` are decided
    unsigned long long rank;      // rank of the answer among the keys matching prefix/mask
    unsigned long long n_less;    // candidates strictly below every key matching prefix/mask
    unsigned long long max_less;  // largest candidate (as key - base) below the answer once all bits are decided
    unsigned long long need_scan; // the predecessor is not in the answer's last digit group: select_max_less_kernel must scan
    unsigned long long hist[SELECT_MAX_BINS];
};

struct MedianResult {
    unsigned long long key_lo, key_hi; // bit patterns of the two middle D2 (equal when n^2 is odd)
    double d2_lo, d2_hi;
    double median;                     // of the distances
    double scale;                      // a
};

// The verdict on a collecting pass over a PREDICTED bracket, taken on the device so that the step needs no host round trip:
// the pass returned exact counts (all-reduced over the ranks), so whether the bracket held the median is arithmetic on them.
struct StepDecision {
    unsigned long long kk;      // rank of the upper middle value among the candidates (k_hi - below)
    unsigned long long m_local; // this rank's candidates
    unsigned long long below, mid; // the pass's global counts (for the host's bookkeeping, read later)
    int hit;                    // the bracket held: the select below works on the right candidates
    int pad;
};
// pass_words = [below, cand_total] (global).  need_both: the pass does not track the largest value below the bracket, so for an
// even count the lower middle value must be a candidate too.  miss_sticky is raised on a miss: everything that would change
// persistent state (the optimizer update) checks it, the host reads it after the step and repeats the step the slow way.
__global__ void median_decide_kernel(const unsigned long long *pass_words, const unsigned long long *cand_count_local, unsigned long long k_hi,
                                     int even, int need_both, unsigned long long capacity, StepDecision *dec, int *miss_sticky)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long b = pass_words[0], mid = pass_words[1];
    const bool hit = (need_both ? b < k_hi : b <= k_hi) && (k_hi < b + mid) && (mid <= capacity) && (!even || k_hi >= 1ull);
    dec->hit = hit ? 1 : 0;
    dec->kk = hit ? k_hi - b : 0ull;
    const unsigned long long ml = *cand_count_local;
    dec->m_local = ml < capacity ? ml : capacity;
    dec->below = b;
    dec->mid = mid;
    if (!hit) *miss_sticky = 1;
}

// dec != nullptr: the rank comes from the device-side verdict instead of `rank`
__global__ void select_init_kernel(SelectState *st, unsigned long long base, unsigned long long prefix, unsigned long long mask,
                                   unsigned long long rank, const StepDecision *dec = nullptr)
{
    for (int t = threadIdx.x; t < SELECT_MAX_BINS; t += blockDim.x) st->hist[t] = 0ull;
    if (threadIdx.x == 0) {
        st->base = base; st->prefix = prefix; st->mask = mask; st->rank = dec ? dec->kk : rank;
        st->n_less = 0ull; st->max_less = 0ull; st->need_scan = 1ull;
    }
}

// histogram of the `bits`-wide digit at `shift` over the candidates that match the decided prefix
__global__ void __launch_bounds__(256)
select_hist_kernel(const unsigned long long *__restrict__ cand, unsigned long long m, int shift, int bits, SelectState *st,
                   const StepDecision *dec = nullptr)
{
    if (dec) m = dec->m_local;
    __shared__ unsigned int sh[SELECT_MAX_BINS];
    const int nbins = 1 << bits;
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    const unsigned long long base = st->base, prefix = st->prefix, mask = st->mask, dmask = (unsigned long long)(nbins - 1);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < m;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long key = cand[t] - base;
        if ((key & mask) == prefix) atomicAdd(&sh[(unsigned int)((key >> shift) & dmask)], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
        if (sh[b]) atomicAdd(&st->hist[b], (unsigned long long)sh[b]);
}

// one block: pick the digit holding `rank`, extend the prefix, reset the histogram.  On the last digit (`last` != 0) the
// predecessor of the answer among the candidates is the largest non-empty smaller digit of the same group, if there is
// one; otherwise need_scan stays set and select_max_less_kernel looks for it among the smaller groups.
__global__ void __launch_bounds__(256) select_pick_kernel(SelectState *st, int shift, int bits, int last)
{
    __shared__ unsigned long long sh[SELECT_MAX_BINS];
    __shared__ unsigned long long part[256];
    const int nbins = 1 << bits, per = nbins / 256 > 0 ? nbins / 256 : 1, t = threadIdx.x;
    for (int b = t; b < nbins; b += 256) sh[b] = st->hist[b];
    __syncthreads();
    unsigned long long mine = 0ull; // thread t owns bins [t * per, t * per + per)
    if (t * per < nbins)
        for (int b = t * per; b < t * per + per; ++b) mine += sh[b];
    part[t] = mine;
    __syncthreads();
    if (t == 0) {
        const unsigned long long rank = st->rank;
        unsigned long long cum = 0ull;
        int seg = 255;
        for (int q = 0; q < 256; ++q) {
            if (rank < cum + part[q]) { seg = q; break; }
            cum += part[q];
        }
        int digit = nbins - 1;
        const int b0 = seg * per < nbins ? seg * per : nbins - per;
        for (int b = b0; b < b0 + per; ++b) {
            if (rank < cum + sh[b]) { digit = b; break; }
            cum += sh[b];
        }
        int pred = -1;
        for (int b = digit - 1; b >= 0; --b)
            if (sh[b] != 0ull) { pred = b; break; }
        st->rank = rank - cum;
        st->n_less += cum;
        const unsigned long long np = st->prefix | (((unsigned long long)digit) << shift);
        st->prefix = np;
        st->mask |= ((unsigned long long)(nbins - 1)) << shift;
        if (last && pred >= 0) {
            st->max_less = (np & ~(((unsigned long long)(nbins - 1)) << shift)) | (((unsigned long long)pred) << shift);
            st->need_scan = 0ull;
        }
    }
    __syncthreads();
    for (int b = t; b < nbins; b += 256) st->hist[b] = 0ull;
}

// largest candidate below the answer, only when the last pick could not tell (the answer is the smallest key of its group)
__global__ void __launch_bounds__(256)
select_max_less_kernel(const unsigned long long *__restrict__ cand, unsigned long long m, SelectState *st, const StepDecision *dec = nullptr)
{
    if (st->need_scan == 0ull) return;
    if (dec) m = dec->m_local;
    const unsigned long long base = st->base, key_hi = st->prefix;
    unsigned long long best = 0ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < m;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long key = cand[t] - base;
        if (key < key_hi && key >= best) best = key; // (key - base may be 0: the atomicMax below still records it)
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
                                       double *a_out, const StepDecision *dec = nullptr)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (dec) kk = dec->kk;
    unsigned long long key_hi = direct ? direct_key : st->base + st->prefix;
    unsigned long long key_lo = key_hi;
    if (even) {
        if (kk == 0ull) key_lo = *max_below_global;                 // predecessor lies below the bracket
        else if (!direct && st->n_less == kk) key_lo = st->base + st->max_less; // predecessor is a smaller candidate
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
