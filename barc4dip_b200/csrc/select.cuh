// Shared pieces of the exact order-statistics machinery (select.cu) that other kernels fuse into their epilogues:
// the per-frame bracket state, the order-preserving float keys, and the census of register-resident values against
// a bracket (used by rows_inv_kernel to take the median of |corr| without materialising the map).
#pragma once

#include "common.cuh"

constexpr int SEL_MAXR = 4;           // ranks per frame (2 per quantile)
constexpr int SEL_MAXQ = 2;
constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;
constexpr int SEL_SAMPLES = 16384;
constexpr int64_t SEL_FAST_MIN = 65536;

struct SelState {                     // per frame, device resident (radix path)
    unsigned prefix[SEL_MAXR];        // key bits fixed so far (left aligned)
    long long rank[SEL_MAXR];         // residual rank inside the prefix bucket
    long long n_valid;
};

struct SelFast {                      // per frame, device resident (bracket path)
    unsigned L[SEL_MAXQ], U[SEL_MAXQ];
    unsigned long long below[SEL_MAXQ];
    unsigned ncand[SEL_MAXQ];
    unsigned long long n_valid;
    int need_fallback;
};

// order-preserving key; -0.0 maps onto +0.0 so that key order and float order agree for every non-NaN value
__device__ __forceinline__ unsigned key_of(float v, int use_abs) {
    unsigned b = __float_as_uint(v);
    if (use_abs) return b & 0x7fffffffu;
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float value_of(unsigned k, int use_abs) {
    if (use_abs) return __uint_as_float(k);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct FusedMedian {                  // device scratch of one batch of the fused median (b4d_fused_median_*)
    double* q_dev;                    // the quantile (0.5)
    SelFast* st;                      // per frame: bracket [L, U]; ncand / below / n_valid of the SAMPLE rows
    int* need;
    unsigned* bhist;                  // (T, SEL_BINS) bracket histogram
    unsigned* cnt3;                   // (T, regions, 3): candidates, values below the bracket, valid values of a region
    unsigned* cand;                   // (T, regions * FM_REGION + FM_SAMPLE_CAP) keys: warp regions, then the sample rows'
    int regions;                      // warp regions per frame
};

// numpy's linear method: h = n*q + (1 + q*(1-1-1)) - 1, lo = floor(h), hi = min(lo+1, n-1)
__device__ __forceinline__ void target_ranks(unsigned long long n, double q, long long& lo, long long& hi) {
    lo = hi = 0;
    if (n == 0) return;
    const double hh = __dadd_rn(__dadd_rn(__dmul_rn((double)n, q), __dadd_rn(1.0, __dmul_rn(q, -1.0))), -1.0);
    lo = (long long)floor(hh);
    if (lo < 0) lo = 0;
    if (lo > (long long)n - 1) lo = (long long)n - 1;
    hi = lo + 1 > (long long)n - 1 ? (long long)n - 1 : lo + 1;
}


// bin of a candidate inside its bracket [L, U]: (key - L) >> shift with the smallest shift that fits SEL_BINS bins
__device__ __forceinline__ int bracket_shift(unsigned L, unsigned U) {
    const int width = 32 - __clz((U - L) | 1u);
    return width > 11 ? width - 11 : 0;
}

// Census of NV register-resident values against bracket q of a frame, whole warps only: counts the values below the
// bracket into `below`, appends the keys of the values inside it to the CTA's shared-memory stage (one warp scan and
// one shared atomic per call). NaNs are neither below nor inside; the caller counts valid values itself.
template <int NV>
__device__ __forceinline__ void census_values(const float (&v)[NV], float Lf, float Uf, int use_abs, unsigned& below,
                                              unsigned* s_keys, unsigned* s_n, unsigned stage_cap, int lane) {
    unsigned c = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        below += (v[i] < Lf);
        c += (v[i] >= Lf && v[i] <= Uf);
    }
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    const unsigned wtot = __shfl_sync(0xffffffffu, incl, 31);
    if (wtot) {
        unsigned base = 0;
        if (lane == 31) base = atomicAdd(s_n, wtot);
        base = __shfl_sync(0xffffffffu, base, 31);
        if (c) {
            unsigned pos = base + incl - c;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (v[i] >= Lf && v[i] <= Uf) {
                    if (pos < stage_cap) s_keys[pos] = key_of(v[i], use_abs);
                    ++pos;
                }
            }
        }
    }
}

// Flush of a CTA's stage into the frame's candidate list (one global reservation) and bracket histogram.
// Call with the whole CTA after a barrier that made the stage complete; s_base is a shared scratch word.
__device__ __forceinline__ void census_flush(SelFast* s, int q, const unsigned* s_keys, unsigned s_count, unsigned stage_cap,
                                             unsigned* cand_q, unsigned cap, unsigned* hist_q, unsigned* s_base) {
    if (threadIdx.x == 0) {
        unsigned base = 0;
        if (s_count > stage_cap) s->need_fallback = 1;
        else if (s_count) base = atomicAdd(&s->ncand[q], s_count);
        *s_base = base;
    }
    __syncthreads();
    const unsigned cnt = min(s_count, stage_cap), base = *s_base;
    const unsigned L = s->L[q];
    const int shift = bracket_shift(L, s->U[q]);
    for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
        const unsigned key = s_keys[i];
        if (base + i < cap) cand_q[base + i] = key;
        atomicAdd(hist_q + min((key - L) >> shift, (unsigned)(SEL_BINS - 1)), 1u);
    }
}

// Warp-level census into the warp's OWN region of the frame's candidate store: no reservation, no atomic whose result
// anybody waits for. Lane counts -> one warp scan -> keys written at the scan offsets, (count, below, valid) of the
// region stored by lane 31, bracket histogram by fire-and-forget RED. A region holds FM_REGION keys; a count above that
// marks the frame for the map-based path. Whole warps only.
constexpr int FM_REGION = 64;          // keys per warp and call (a warp holds 512 values, < 5 % of them inside a bracket)
constexpr int FM_SAMPLE_CAP = 4096;    // keys the sample rows may contribute

template <int NV, int COMP>
__device__ __forceinline__ void census_values_region(const float2 (&x)[NV], unsigned Lkey, unsigned Ukey, int shift,
                                                     unsigned* __restrict__ region, unsigned* __restrict__ cnt3,
                                                     unsigned* __restrict__ hist_q, int lane) {
    // The values are magnitudes (non-negative or NaN), so their bit patterns are the order-preserving keys and one
    // subtraction d = key - L gives both tests: below the bracket <=> bit 31 of d (keys and L are < 2^31), inside <=>
    // d <= U - L (unsigned). The keys inside are compacted slot by slot with a warp vote: a slot nobody is inside of
    // costs five instructions.
    const unsigned W = min(Ukey, 0x7fffffffu) - Lkey;
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    unsigned below = 0, base = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const unsigned d = __float_as_uint(COMP ? x[i].y : x[i].x) - Lkey;
        below += d >> 31;
        const bool in = d <= W;
        if (__any_sync(0xffffffffu, in)) {
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) {
                const unsigned pos = base + __popc(m & lt);
                if (pos < FM_REGION) region[pos] = d + Lkey;
                atomicAdd(hist_q + min(d >> shift, (unsigned)(SEL_BINS - 1)), 1u);
            }
            base += __popc(m);
        }
    }
    // A NaN anywhere in the frame reaches every output of its inverse transform, so one value per lane tells whether the
    // warp's values are valid.
    const float v0 = COMP ? x[0].y : x[0].x;
    const unsigned nvalid = (unsigned)__popc(__ballot_sync(0xffffffffu, v0 == v0)) * (unsigned)NV;
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 31) { cnt3[0] = base; cnt3[1] = below; cnt3[2] = nvalid; }
}
