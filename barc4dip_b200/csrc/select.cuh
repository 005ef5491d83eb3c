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
    unsigned* cand;                   // (T, regions * FM_REGION + FM_SAMPLE_CAP) keys: CTA regions, then the sample rows'
    int regions;                      // CTA regions per frame (row blocks of rows_inv_kernel)
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

// Census of register-resident magnitudes against a frame's bracket, one value at a time and without a vote, a branch or
// an atomic: every thread owns a column of the CTA's shared-memory stage (slot k of thread `tid` at word k * nthreads +
// tid; the caller provides as many slots as a thread has values, so nothing can overflow) and appends d = key - L of the
// values inside the bracket with a predicated store and a predicated pointer bump. A value outside the bracket costs five
// issue slots: key - L, its sign bit into `below`, one unsigned compare covering both ends, the two predicated-off
// instructions. The CTA then compacts the columns into its own region of the frame's candidate store
// (census_stage_flush). A region holds FM_REGION keys (a CTA of rows_inv_kernel owns 16384 values, ~560 of them inside a
// +-6 sigma bracket of 32 K samples); a count above that marks the frame for the map-based path.
constexpr int FM_REGION = 1536;        // keys per CTA region of the candidate store
constexpr int FM_SAMPLE_CAP = 4096;    // keys the sample rows may contribute

template <int STRIDE_BYTES>
__device__ __forceinline__ void census_stage_value(float v, unsigned Lkey, unsigned W, unsigned& below, unsigned& slot_addr) {
    // The values are magnitudes (non-negative or NaN), so their bit patterns are the order-preserving keys and one
    // subtraction d = key - L gives both tests: below the bracket <=> bit 31 of d (keys and L are < 2^31), inside <=>
    // d <= U - L (unsigned).
    const unsigned d = __float_as_uint(v) - Lkey;
    below += d >> 31;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ls.u32 p, %1, %2;\n\t"
        "@p st.shared.u32 [%0], %1;\n\t"
        "@p add.u32 %0, %0, %3;\n\t"
        "}" : "+r"(slot_addr) : "r"(d), "r"(W), "n"(STRIDE_BYTES) : "memory");
}

// Compaction of the CTA's stage columns (cnt entries of this thread, d = key - L each) into its region of the frame's
// candidate store, in thread order, and into the frame's bracket histogram; cnt3 = (candidates, values below the bracket,
// valid values) of the region. Whole CTA of NT threads (a multiple of 32, at most 1024); s_w: NT / 32 shared words;
// below / nvalid: the CTA's totals, read by thread 0 only. Contains one barrier.
template <int NT>
__device__ __forceinline__ void census_stage_flush(const unsigned* stage, unsigned cnt, unsigned* s_w, unsigned below, unsigned nvalid,
                                                   unsigned Lkey, int shift, unsigned* __restrict__ region,
                                                   unsigned* __restrict__ cnt3, unsigned* __restrict__ hist_q) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned off = incl - cnt, n = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const unsigned v = s_w[w];
        if (w < warp) off += v;
        n += v;
    }
    for (unsigned k = 0; k < cnt; ++k) {
        const unsigned d = stage[k * NT + tid];
        if (off + k < (unsigned)FM_REGION) region[off + k] = d + Lkey;
        atomicAdd(hist_q + min(d >> shift, (unsigned)(SEL_BINS - 1)), 1u);
    }
    if (tid == 0) { cnt3[0] = n; cnt3[1] = below; cnt3[2] = nvalid; }
}
