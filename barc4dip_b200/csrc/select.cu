// Exact order statistics per frame by radix select on the float bit patterns.
//
// Behind np.nanpercentile(img, 0.05 / 99.95) in amplitude() (metrics/speckles.py:647 via
// utils/range.py:51-54), np.median(|corr|) in the tracker's SNR (signal/tracking.py:319) and the
// two medians of flat_field_correction (preprocessing/normalize.py:110,126).
//
// Three histogram passes over the frame (11 + 11 + 10 key bits). After each pass a one-CTA-per-
// frame kernel walks the cumulative histogram and narrows every requested rank to a key prefix.
// Keys are order-preserving uint32 images of the floats; NaNs are skipped (nanpercentile
// semantics) and counted out of n_valid. Hot bins (speckle intensities share a few exponents)
// are handled with warp-aggregated shared-memory atomics.
#include "common.cuh"

namespace {

constexpr int SEL_MAXR = 4;           // ranks per frame (2 per quantile)
constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;

struct SelState {                     // per frame, device resident
    unsigned prefix[SEL_MAXR];        // key bits fixed so far (left aligned)
    long long rank[SEL_MAXR];         // residual rank inside the prefix bucket
    long long n_valid;
};

__device__ __forceinline__ unsigned key_of(float v, int use_abs) {
    unsigned b = __float_as_uint(v);
    if (use_abs) return b & 0x7fffffffu;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float value_of(unsigned k, int use_abs) {
    if (use_abs) return __uint_as_float(k);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ void hist_add(unsigned* h, unsigned bin) {
    const unsigned m = __match_any_sync(__activemask(), bin);
    if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) atomicAdd(h + bin, (unsigned)__popc(m));
}

// PASS 0: bits 31..21, PASS 1: bits 20..10 (given 11-bit prefix), PASS 2: bits 9..0 (given 22-bit prefix)
template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) sel_hist_kernel(const float* __restrict__ stack, int64_t n,
                                                               int nr, int use_abs,
                                                               const SelState* __restrict__ st,
                                                               unsigned* __restrict__ hist) {
    extern __shared__ unsigned sh[];                      // nr_eff * SEL_BINS
    const int64_t t = blockIdx.y;
    const int nre = (PASS == 0) ? 1 : nr;
    for (int i = threadIdx.x; i < nre * SEL_BINS; i += blockDim.x) sh[i] = 0;
    unsigned pre[SEL_MAXR];
#pragma unroll
    for (int r = 0; r < SEL_MAXR; ++r) pre[r] = (PASS > 0 && r < nr) ? st[t].prefix[r] : 0u;
    __syncthreads();

    const float* f = stack + t * n;
    constexpr int SHIFT = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    constexpr int PSHIFT = (PASS == 1) ? 21 : 10;         // bits below the known prefix
    constexpr unsigned MASK = (PASS == 2) ? 1023u : 2047u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(f + i);
        if (v != v) continue;
        const unsigned k = key_of(v, use_abs);
        if (PASS == 0) {
            hist_add(sh, k >> SHIFT);
        } else {
#pragma unroll
            for (int r = 0; r < SEL_MAXR; ++r) {
                if (r < nr && (k >> PSHIFT) == (pre[r] >> PSHIFT)) {
                    // duplicate prefixes are counted once, into the first rank that owns them
                    bool first = true;
#pragma unroll
                    for (int q = 0; q < r; ++q) first = first && (pre[q] >> PSHIFT) != (pre[r] >> PSHIFT);
                    if (first) hist_add(sh + r * SEL_BINS, (k >> SHIFT) & MASK);
                }
            }
        }
    }
    __syncthreads();
    unsigned* g = hist + (size_t)t * SEL_MAXR * SEL_BINS;
    for (int i = threadIdx.x; i < nre * SEL_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(g + i, sh[i]);
}

// one CTA per frame: locate each rank's bin in the (shared) histogram of its prefix
template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) sel_scan_kernel(unsigned* __restrict__ hist, int nr, int n_q,
                                                               const double* __restrict__ quant, int use_abs,
                                                               SelState* __restrict__ st, float* __restrict__ out,
                                                               long long* __restrict__ n_valid_out) {
    const int64_t t = blockIdx.x;
    unsigned* g = hist + (size_t)t * SEL_MAXR * SEL_BINS;
    __shared__ unsigned long long cum[SEL_BINS];
    __shared__ SelState s;
    constexpr int PSHIFT = (PASS == 1) ? 21 : 10;
    constexpr int SHIFT = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
    if (threadIdx.x == 0) s = st[t];
    __syncthreads();

    for (int r = 0; r < nr; ++r) {
        // histogram owner: first rank with the same prefix (pass 0: rank 0)
        int owner = r;
        if (PASS == 0) owner = 0;
        else for (int q = r - 1; q >= 0; --q) if ((s.prefix[q] >> PSHIFT) == (s.prefix[r] >> PSHIFT)) owner = q;
        const unsigned* h = g + owner * SEL_BINS;
        // inclusive scan of 2048 bins by 256 threads (8 bins each) -- sizes are tiny, keep it simple
        unsigned long long loc[8], run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { run += h[threadIdx.x * 8 + i]; loc[i] = run; }
        __shared__ unsigned long long part[SEL_THREADS];
        part[threadIdx.x] = run;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long a = 0;
            for (int i = 0; i < SEL_THREADS; ++i) { unsigned long long v = part[i]; part[i] = a; a += v; }
            if (PASS == 0 && r == 0) {
                s.n_valid = (long long)a;
                // numpy's linear method: h = n*q + (1 + q*(1-1-1)) - 1, lo = floor(h), hi = min(lo+1, n-1)
                for (int q = 0; q < n_q; ++q) {
                    long long lo = 0, hi = 0;
                    if (a > 0) {
                        const double qq = quant[q];
                        const double hh = __dadd_rn(__dadd_rn(__dmul_rn((double)a, qq), __dadd_rn(1.0, __dmul_rn(qq, -1.0))), -1.0);
                        lo = (long long)floor(hh);
                        if (lo < 0) lo = 0;
                        if (lo > (long long)a - 1) lo = (long long)a - 1;
                        hi = lo + 1 > (long long)a - 1 ? (long long)a - 1 : lo + 1;
                    }
                    s.rank[2 * q] = lo; s.rank[2 * q + 1] = hi;
                    s.prefix[2 * q] = 0; s.prefix[2 * q + 1] = 0;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) cum[threadIdx.x * 8 + i] = loc[i] + part[threadIdx.x];
        __syncthreads();
        const long long rk = s.rank[r];
        // the bin b with cum[b-1] <= rk < cum[b]
        for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) {
            const unsigned long long lo = b ? cum[b - 1] : 0ull, hi = cum[b];
            if ((unsigned long long)rk >= lo && (unsigned long long)rk < hi) {
                s.prefix[r] |= ((unsigned)b) << SHIFT;
                s.rank[r] = rk - (long long)lo;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st[t] = s;
        if (PASS == 2) {
            for (int r = 0; r < nr; ++r)
                out[t * nr + r] = s.n_valid > 0 ? value_of(s.prefix[r], use_abs) : __uint_as_float(0x7fc00000u);
            if (n_valid_out) n_valid_out[t] = s.n_valid;
        }
    }
    __syncthreads();
    // clear for the next pass
    for (int i = threadIdx.x; i < SEL_MAXR * SEL_BINS; i += blockDim.x) g[i] = 0;
}

}  // namespace

int b4d_select_impl(b4d_ctx* ctx, const float* stack, int64_t T, int64_t n, const double* q_dev, int n_q,
                    int use_abs, float* out, int64_t* n_valid) {
    const int nr = 2 * n_q;
    void* p = nullptr;
    const size_t hist_bytes = (size_t)T * SEL_MAXR * SEL_BINS * sizeof(unsigned);
    const size_t st_bytes = (size_t)T * sizeof(SelState);
    int rc = b4d_scratch(ctx, SCR_SELECT, hist_bytes + st_bytes, &p);
    if (rc) return rc;
    unsigned* hist = static_cast<unsigned*>(p);
    SelState* st = reinterpret_cast<SelState*>(static_cast<char*>(p) + hist_bytes);
    B4D_CUDA(ctx, cudaMemsetAsync(p, 0, hist_bytes + st_bytes, ctx->stream));

    int bpf = (int)((n + SEL_THREADS * 16 - 1) / (SEL_THREADS * 16));
    const int cap = ctx->sm_count * 8;
    if ((int64_t)bpf * T > cap) bpf = (int)((cap + T - 1) / T);
    if (bpf < 1) bpf = 1;
    for (int64_t t0 = 0; t0 < T; t0 += 32768) {
        const unsigned tc = (unsigned)((T - t0 < 32768) ? T - t0 : 32768);
        dim3 grid((unsigned)bpf, tc);
        const float* s0 = stack + t0 * n;
        unsigned* h0 = hist + (size_t)t0 * SEL_MAXR * SEL_BINS;
        SelState* st0 = st + t0;
        float* o0 = out + t0 * nr;
        long long* nv0 = n_valid ? reinterpret_cast<long long*>(n_valid) + t0 : nullptr;
        { ProfScope ps(ctx, KC_SELECT_HIST);
        sel_hist_kernel<0><<<grid, SEL_THREADS, SEL_BINS * sizeof(unsigned), ctx->stream>>>(s0, n, nr, use_abs, st0, h0); }
        B4D_LAUNCH_CHECK(ctx);
        { ProfScope ps(ctx, KC_SELECT_SCAN);
        sel_scan_kernel<0><<<tc, SEL_THREADS, 0, ctx->stream>>>(h0, nr, n_q, q_dev, use_abs, st0, o0, nv0); }
        B4D_LAUNCH_CHECK(ctx);
        { ProfScope ps(ctx, KC_SELECT_HIST);
        sel_hist_kernel<1><<<grid, SEL_THREADS, nr * SEL_BINS * sizeof(unsigned), ctx->stream>>>(s0, n, nr, use_abs, st0, h0); }
        B4D_LAUNCH_CHECK(ctx);
        { ProfScope ps(ctx, KC_SELECT_SCAN);
        sel_scan_kernel<1><<<tc, SEL_THREADS, 0, ctx->stream>>>(h0, nr, n_q, q_dev, use_abs, st0, o0, nv0); }
        B4D_LAUNCH_CHECK(ctx);
        { ProfScope ps(ctx, KC_SELECT_HIST);
        sel_hist_kernel<2><<<grid, SEL_THREADS, nr * SEL_BINS * sizeof(unsigned), ctx->stream>>>(s0, n, nr, use_abs, st0, h0); }
        B4D_LAUNCH_CHECK(ctx);
        { ProfScope ps(ctx, KC_SELECT_SCAN);
        sel_scan_kernel<2><<<tc, SEL_THREADS, 0, ctx->stream>>>(h0, nr, n_q, q_dev, use_abs, st0, o0, nv0); }
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

extern "C" int b4d_select_ranks(b4d_ctx* ctx, const float* stack, int64_t n_frames, int64_t frame_elems,
                                const double* quantiles_host, int n_q, int use_abs, float* out, int64_t* n_valid) {
    if (!ctx) return B4D_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->lock);
    if (!stack || !out || !quantiles_host || n_frames < 1 || frame_elems < 1 || n_q < 1 || 2 * n_q > SEL_MAXR)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_select_ranks: bad arguments (n_q must be 1..%d)", SEL_MAXR / 2);
    for (int i = 0; i < n_q; ++i)
        if (!(quantiles_host[i] >= 0.0 && quantiles_host[i] <= 1.0))
            return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_select_ranks: quantile %d outside [0, 1]", i);
    void* p = nullptr;
    int rc = b4d_scratch(ctx, SCR_MISC, 64 * sizeof(double), &p);
    if (rc) return rc;
    B4D_CUDA(ctx, cudaMemcpyAsync(p, quantiles_host, n_q * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return b4d_select_impl(ctx, stack, n_frames, frame_elems, static_cast<const double*>(p), n_q, use_abs, out, n_valid);
}
