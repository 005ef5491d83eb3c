// Exact order statistics per frame.
//
// Behind np.nanpercentile(img, 0.05 / 99.95) in amplitude() (metrics/speckles.py:647 via
// utils/range.py:51-54), np.median(|corr|) in the tracker's SNR (signal/tracking.py:319) and the
// two medians of flat_field_correction (preprocessing/normalize.py:110,126).
//
// Fast path (frames of >= 64 K elements): ONE streaming pass per frame.
//   1. sel_sample_kernel : 16384 strided samples per frame are sorted in shared memory; for every requested
//      quantile a key bracket [L, U] is taken +-6 sigma (binomial) around the sample quantile.
//   2. sel_collect_kernel: one pass over the frame counts the valid (non-NaN) elements, the elements below L,
//      and appends the elements inside [L, U] (0.2 % .. 5 % of the frame) to a candidate list.
//   3. sel_final_kernel  : the two target ranks (numpy's 'linear' method: floor(h), floor(h)+1) are located
//      inside the candidate list by a shared-memory radix select. If a rank falls outside the bracket (or the
//      list overflowed) the frame is flagged and
// Fallback (flagged frames and small frames): three histogram passes over the frame (11 + 11 + 10 key bits)
// with a one-CTA scan after each. Both paths are exact.
// Keys are order-preserving uint32 images of the floats; NaNs are skipped (nanpercentile semantics).
#include "common.cuh"
#include "select.cuh"

namespace {

__device__ __forceinline__ void hist_add(unsigned* h, unsigned bin) {
    const unsigned m = __match_any_sync(__activemask(), bin);
    if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) atomicAdd(h + bin, (unsigned)__popc(m));
}

// =================================================================================================
// fast path
// =================================================================================================
// Rank r among c keys that all lie in [L, U]: whole CTA (1024 threads) cooperates. Digits are taken from
// k - L, which is spread almost uniformly over [0, U - L] (the bracket is a narrow slice of the density), so the
// shared-memory histogram sees no hot bin. Returns the key.
__device__ unsigned cta_bracket_select(const unsigned* __restrict__ keys, unsigned c, unsigned r, unsigned L, unsigned U,
                                       unsigned* hist /*2048 smem*/, unsigned* s_tmp /*34 smem*/) {
    const unsigned range = U - L;                        // keys' offsets are in [0, range]
    int width = 32 - __clz(range | 1u);                  // bits needed for an offset
    unsigned base = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    while (true) {
        const int s = width > 11 ? width - 11 : 0;
        for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        for (unsigned i = threadIdx.x; i < c; i += blockDim.x) {
            const unsigned long long off = (unsigned long long)(keys[i] - L) - (unsigned long long)base;
            if (keys[i] - L >= base && (off >> width) == 0ull) atomicAdd(&hist[(unsigned)(off >> s)], 1u);
        }
        __syncthreads();
        // block-wide exclusive scan of the 2048 bins (2 per thread), then locate the bin holding rank r
        const unsigned h0 = hist[2 * threadIdx.x], h1 = hist[2 * threadIdx.x + 1];
        unsigned incl = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_tmp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_tmp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            s_tmp[lane] = w;                              // inclusive totals per warp
        }
        __syncthreads();
        const unsigned before = (warp ? s_tmp[warp - 1] : 0u) + incl - (h0 + h1);   // exclusive prefix of bin 2*tid
        if (r >= before && r < before + h0) { s_tmp[32] = 2 * threadIdx.x; s_tmp[33] = r - before; }
        else if (r >= before + h0 && r < before + h0 + h1) { s_tmp[32] = 2 * threadIdx.x + 1; s_tmp[33] = r - before - h0; }
        __syncthreads();
        base += s_tmp[32] << s;
        r = s_tmp[33];
        width = s;
        __syncthreads();
        if (s == 0) break;
    }
    return base + L;
}

__global__ void __launch_bounds__(1024) sel_sample_kernel(const float* __restrict__ stack, int64_t n, int n_q,
                                                          const double* __restrict__ quant, int use_abs,
                                                          SelFast* __restrict__ st) {
    extern __shared__ unsigned keys[];              // SEL_SAMPLES
    __shared__ unsigned hist[SEL_BINS];
    __shared__ unsigned tmp[34];
    __shared__ unsigned s_min, s_max;
    __shared__ int s_valid;
    const int64_t t = blockIdx.x;
    const float* f = stack + t * n;
    if (threadIdx.x == 0) { s_valid = 0; s_min = 0xffffffffu; s_max = 0u; }
    __syncthreads();
    int nv = 0;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = threadIdx.x; i < SEL_SAMPLES; i += blockDim.x) {
        const int64_t p = (int64_t)(((__int128)i * n) / SEL_SAMPLES);
        const float v = __ldg(f + p);
        const bool ok = v == v;
        const unsigned k = ok ? key_of(v, use_abs) : 0xffffffffu;     // NaNs sort last (and are never selected)
        keys[i] = k;
        nv += ok;
        if (ok) { kmin = min(kmin, k); kmax = max(kmax, k); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_valid, nv); atomicMin(&s_min, kmin); atomicMax(&s_max, kmax); }
    __syncthreads();
    const int m = s_valid;
    const unsigned lo_key = s_min, hi_key = s_max;
    // bracket [L, U] = sample order statistics +-6 sigma (binomial) around each quantile, found by selection
    unsigned Lq[SEL_MAXQ], Uq[SEL_MAXQ];
    for (int q = 0; q < SEL_MAXQ; ++q) { Lq[q] = 0u; Uq[q] = 0xfffffffeu; }
    if (m >= 1024) {
        for (int q = 0; q < n_q; ++q) {
            const double qq = quant[q];
            const double c = qq * (double)(m - 1);
            const double d = 6.0 * sqrt(qq * (1.0 - qq) * (double)m) + 8.0;
            const long long il = (long long)floor(c - d), iu = (long long)ceil(c + d);
            if (il > 0) Lq[q] = cta_bracket_select(keys, SEL_SAMPLES, (unsigned)il, lo_key, hi_key, hist, tmp);
            if (iu < m - 1) Uq[q] = cta_bracket_select(keys, SEL_SAMPLES, (unsigned)iu, lo_key, hi_key, hist, tmp);
        }
    }
    if (threadIdx.x == 0) {
        SelFast s;
        for (int q = 0; q < SEL_MAXQ; ++q) { s.L[q] = Lq[q]; s.U[q] = Uq[q]; s.below[q] = 0ull; s.ncand[q] = 0u; }
        s.n_valid = 0ull;
        s.need_fallback = 0;
        st[t] = s;
    }
}

// One pass over the frame. A CTA works through chunks of COL_CHUNK elements and keeps, per bracket, the number of
// elements below it and a shared-memory stage of the candidates' keys: the lanes of a warp that hold a candidate are
// compacted by one ballot and one shared atomic per element slot. At the end the stage is flushed with ONE global
// reservation per bracket; the flush also counts the candidates into a 2048-bin histogram over the bracket, which
// lets sel_final_kernel go straight to the bins that hold the wanted ranks.
constexpr int COL_THREADS = 256;
constexpr int COL_VEC = 4;
constexpr int COL_ITERS = 4;
constexpr int COL_CHUNK = COL_THREADS * COL_VEC * COL_ITERS;     // 4096 elements
constexpr int COL_STAGE = 4096;                                  // staged candidates per CTA and bracket

__global__ void __launch_bounds__(COL_THREADS) sel_collect_kernel(const float* __restrict__ stack, int64_t n, int n_q,
                                                                  int use_abs, SelFast* __restrict__ st,
                                                                  unsigned* __restrict__ cand, unsigned cap,
                                                                  unsigned* __restrict__ hist) {
    __shared__ unsigned s_keys[SEL_MAXQ][COL_STAGE];
    __shared__ unsigned s_n[SEL_MAXQ], s_base[SEL_MAXQ];
    __shared__ unsigned long long tot[3];
    const int64_t t = blockIdx.y;
    const float* f = stack + t * n;
    SelFast* s = st + t;
    float Lf[SEL_MAXQ], Uf[SEL_MAXQ];
#pragma unroll
    for (int q = 0; q < SEL_MAXQ; ++q) {
        // L = 0 / U = 0xfffffffe are the open ends of a bracket
        Lf[q] = s->L[q] == 0u ? -INFINITY : value_of(s->L[q], use_abs);
        Uf[q] = s->U[q] == 0xfffffffeu ? INFINITY : value_of(s->U[q], use_abs);
    }
    unsigned nvalid = 0, below[SEL_MAXQ] = {0u, 0u};
    const int lane = threadIdx.x & 31;
    const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
    if (threadIdx.x < 3) tot[threadIdx.x] = 0ull;
    if (threadIdx.x < SEL_MAXQ) s_n[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t nchunks = (n + COL_CHUNK - 1) / COL_CHUNK;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t e0 = ch * COL_CHUNK;
        float v[COL_ITERS][COL_VEC];
#pragma unroll
        for (int it = 0; it < COL_ITERS; ++it) {
            const int64_t e = e0 + ((int64_t)it * COL_THREADS + threadIdx.x) * COL_VEC;
            if (vec && e + COL_VEC <= n) {
                const float4 x = __ldcs(reinterpret_cast<const float4*>(f + e));
                v[it][0] = x.x; v[it][1] = x.y; v[it][2] = x.z; v[it][3] = x.w;
            } else {
#pragma unroll
                for (int k = 0; k < COL_VEC; ++k) v[it][k] = (e + k < n) ? __ldcs(f + e + k) : __uint_as_float(0x7fc00000u);
            }
        }
        if (use_abs) {
#pragma unroll
            for (int it = 0; it < COL_ITERS; ++it)
#pragma unroll
                for (int k = 0; k < COL_VEC; ++k) v[it][k] = fabsf(v[it][k]);
        }
        // NaN census per quad: the sum of magnitudes is NaN iff one of them is
#pragma unroll
        for (int it = 0; it < COL_ITERS; ++it) {
            const float sm4 = (fabsf(v[it][0]) + fabsf(v[it][1])) + (fabsf(v[it][2]) + fabsf(v[it][3]));
            if (sm4 == sm4) nvalid += 4;
            else nvalid += (v[it][0] == v[it][0]) + (v[it][1] == v[it][1]) + (v[it][2] == v[it][2]) + (v[it][3] == v[it][3]);
        }
#pragma unroll
        for (int q = 0; q < SEL_MAXQ; ++q) {
            if (q >= n_q) continue;
            // per-lane census, one warp scan and one shared atomic per chunk and bracket, then the lane's own writes
            const float L = Lf[q], U = Uf[q];
            unsigned c = 0;
#pragma unroll
            for (int it = 0; it < COL_ITERS; ++it)
#pragma unroll
                for (int k = 0; k < COL_VEC; ++k) {
                    const float x = v[it][k];
                    below[q] += (x < L);
                    c += (x >= L && x <= U);
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
                if (lane == 31) base = atomicAdd(&s_n[q], wtot);
                base = __shfl_sync(0xffffffffu, base, 31);
                if (c) {
                    unsigned pos = base + incl - c;
#pragma unroll
                    for (int it = 0; it < COL_ITERS; ++it)
#pragma unroll
                        for (int k = 0; k < COL_VEC; ++k) {
                            const float x = v[it][k];
                            if (x >= L && x <= U) {
                                if (pos < COL_STAGE) s_keys[q][pos] = key_of(x, use_abs);
                                ++pos;
                            }
                        }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        below[0] += __shfl_xor_sync(0xffffffffu, below[0], o);
        below[1] += __shfl_xor_sync(0xffffffffu, below[1], o);
    }
    if (lane == 0) {
        atomicAdd(&tot[0], (unsigned long long)nvalid);
        atomicAdd(&tot[1], (unsigned long long)below[0]);
        atomicAdd(&tot[2], (unsigned long long)below[1]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&s->n_valid, tot[0]);
        atomicAdd(&s->below[0], tot[1]);
        atomicAdd(&s->below[1], tot[2]);
    }
    // flush: one reservation per bracket, keys + their histogram
    if (threadIdx.x < SEL_MAXQ && threadIdx.x < n_q) {
        const int q = threadIdx.x;
        const unsigned cnt = s_n[q];
        unsigned base = 0;
        if (cnt > COL_STAGE) s->need_fallback = 1;
        else if (cnt) base = atomicAdd(&s->ncand[q], cnt);
        s_base[q] = base;
    }
    __syncthreads();
    for (int q = 0; q < n_q; ++q) {
        const unsigned cnt = min(s_n[q], (unsigned)COL_STAGE), base = s_base[q];
        const unsigned L = s->L[q];
        const int shift = bracket_shift(L, s->U[q]);
        unsigned* dst = cand + ((size_t)t * SEL_MAXQ + q) * cap;
        unsigned* h = hist + ((size_t)t * SEL_MAXQ + q) * SEL_BINS;
        for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
            const unsigned key = s_keys[q][i];
            if (base + i < cap) dst[base + i] = key;
            atomicAdd(h + ((key - L) >> shift), 1u);
        }
    }
}

// Final ranks of one frame from its candidate lists. The histogram of the bracket names the bin of each wanted rank;
// one pass over the candidates gathers the keys of those bins (a few dozen) into shared memory, where the ranks are
// resolved by the bracket select. Frames whose bracket missed (or whose lists / bins overflowed) are flagged.
constexpr int FIN_CAP = 4096;
__global__ void __launch_bounds__(1024) sel_final_kernel(SelFast* __restrict__ st, const unsigned* __restrict__ cand, unsigned cap,
                                                         const unsigned* __restrict__ hist, int n_q,
                                                         const double* __restrict__ quant, int use_abs,
                                                         float* __restrict__ out, long long* __restrict__ n_valid_out,
                                                         int* __restrict__ need) {
    __shared__ unsigned shist[SEL_BINS];
    __shared__ unsigned bc[34];
    __shared__ unsigned s_keys[FIN_CAP];
    __shared__ unsigned s_n;
    __shared__ unsigned s_bin[2], s_before[2];
    const int64_t t = blockIdx.x;
    SelFast* s = st + t;
    const unsigned long long nv = s->n_valid;
    bool ok = nv > 0 && !s->need_fallback;
    long long lo[SEL_MAXQ], hi[SEL_MAXQ];
    for (int q = 0; q < n_q; ++q) {
        target_ranks(nv, quant[q], lo[q], hi[q]);
        const long long b = (long long)s->below[q], c = (long long)s->ncand[q];
        ok = ok && c <= (long long)cap && lo[q] >= b && hi[q] < b + c;
    }
    float res[2 * SEL_MAXQ];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = 0; q < n_q && ok; ++q) {
        const unsigned L = s->L[q], U = s->U[q], c = s->ncand[q];
        const int shift = bracket_shift(L, U);
        const unsigned ra = (unsigned)(lo[q] - (long long)s->below[q]), rb = (unsigned)(hi[q] - (long long)s->below[q]);
        // exclusive scan of the bracket histogram (2 bins per thread), locate the bins of both ranks
        const unsigned* h = hist + ((size_t)t * SEL_MAXQ + q) * SEL_BINS;
        const unsigned h0 = h[2 * threadIdx.x], h1 = h[2 * threadIdx.x + 1];
        unsigned incl = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        __syncthreads();
        if (lane == 31) bc[warp] = incl;
        if (threadIdx.x == 0) s_n = 0u;
        __syncthreads();
        if (warp == 0) {
            unsigned w = bc[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            bc[lane] = w;
        }
        __syncthreads();
        const unsigned before = (warp ? bc[warp - 1] : 0u) + incl - (h0 + h1);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned rk = r ? rb : ra;
            if (rk >= before && rk < before + h0) { s_bin[r] = 2 * threadIdx.x; s_before[r] = before; }
            else if (rk >= before + h0 && rk < before + h0 + h1) { s_bin[r] = 2 * threadIdx.x + 1; s_before[r] = before + h0; }
        }
        __syncthreads();
        const unsigned bin_a = s_bin[0], bin_b = s_bin[1];
        unsigned ka, kb;
        if (shift == 0) {                     // one key per bin: the bin is the answer
            ka = L + bin_a;
            kb = L + bin_b;
        } else {
            const unsigned* keys = cand + ((size_t)t * SEL_MAXQ + q) * cap;
            for (unsigned i = threadIdx.x; i < c; i += blockDim.x) {
                const unsigned key = keys[i], b = (key - L) >> shift;
                if (b == bin_a || b == bin_b) {
                    const unsigned p = atomicAdd(&s_n, 1u);
                    if (p < FIN_CAP) s_keys[p] = key;
                }
            }
            __syncthreads();
            const unsigned m = s_n;
            if (m > FIN_CAP) { ok = false; break; }      // heavy ties inside one bin: the radix path resolves them
            // ranks inside the gathered set: the keys of bin_a precede those of bin_b
            const unsigned la = ra - s_before[0];
            const unsigned lb = bin_b == bin_a ? rb - s_before[0] : h[bin_a] + (rb - s_before[1]);
            const unsigned Lb = L + (bin_a << shift), Ub = L + (((bin_b + 1u) << shift) - 1u);
            ka = cta_bracket_select(s_keys, m, la, Lb, Ub < Lb ? 0xffffffffu : Ub, shist, bc);
            kb = (rb == ra) ? ka : cta_bracket_select(s_keys, m, lb, Lb, Ub < Lb ? 0xffffffffu : Ub, shist, bc);
        }
        res[2 * q] = value_of(ka, use_abs);
        res[2 * q + 1] = value_of(kb, use_abs);
    }
    if (threadIdx.x == 0) {
        need[t] = ok ? 0 : 1;
        if (n_valid_out) n_valid_out[t] = (long long)nv;
        if (ok)
            for (int i = 0; i < 2 * n_q; ++i) out[t * 2 * n_q + i] = res[i];
    }
}

// =================================================================================================
// fused median of a map that is never materialised (the tracker's |corr|, signal/tracking.py:319)
// =================================================================================================
// rows_inv_kernel first produces a few sample rows of every frame; this kernel turns them into the bracket [L, U]
// around the wanted quantile (same +-6 sigma rule as sel_sample_kernel) and takes the census of the sample rows
// themselves. The main rows_inv launch then counts / collects every other value against the bracket in its epilogue
// (census_values_global) and sel_final_kernel reads the exact order statistics from the candidates.
__global__ void __launch_bounds__(1024) sel_bracket_samples_kernel(const float* __restrict__ samples, int m,
                                                                   const double* __restrict__ quant,
                                                                   SelFast* __restrict__ st, unsigned* __restrict__ cand,
                                                                   int regions, unsigned* __restrict__ bhist) {
    extern __shared__ unsigned keys[];              // m
    __shared__ unsigned hist[SEL_BINS];
    __shared__ unsigned tmp[34];
    __shared__ unsigned s_min, s_max, s_cnt, s_below;
    __shared__ int s_valid;
    const int64_t t = blockIdx.x;
    const float* f = samples + t * m;
    if (threadIdx.x == 0) { s_valid = 0; s_min = 0xffffffffu; s_max = 0u; s_cnt = 0u; s_below = 0u; }
    __syncthreads();
    int nv = 0;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const float v = f[i];
        const bool ok = v == v;
        const unsigned k = ok ? key_of(v, 1) : 0xffffffffu;
        keys[i] = k;
        nv += ok;
        if (ok) { kmin = min(kmin, k); kmax = max(kmax, k); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_valid, nv); atomicMin(&s_min, kmin); atomicMax(&s_max, kmax); }
    __syncthreads();
    const int mv = s_valid;
    unsigned L = 0u, U = 0xfffffffeu;
    if (mv >= 1024) {
        const double qq = quant[0];
        const double c = qq * (double)(mv - 1);
        const double d = 6.0 * sqrt(qq * (1.0 - qq) * (double)mv) + 8.0;
        const long long il = (long long)floor(c - d), iu = (long long)ceil(c + d);
        if (il > 0) L = cta_bracket_select(keys, (unsigned)m, (unsigned)il, s_min, s_max, hist, tmp);
        if (iu < mv - 1) U = cta_bracket_select(keys, (unsigned)m, (unsigned)iu, s_min, s_max, hist, tmp);
    }
    // census of the sample rows themselves: their keys go to the tail of the frame's candidate store
    const int shift = bracket_shift(L, U);
    unsigned* dst = cand + (size_t)t * ((size_t)regions * FM_REGION + FM_SAMPLE_CAP) + (size_t)regions * FM_REGION;
    unsigned* h = bhist + (size_t)t * SEL_BINS;
    unsigned below = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned k = keys[i];
        below += (k < L);
        if (k >= L && k <= U) {
            const unsigned p = atomicAdd(&s_cnt, 1u);
            if (p < FM_SAMPLE_CAP) dst[p] = k;
            atomicAdd(h + min((k - L) >> shift, (unsigned)(SEL_BINS - 1)), 1u);
        }
    }
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_below, below);
    __syncthreads();
    if (threadIdx.x == 0) {
        SelFast* s = st + t;
        s->L[0] = L; s->U[0] = U;
        s->ncand[0] = s_cnt;
        s->below[0] = (unsigned long long)s_below;
        s->n_valid = (unsigned long long)mv;
    }
}

// Exact ranks of the fused median: totals from the per-region counters, bins of the wanted ranks from the bracket
// histogram, one pass over the region store for the keys of those bins, bracket select among them.
// out[t] = the two middle order statistics; need[t] = 1 where the bracket missed or a region / bin overflowed.
__global__ void __launch_bounds__(1024) fused_median_final_kernel(const SelFast* __restrict__ st, const unsigned* __restrict__ cnt3,
                                                                  const unsigned* __restrict__ cand, int regions,
                                                                  const unsigned* __restrict__ bhist,
                                                                  const double* __restrict__ quant, float* __restrict__ out,
                                                                  long long* __restrict__ n_valid_out, int* __restrict__ need) {
    __shared__ unsigned shist[SEL_BINS];
    __shared__ unsigned bc[34];
    __shared__ unsigned s_keys[FIN_CAP];
    __shared__ unsigned s_n;
    __shared__ unsigned s_bin[2], s_before[2];
    __shared__ unsigned long long s_tot[3];
    __shared__ int s_over;
    const int64_t t = blockIdx.x;
    const SelFast* s = st + t;
    const unsigned* c3 = cnt3 + (size_t)t * regions * 3;
    const unsigned* store = cand + (size_t)t * ((size_t)regions * FM_REGION + FM_SAMPLE_CAP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) { s_over = 0; s_n = 0u; }
    __syncthreads();
    // 1. totals (fixed data, order-independent integer sums)
    unsigned long long a0 = 0, a1 = 0, a2 = 0;
    int over = 0;
    for (int r = threadIdx.x; r < regions; r += blockDim.x) {
        const unsigned c = c3[3 * r];
        a0 += c; a1 += c3[3 * r + 1]; a2 += c3[3 * r + 2];
        over |= c > (unsigned)FM_REGION;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) { atomicAdd(&s_tot[0], a0); atomicAdd(&s_tot[1], a1); atomicAdd(&s_tot[2], a2); }
    if (over) s_over = 1;
    __syncthreads();
    const unsigned sc = s->ncand[0];
    const unsigned long long ncand = s_tot[0] + sc, below = s_tot[1] + s->below[0], nv = s_tot[2] + s->n_valid;
    bool ok = nv > 0 && !s_over && sc <= (unsigned)FM_SAMPLE_CAP;
    long long lo, hi;
    target_ranks(nv, quant[0], lo, hi);
    ok = ok && lo >= (long long)below && hi < (long long)(below + ncand);
    float res[2] = {0.f, 0.f};
    if (ok) {
        const unsigned L = s->L[0], U = s->U[0];
        const int shift = bracket_shift(L, U);
        const unsigned ra = (unsigned)(lo - (long long)below), rb = (unsigned)(hi - (long long)below);
        const unsigned* h = bhist + (size_t)t * SEL_BINS;
        const unsigned h0 = h[2 * threadIdx.x], h1 = h[2 * threadIdx.x + 1];
        unsigned incl = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) bc[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = bc[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            bc[lane] = w;
        }
        __syncthreads();
        const unsigned before = (warp ? bc[warp - 1] : 0u) + incl - (h0 + h1);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned rk = r ? rb : ra;
            if (rk >= before && rk < before + h0) { s_bin[r] = 2 * threadIdx.x; s_before[r] = before; }
            else if (rk >= before + h0 && rk < before + h0 + h1) { s_bin[r] = 2 * threadIdx.x + 1; s_before[r] = before + h0; }
        }
        __syncthreads();
        const unsigned bin_a = s_bin[0], bin_b = s_bin[1];
        unsigned ka, kb;
        if (shift == 0) {                     // one key per bin: the bin is the answer
            ka = L + bin_a;
            kb = L + bin_b;
        } else {
            auto take = [&](unsigned key) {
                const unsigned b = min((key - L) >> shift, (unsigned)(SEL_BINS - 1));
                if (b == bin_a || b == bin_b) {
                    const unsigned p = atomicAdd(&s_n, 1u);
                    if (p < FIN_CAP) s_keys[p] = key;
                }
            };
            // a warp walks whole CTA regions (256 per frame at 2048^2, ~560 keys each), 128 keys per step with all four
            // loads in flight together
            for (int r = warp; r < regions; r += 32) {
                const unsigned c = min(c3[3 * r], (unsigned)FM_REGION);
                const unsigned* reg = store + (size_t)r * FM_REGION;
                for (unsigned i0 = 0; i0 < c; i0 += 128) {
                    unsigned k[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) k[u] = i0 + 32 * u + lane < c ? reg[i0 + 32 * u + lane] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i0 + 32 * u + lane < c) take(k[u]);
                }
            }
            for (unsigned i = threadIdx.x; i < sc; i += blockDim.x) take(store[(size_t)regions * FM_REGION + i]);
            __syncthreads();
            const unsigned m = s_n;
            if (m > FIN_CAP) ok = false;      // heavy ties inside one bin: the map-based path resolves them
            else {
                const unsigned la = ra - s_before[0];
                const unsigned lb = bin_b == bin_a ? rb - s_before[0] : h[bin_a] + (rb - s_before[1]);
                const unsigned Lb = L + (bin_a << shift), Ub = L + (((bin_b + 1u) << shift) - 1u);
                ka = cta_bracket_select(s_keys, m, la, Lb, Ub < Lb ? 0xffffffffu : max(Ub, U), shist, bc);
                kb = (rb == ra) ? ka : cta_bracket_select(s_keys, m, lb, Lb, Ub < Lb ? 0xffffffffu : max(Ub, U), shist, bc);
            }
        }
        if (ok) { res[0] = __uint_as_float(ka); res[1] = __uint_as_float(kb); }
    }
    if (threadIdx.x == 0) {
        need[t] = ok ? 0 : 1;
        n_valid_out[t] = (long long)nv;
        out[2 * t] = ok ? res[0] : __uint_as_float(0x7fc00000u);
        out[2 * t + 1] = ok ? res[1] : __uint_as_float(0x7fc00000u);
    }
}

// =================================================================================================
// tails: thresholds for the fused tail collection of frame_reduce2_kernel, and the final order statistics
// =================================================================================================
// thr[t] = (thr_lo, thr_hi): every pixel < thr_lo / > thr_hi is a candidate, the pixels equal to a threshold are counted
// (ties: masked, dark or saturated regions of integer detector frames). They are the sample order statistics
// +-6 sigma (binomial) beyond the sample quantiles, so the wanted ranks fall inside the candidate lists unless the
// sample was unlucky (then tails_final_kernel flags the frame).
__global__ void __launch_bounds__(1024) tails_probe_kernel(const float* __restrict__ stack, const float* __restrict__ gain,
                                                           const float* __restrict__ dark, int64_t n, double q_lo, double q_hi,
                                                           float* __restrict__ thr) {
    extern __shared__ unsigned keys[];              // SEL_SAMPLES
    __shared__ unsigned hist[SEL_BINS];
    __shared__ unsigned tmp[34];
    __shared__ unsigned s_min, s_max;
    __shared__ int s_valid;
    const int64_t t = blockIdx.x;
    const float* f = stack + t * n;
    if (threadIdx.x == 0) { s_valid = 0; s_min = 0xffffffffu; s_max = 0u; }
    __syncthreads();
    int nv = 0;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = threadIdx.x; i < SEL_SAMPLES; i += blockDim.x) {
        const int64_t p = (int64_t)(((__int128)i * n) / SEL_SAMPLES);
        float v = __ldg(f + p);
        if (gain) v = (v - (dark ? __ldg(dark + p) : 0.f)) * __ldg(gain + p);
        const bool ok = v == v;
        const unsigned k = ok ? key_of(v, 0) : 0xffffffffu;
        keys[i] = k;
        nv += ok;
        if (ok) { kmin = min(kmin, k); kmax = max(kmax, k); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_valid, nv); atomicMin(&s_min, kmin); atomicMax(&s_max, kmax); }
    __syncthreads();
    const int m = s_valid;
    float lo = INFINITY, hi = -INFINITY;            // degenerate sample: everything is a candidate -> overflow -> flagged
    if (m >= 1024) {
        const double cl = q_lo * (double)(m - 1), dl = 6.0 * sqrt(q_lo * (1.0 - q_lo) * (double)m) + 8.0;
        const double ch = q_hi * (double)(m - 1), dh = 6.0 * sqrt(q_hi * (1.0 - q_hi) * (double)m) + 8.0;
        const long long iu = (long long)ceil(cl + dl), il = (long long)floor(ch - dh);
        if (iu < m - 1) lo = value_of(cta_bracket_select(keys, SEL_SAMPLES, (unsigned)iu, s_min, s_max, hist, tmp), 0);
        if (il > 0) hi = value_of(cta_bracket_select(keys, SEL_SAMPLES, (unsigned)il, s_min, s_max, hist, tmp), 0);
    }
    if (threadIdx.x == 0) { thr[2 * t] = lo; thr[2 * t + 1] = hi; }
}

constexpr unsigned TAIL_GCAP = 16384;   // candidates per frame and tail (64 KB of keys in shared memory)

// out[t] = (v[lo], v[hi]) of q_lo, (v[lo], v[hi]) of q_hi (numpy 'linear' neighbours); nvalid_out[t] = number of
// non-NaN pixels, or -1 when the frame's candidate lists do not contain the wanted ranks.
__global__ void __launch_bounds__(1024) tails_final_kernel(const float* __restrict__ cand, const unsigned* __restrict__ cnt,
                                                           const unsigned* __restrict__ eq, const float* __restrict__ thr,
                                                           const int* __restrict__ flag, const double* __restrict__ fr,
                                                           double q_lo, double q_hi, float* __restrict__ out,
                                                           long long* __restrict__ nvalid_out) {
    extern __shared__ unsigned keys[];              // TAIL_GCAP
    __shared__ unsigned hist[SEL_BINS];
    __shared__ unsigned tmp[34];
    __shared__ unsigned s_min, s_max;
    const int64_t t = blockIdx.x;
    const long long nv = (long long)(fr[t * B4D_FR_NCOLS + B4D_FR_NPIX] - fr[t * B4D_FR_NCOLS + B4D_FR_NNAN]);
    bool ok = nv > 0 && !flag[t];
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tail = 0; tail < 2 && ok; ++tail) {
        // ascending order of a tail: lower = [c candidates < thr][e ties == thr] ...; upper = ... [e ties == thr][c candidates > thr]
        const unsigned c = cnt[2 * t + tail], e = eq[2 * t + tail];
        const float tv = thr[2 * t + tail];
        long long lo, hi;
        target_ranks((unsigned long long)nv, tail ? q_hi : q_lo, lo, hi);
        // ascending ranks inside the candidate list; the run of ties sits right after it (lower) / right before it (upper)
        const long long off = tail ? nv - (long long)c : 0;
        lo -= off; hi -= off;
        bool need_lo, need_hi, tie_lo, tie_hi;
        if (!tail) {
            need_lo = lo < (long long)c; need_hi = hi < (long long)c;
            tie_lo = !need_lo && lo < (long long)c + (long long)e; tie_hi = !need_hi && hi < (long long)c + (long long)e;
        } else {
            need_lo = lo >= 0; need_hi = hi >= 0;
            tie_lo = !need_lo && lo >= -(long long)e; tie_hi = !need_hi && hi >= -(long long)e;
        }
        if (c > TAIL_GCAP || lo < -(long long)e || hi >= (long long)c + (tail ? 0 : (long long)e) || !(need_lo || tie_lo) || !(need_hi || tie_hi)) {
            ok = false;
            break;
        }
        unsigned ka = 0u, kb = 0u;
        if (need_lo || need_hi) {
            __syncthreads();
            if (threadIdx.x == 0) { s_min = 0xffffffffu; s_max = 0u; }
            __syncthreads();
            unsigned kmin = 0xffffffffu, kmax = 0u;
            const float* src = cand + ((size_t)t * 2 + tail) * TAIL_GCAP;
            for (unsigned i = threadIdx.x; i < c; i += blockDim.x) {
                const unsigned k = key_of(src[i], 0);
                keys[i] = k;
                kmin = min(kmin, k); kmax = max(kmax, k);
            }
            for (int o = 16; o > 0; o >>= 1) {
                kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
                kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
            }
            if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, kmin); atomicMax(&s_max, kmax); }
            __syncthreads();
            if (need_lo) ka = cta_bracket_select(keys, c, (unsigned)lo, s_min, s_max, hist, tmp);
            if (need_hi) kb = (need_lo && hi == lo) ? ka : cta_bracket_select(keys, c, (unsigned)hi, s_min, s_max, hist, tmp);
        }
        res[2 * tail] = need_lo ? value_of(ka, 0) : tv;
        res[2 * tail + 1] = need_hi ? value_of(kb, 0) : tv;
    }
    if (threadIdx.x == 0) {
        nvalid_out[t] = ok ? nv : -1;
        for (int i = 0; i < 4; ++i) out[4 * t + i] = ok ? res[i] : __uint_as_float(0x7fc00000u);
    }
}

// =================================================================================================
// radix path (small frames, flagged frames)
// =================================================================================================
// PASS 0: bits 31..21, PASS 1: bits 20..10 (given 11-bit prefix), PASS 2: bits 9..0 (given 22-bit prefix)
template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) sel_hist_kernel(const float* __restrict__ stack, int64_t n,
                                                               int nr, int use_abs,
                                                               const SelState* __restrict__ st,
                                                               unsigned* __restrict__ hist, const int* __restrict__ need) {
    extern __shared__ unsigned sh[];                      // nr_eff * SEL_BINS
    const int64_t t = blockIdx.y;
    if (need && !need[t]) return;
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
                                                               long long* __restrict__ n_valid_out,
                                                               const int* __restrict__ need) {
    const int64_t t = blockIdx.x;
    if (need && !need[t]) return;
    unsigned* g = hist + (size_t)t * SEL_MAXR * SEL_BINS;
    __shared__ unsigned long long cum[SEL_BINS];
    __shared__ unsigned long long part[SEL_THREADS];
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
        unsigned long long loc[8], run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { run += h[threadIdx.x * 8 + i]; loc[i] = run; }
        part[threadIdx.x] = run;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long a = 0;
            for (int i = 0; i < SEL_THREADS; ++i) { unsigned long long v = part[i]; part[i] = a; a += v; }
            if (PASS == 0 && r == 0) {
                s.n_valid = (long long)a;
                for (int q = 0; q < n_q; ++q) {
                    long long lo, hi;
                    target_ranks(a, quant[q], lo, hi);
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
    for (int i = threadIdx.x; i < SEL_MAXR * SEL_BINS; i += blockDim.x) g[i] = 0;   // clear for the next pass
}

int radix_path(b4d_ctx* ctx, const float* stack, int64_t T, int64_t n, const double* q_dev, int n_q, int use_abs, float* out,
               long long* n_valid, const int* need, unsigned* hist, SelState* st) {
    const int nr = 2 * n_q;
    int bpf = (int)((n + SEL_THREADS * 16 - 1) / (SEL_THREADS * 16));
    const int cap = ctx->sm_count * 8;
    if ((int64_t)bpf * T > cap) bpf = (int)((cap + T - 1) / T);
    if (bpf < 1) bpf = 1;
    dim3 grid((unsigned)bpf, (unsigned)T);
    const unsigned tc = (unsigned)T;
    { ProfScope ps(ctx, KC_SELECT_HIST);
      sel_hist_kernel<0><<<grid, SEL_THREADS, SEL_BINS * sizeof(unsigned), ctx->stream>>>(stack, n, nr, use_abs, st, hist, need); }
    B4D_LAUNCH_CHECK(ctx);
    { ProfScope ps(ctx, KC_SELECT_SCAN);
      sel_scan_kernel<0><<<tc, SEL_THREADS, 0, ctx->stream>>>(hist, nr, n_q, q_dev, use_abs, st, out, n_valid, need); }
    B4D_LAUNCH_CHECK(ctx);
    { ProfScope ps(ctx, KC_SELECT_HIST);
      sel_hist_kernel<1><<<grid, SEL_THREADS, nr * SEL_BINS * sizeof(unsigned), ctx->stream>>>(stack, n, nr, use_abs, st, hist, need); }
    B4D_LAUNCH_CHECK(ctx);
    { ProfScope ps(ctx, KC_SELECT_SCAN);
      sel_scan_kernel<1><<<tc, SEL_THREADS, 0, ctx->stream>>>(hist, nr, n_q, q_dev, use_abs, st, out, n_valid, need); }
    B4D_LAUNCH_CHECK(ctx);
    { ProfScope ps(ctx, KC_SELECT_HIST);
      sel_hist_kernel<2><<<grid, SEL_THREADS, nr * SEL_BINS * sizeof(unsigned), ctx->stream>>>(stack, n, nr, use_abs, st, hist, need); }
    B4D_LAUNCH_CHECK(ctx);
    { ProfScope ps(ctx, KC_SELECT_SCAN);
      sel_scan_kernel<2><<<tc, SEL_THREADS, 0, ctx->stream>>>(hist, nr, n_q, q_dev, use_abs, st, out, n_valid, need); }
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

}  // namespace


// ---- tails (called from reduce.cu) ------------------------------------------------------------------
unsigned b4d_tails_gcap() { return TAIL_GCAP; }

int b4d_tails_probe_launch(b4d_ctx* ctx, const float* stack, int64_t T, int64_t npix, const float* gain, const float* dark,
                           double q_lo, double q_hi, float* thr) {
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) {
        B4D_CUDA(ctx, cudaFuncSetAttribute(tails_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(SEL_SAMPLES * sizeof(unsigned))));
        attr[ctx->device] = true;
    }
    ProfScope ps(ctx, KC_SELECT_SAMPLE);
    tails_probe_kernel<<<(unsigned)T, 1024, SEL_SAMPLES * sizeof(unsigned), ctx->stream>>>(stack, gain, dark, npix, q_lo, q_hi, thr);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int b4d_tails_final_launch(b4d_ctx* ctx, const float* cand, const unsigned* cnt, const unsigned* eq, const float* thr, const int* flag,
                           const double* fr, int64_t T, double q_lo, double q_hi, float* out, int64_t* nvalid_out) {
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) {
        B4D_CUDA(ctx, cudaFuncSetAttribute(tails_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(TAIL_GCAP * sizeof(unsigned))));
        attr[ctx->device] = true;
    }
    ProfScope ps(ctx, KC_SELECT_FINAL);
    tails_final_kernel<<<(unsigned)T, 1024, TAIL_GCAP * sizeof(unsigned), ctx->stream>>>(cand, cnt, eq, thr, flag, fr, q_lo, q_hi, out,
                                                                                      reinterpret_cast<long long*>(nvalid_out));
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

// ---- fused median (called from spectral.cu) -----------------------------------------------------------
// Scratch of one batch: bracket state, fallback flags, bracket histograms and region counters (zeroed), region store.
// A few doubles from the host to device memory as kernel arguments: a cudaMemcpyAsync from pageable memory would queue
// on the copy engine behind whatever bulk host->device transfer another stream has in flight (StackAnalyzer.run
// overlaps the next chunk's upload with these kernels) and stall the whole compute stream for its duration.
__global__ void put_doubles_kernel(double* dst, int n, double a, double b, double c, double d) {
    const double v[4] = {a, b, c, d};
    if (threadIdx.x < n) dst[threadIdx.x] = v[threadIdx.x];
}
int b4d_put_doubles(b4d_ctx* ctx, double* dst, const double* src_host, int n) {
    if (n < 1 || n > 4) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_put_doubles: 1..4 values");
    double v[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) v[i] = src_host[i];
    put_doubles_kernel<<<1, 4, 0, ctx->stream>>>(dst, n, v[0], v[1], v[2], v[3]);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int b4d_fused_median_begin(b4d_ctx* ctx, int64_t T, int regions, FusedMedian* fm) {
    const size_t st_bytes = ((size_t)T * sizeof(SelFast) + 255) & ~size_t(255);
    const size_t need_bytes = ((size_t)T * sizeof(int) + 255) & ~size_t(255);
    const size_t bh_bytes = (size_t)T * SEL_BINS * sizeof(unsigned);
    const size_t c3_bytes = ((size_t)T * regions * 3 * sizeof(unsigned) + 255) & ~size_t(255);
    const size_t cand_bytes = (size_t)T * ((size_t)regions * FM_REGION + FM_SAMPLE_CAP) * sizeof(unsigned);
    void* p = nullptr;
    int rc = b4d_scratch(ctx, SCR_SELECT, 256 + st_bytes + need_bytes + bh_bytes + c3_bytes + cand_bytes, &p);
    if (rc) return rc;
    char* base = static_cast<char*>(p);
    fm->q_dev = reinterpret_cast<double*>(base);
    fm->st = reinterpret_cast<SelFast*>(base + 256);
    fm->need = reinterpret_cast<int*>(base + 256 + st_bytes);
    fm->bhist = reinterpret_cast<unsigned*>(base + 256 + st_bytes + need_bytes);
    fm->cnt3 = reinterpret_cast<unsigned*>(base + 256 + st_bytes + need_bytes + bh_bytes);
    fm->cand = reinterpret_cast<unsigned*>(base + 256 + st_bytes + need_bytes + bh_bytes + c3_bytes);
    fm->regions = regions;
    B4D_CUDA(ctx, cudaMemsetAsync(p, 0, 256 + st_bytes + need_bytes + bh_bytes + c3_bytes, ctx->stream));
    static const double half = 0.5;
    return b4d_put_doubles(ctx, fm->q_dev, &half, 1);
}

int b4d_fused_median_bracket(b4d_ctx* ctx, const FusedMedian& fm, const float* samples, int m, int64_t T) {
    static bool attr[B4D_MAX_DEVICES] = {};
    if (m > 32768) return b4d_fail(ctx, B4D_ERR_INVALID, "fused median: at most 32768 samples per frame");
    if (!attr[ctx->device]) {
        B4D_CUDA(ctx, cudaFuncSetAttribute(sel_bracket_samples_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
        attr[ctx->device] = true;
    }
    ProfScope ps(ctx, KC_SELECT_SAMPLE);
    sel_bracket_samples_kernel<<<(unsigned)T, 1024, (size_t)m * sizeof(unsigned), ctx->stream>>>(samples, m, fm.q_dev, fm.st, fm.cand,
                                                                                                fm.regions, fm.bhist);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

// out (T, 2): the two middle order statistics; nvalid (T); fm.need[t] = 1 where the bracket missed
int b4d_fused_median_final(b4d_ctx* ctx, const FusedMedian& fm, int64_t T, float* out, int64_t* nvalid) {
    ProfScope ps(ctx, KC_SELECT_FINAL);
    fused_median_final_kernel<<<(unsigned)T, 1024, 0, ctx->stream>>>(fm.st, fm.cnt3, fm.cand, fm.regions, fm.bhist, fm.q_dev, out,
                                                                    reinterpret_cast<long long*>(nvalid), fm.need);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int b4d_select_impl(b4d_ctx* ctx, const float* stack, int64_t T, int64_t n, const double* q_dev, int n_q,
                    int use_abs, float* out, int64_t* n_valid) {
    static bool sample_attr[B4D_MAX_DEVICES] = {};
    const bool fast = n >= SEL_FAST_MIN;
    const unsigned cap = fast ? (unsigned)(n / 8) : 0u;
    const int64_t TC = 16384;
    for (int64_t t0 = 0; t0 < T; t0 += TC) {
        const int64_t tc = T - t0 < TC ? T - t0 : TC;
        void* p = nullptr;
        const size_t hist_bytes = (size_t)tc * SEL_MAXR * SEL_BINS * sizeof(unsigned);
        const size_t st_bytes = ((size_t)tc * sizeof(SelState) + 255) & ~size_t(255);
        const size_t fast_bytes = ((size_t)tc * sizeof(SelFast) + 255) & ~size_t(255);
        const size_t need_bytes = ((size_t)tc * sizeof(int) + 255) & ~size_t(255);
        const size_t cand_bytes = (size_t)tc * SEL_MAXQ * cap * sizeof(unsigned);
        const size_t bh_bytes = fast ? (size_t)tc * SEL_MAXQ * SEL_BINS * sizeof(unsigned) : 0;   // bracket histograms
        int rc = b4d_scratch(ctx, SCR_SELECT, hist_bytes + st_bytes + fast_bytes + need_bytes + bh_bytes + cand_bytes, &p);
        if (rc) return rc;
        char* base = static_cast<char*>(p);
        unsigned* hist = reinterpret_cast<unsigned*>(base);
        SelState* st = reinterpret_cast<SelState*>(base + hist_bytes);
        SelFast* sf = reinterpret_cast<SelFast*>(base + hist_bytes + st_bytes);
        int* need = reinterpret_cast<int*>(base + hist_bytes + st_bytes + fast_bytes);
        unsigned* bhist = reinterpret_cast<unsigned*>(base + hist_bytes + st_bytes + fast_bytes + need_bytes);
        unsigned* cand = reinterpret_cast<unsigned*>(base + hist_bytes + st_bytes + fast_bytes + need_bytes + bh_bytes);
        B4D_CUDA(ctx, cudaMemsetAsync(p, 0, hist_bytes + st_bytes + fast_bytes + need_bytes + bh_bytes, ctx->stream));
        const float* s0 = stack + t0 * n;
        float* o0 = out + t0 * 2 * n_q;
        long long* nv0 = n_valid ? reinterpret_cast<long long*>(n_valid) + t0 : nullptr;
        if (fast) {
            if (!sample_attr[ctx->device]) {
                B4D_CUDA(ctx, cudaFuncSetAttribute(sel_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(SEL_SAMPLES * sizeof(unsigned))));
                sample_attr[ctx->device] = true;
            }
            { ProfScope ps(ctx, KC_SELECT_SAMPLE);
              sel_sample_kernel<<<(unsigned)tc, 1024, SEL_SAMPLES * sizeof(unsigned), ctx->stream>>>(s0, n, n_q, q_dev, use_abs, sf); }
            B4D_LAUNCH_CHECK(ctx);
            // CTAs per frame: enough to fill the GPU, and never more than 8 chunks (32 K elements) per CTA -- the widest
            // bracket (the median's, +-6 sigma of a 16 K sample) holds < 5 % of them, well inside the CTA's stage
            int bpf = (int)((n + COL_CHUNK - 1) / COL_CHUNK);
            const int capb = ctx->sm_count * 12;
            if ((int64_t)bpf * tc > capb) bpf = (int)((capb + tc - 1) / tc);
            const int bmin = (int)((n + 8 * COL_CHUNK - 1) / (8 * COL_CHUNK));
            if (bpf < bmin) bpf = bmin;
            if (bpf < 1) bpf = 1;
            { ProfScope ps(ctx, KC_SELECT_COLLECT);
              sel_collect_kernel<<<dim3((unsigned)bpf, (unsigned)tc), 256, 0, ctx->stream>>>(s0, n, n_q, use_abs, sf, cand, cap, bhist); }
            B4D_LAUNCH_CHECK(ctx);
            { ProfScope ps(ctx, KC_SELECT_FINAL);
              sel_final_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(sf, cand, cap, bhist, n_q, q_dev, use_abs, o0, nv0, need); }
            B4D_LAUNCH_CHECK(ctx);
        }
        // frames the bracket did not resolve (or every frame when they are small) take the three-pass radix select;
        // resolved frames return at the first instruction of each kernel
        rc = radix_path(ctx, s0, tc, n, q_dev, n_q, use_abs, o0, nv0, fast ? need : nullptr, hist, st);
        if (rc) return rc;
    }
    return B4D_OK;
}

extern "C" int b4d_select_ranks(b4d_ctx* ctx, const float* stack, int64_t n_frames, int64_t frame_elems,
                                const double* quantiles_host, int n_q, int use_abs, float* out, int64_t* n_valid) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!stack || !out || !quantiles_host || n_frames < 1 || frame_elems < 1 || n_q < 1 || n_q > SEL_MAXQ)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_select_ranks: bad arguments (n_q must be 1..%d)", SEL_MAXQ);
    for (int i = 0; i < n_q; ++i)
        if (!(quantiles_host[i] >= 0.0 && quantiles_host[i] <= 1.0))
            return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_select_ranks: quantile %d outside [0, 1]", i);
    void* p = nullptr;
    int rc = b4d_scratch(ctx, SCR_MISC, 64 * sizeof(double), &p);
    if (rc) return rc;
    if ((rc = b4d_put_doubles(ctx, static_cast<double*>(p), quantiles_host, n_q))) return rc;
    return b4d_select_impl(ctx, stack, n_frames, frame_elems, static_cast<const double*>(p), n_q, use_abs, out, n_valid);
}
