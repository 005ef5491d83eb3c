// Per-frame single-pass reductions: distribution moments, Sobel (Tenengrad) energies, Laplacian
// variance, zero / saturation counts -- one streaming read of each frame (HBM-bound).
//
// Replaces, per frame: distribution_moments (metrics/statistics.py:60-99), tenengrad
// (metrics/sharpness.py:440-465), laplacian_variance (metrics/sharpness.py:510-525) and the
// nanmean/nanstd of amplitude (metrics/speckles.py:636-645) of the reference.
//
// Numerics: every quantity is accumulated on d = x - K, K = a per-frame pilot mean (strided
// sample). The stencils are shift invariant and the moments are re-centred in the finalize
// kernel, so the fp32 work happens at the scale of the frame's standard deviation, not of its
// mean. Per-thread fp32 partial sums cover at most 32 pixels before they are folded into fp64.
#include "common.cuh"

namespace {

constexpr int FR_WARPS = 8;        // warps per CTA, each owns one (strip, band) item
constexpr int FR_STRIP = 128;      // columns per warp: 32 lanes x float4
constexpr int FR_BAND = 64;        // rows per item
constexpr int FR_NACC = 12;        // doubles per partial
constexpr int PILOT_SAMPLES = 2048;

struct FrArgs {
    const float* stack;
    const float* gain;   // nullable
    const float* dark;   // nullable
    const float* pilot;  // per-frame K
    double* partials;    // (T, blocks_per_frame, FR_NACC)
    int ny, nx;
    int nstrips, nitems;
    float sat, zeps;
    int has_sat;
};

__device__ __forceinline__ float ff_apply(float x, const float* gain, const float* dark, size_t p) {
    if (gain) {
        float dk = dark ? __ldg(dark + p) : 0.f;
        x = (x - dk) * __ldg(gain + p);
    }
    return x;
}

// ---- pilot: K[t] = mean of the finite pixels among PILOT_SAMPLES strided samples ----------------
__global__ void __launch_bounds__(256) pilot_kernel(const float* __restrict__ stack, const float* gain,
                                                    const float* dark, int64_t npix, float* __restrict__ pilot) {
    const int64_t t = blockIdx.x;
    const float* f = stack + t * npix;
    const int ns = (int)(npix < PILOT_SAMPLES ? npix : PILOT_SAMPLES);
    double s = 0.0;
    int c = 0;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        size_t p = (size_t)(((__int128)i * npix) / ns);
        float x = ff_apply(__ldg(f + p), gain, dark, p);
        if (isfinite(x)) { s += (double)x; c++; }
    }
    __shared__ double ss[8];
    __shared__ int sc[8];
    s = warp_sum(s);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; int n = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += ss[w]; n += sc[w]; }
        pilot[t] = n > 0 ? (float)(a / n) : 0.f;
    }
}

// One image row as seen by a lane: its 4 columns plus one halo column on each side, already
// shifted by K (d = x - K).
struct RowWin {
    float c[6];
    bool clean;   // all four pixels exist, are finite and are neither zero- nor saturation-candidates
};

struct Acc {
    float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, gx2 = 0.f, gy2 = 0.f, lap = 0.f, lap2 = 0.f;
    int nfin = 0, nzero = 0, nsat = 0, nnan = 0;
};

// Raw loads of one row for a lane: its four pixels (flat-field applied) and, for the two edge lanes of a
// strip, the neighbouring strip's pixel. Nothing here depends on another lane, so several rows can be in
// flight before the first shuffle.
struct RawRow {
    float x[4];
    float hl, hr;
};

template <bool VEC>
__device__ __forceinline__ RawRow fetch_row(const FrArgs& a, const float* frame, int r, int j0, int lane) {
    const int rr = min(max(r, 0), a.ny - 1);          // "reflect" = duplicate the edge sample
    const size_t rowoff = (size_t)rr * a.nx;
    const float* row = frame + rowoff;
    RawRow w;
    if (VEC) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 < a.nx) {
            v = __ldcs(reinterpret_cast<const float4*>(row + j0));
            if (a.gain) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(a.gain + rowoff + j0));
                const float4 dk = a.dark ? __ldg(reinterpret_cast<const float4*>(a.dark + rowoff + j0))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                v.x = (v.x - dk.x) * g.x; v.y = (v.y - dk.y) * g.y;
                v.z = (v.z - dk.z) * g.z; v.w = (v.w - dk.w) * g.w;
            }
        }
        w.x[0] = v.x; w.x[1] = v.y; w.x[2] = v.z; w.x[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int jc = min(j0 + k, a.nx - 1);   // clamped: the right reflection comes for free
            w.x[k] = ff_apply(__ldg(row + jc), a.gain, a.dark, rowoff + jc);
        }
    }
    w.hl = 0.f;
    w.hr = 0.f;
    if (lane == 0 && j0 > 0 && j0 < a.nx) w.hl = ff_apply(__ldg(row + j0 - 1), a.gain, a.dark, rowoff + j0 - 1);
    if (lane == 31 && j0 + 4 < a.nx) w.hr = ff_apply(__ldg(row + j0 + 4), a.gain, a.dark, rowoff + j0 + 4);
    return w;
}

// Shift by K, exchange the halo columns through shuffles, accumulate the pointwise statistics.
__device__ __forceinline__ RowWin finish_row(const FrArgs& a, const RawRow& raw, int j0, int lane, float K,
                                             bool in_band, Acc& acc) {
    RowWin w;
#pragma unroll
    for (int k = 0; k < 4; ++k) w.c[k + 1] = raw.x[k] - K;
    float left = __shfl_up_sync(0xffffffffu, w.c[4], 1);
    float right = __shfl_down_sync(0xffffffffu, w.c[1], 1);
    if (lane == 0) left = (j0 == 0) ? w.c[1] : raw.hl - K;
    // first column past the frame reflects onto the last valid one
    const int last = a.nx - 1 - j0;                   // index of the last valid pixel in this lane, if 0..3
    if (j0 + 4 >= a.nx) right = w.c[1 + min(max(last, 0), 3)];
    else if (lane == 31) right = raw.hr - K;
    if (last >= 0 && last < 3) {                      // ragged last lane (nx % 4 != 0): reflect inside the lane
#pragma unroll
        for (int k = 1; k < 4; ++k) if (k > last) w.c[k + 1] = w.c[1 + last];
    }
    w.c[0] = left;
    w.c[5] = right;
    // Lane-level screening (a handful of instructions per four pixels) so that ordinary pixels skip every
    // per-pixel test: the |x| sum is non-finite iff some pixel is NaN/inf (or absurdly large), min|x| > zeps
    // rules out zero-candidates, max x < sat rules out saturation-candidates.
    const float a0 = fabsf(raw.x[0]), a1 = fabsf(raw.x[1]), a2 = fabsf(raw.x[2]), a3 = fabsf(raw.x[3]);
    bool clean = (j0 + 4 <= a.nx) && ((a0 + a1) + (a2 + a3) <= 3.402823466e38f) &&
                 (fminf(fminf(a0, a1), fminf(a2, a3)) > a.zeps);
    if (a.has_sat) clean = clean && (fmaxf(fmaxf(raw.x[0], raw.x[1]), fmaxf(raw.x[2], raw.x[3])) < a.sat);
    w.clean = clean;
    if (in_band) {
        if (clean) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d = w.c[k + 1];
                const float d2 = d * d;
                acc.s1 += d;
                acc.s2 += d2;
                acc.s3 = fmaf(d2, d, acc.s3);
                acc.s4 = fmaf(d2, d2, acc.s4);
            }
            acc.nfin += 4;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xv = raw.x[k];
                const bool valid = j0 + k < a.nx;
                const bool fin = valid && (fabsf(xv) <= 3.402823466e38f);   // false for NaN / inf
                if (fin) {
                    const float d = w.c[k + 1];
                    const float d2 = d * d;
                    acc.s1 += d;
                    acc.s2 += d2;
                    acc.s3 = fmaf(d2, d, acc.s3);
                    acc.s4 = fmaf(d2, d2, acc.s4);
                    acc.nfin++;
                    acc.nzero += (fabsf(xv) <= a.zeps) ? 1 : 0;
                    acc.nsat += (a.has_sat && xv >= a.sat) ? 1 : 0;
                } else if (valid && xv != xv) {
                    acc.nnan++;
                }
            }
        }
    }
    return w;
}

__device__ __forceinline__ void stencil_row(const RowWin& up, const RowWin& mid, const RowWin& dn, int j0,
                                            int nx, Acc& acc) {
    float s[6], dv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        s[k] = up.c[k] + 2.f * mid.c[k] + dn.c[k];
        dv[k] = dn.c[k] - up.c[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float ctr = mid.c[k + 1];
        // the reference averages over pixels whose own value is finite; non-finite neighbours
        // propagate into the sums exactly as they do through scipy.ndimage.
        const bool fin = mid.clean || ((j0 + k < nx) && (fabsf(ctr) <= 3.402823466e38f));
        if (fin) {
            const float gx = s[k + 2] - s[k];
            const float gy = fmaf(2.f, dv[k + 1], dv[k] + dv[k + 2]);
            const float lp = fmaf(-4.f, ctr, (up.c[k + 1] + dn.c[k + 1]) + (mid.c[k] + mid.c[k + 2]));
            acc.gx2 = fmaf(gx, gx, acc.gx2);
            acc.gy2 = fmaf(gy, gy, acc.gy2);
            acc.lap += lp;
            acc.lap2 = fmaf(lp, lp, acc.lap2);
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(FR_WARPS * 32, 2) frame_reduce_kernel(FrArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * FR_WARPS + warp;
    const int64_t t = blockIdx.y;
    const float* frame = a.stack + (size_t)t * a.ny * a.nx;
    const float K = __ldg(a.pilot + t);

    double d[FR_NACC];
#pragma unroll
    for (int i = 0; i < FR_NACC; ++i) d[i] = 0.0;

    if (item < a.nitems) {
        const int strip = item % a.nstrips, band = item / a.nstrips;
        const int j0 = strip * FR_STRIP + lane * 4;
        const int r0 = band * FR_BAND;
        const int r1 = min(r0 + FR_BAND, a.ny);
        Acc acc;
        RawRow raw0 = fetch_row<VEC>(a, frame, r0 - 1, j0, lane), raw1 = fetch_row<VEC>(a, frame, r0, j0, lane);
        RawRow raw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[q] = fetch_row<VEC>(a, frame, r0 + 1 + q, j0, lane);
        RowWin up = finish_row(a, raw0, j0, lane, K, false, acc);
        RowWin mid = finish_row(a, raw1, j0, lane, K, true, acc);
        for (int rb = r0; rb < r1; rb += 4) {
            // the four rows fetched one iteration ago are consumed while the next four are already in flight
            RawRow cur[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) cur[q] = raw[q];
            if (rb + 4 < r1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) raw[q] = fetch_row<VEC>(a, frame, rb + 5 + q, j0, lane);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const RowWin nxt = finish_row(a, cur[q], j0, lane, K, rb + 1 + q < r1, acc);
                if (rb + q < r1) stencil_row(up, mid, nxt, j0, a.nx, acc);
                up = mid;
                mid = nxt;
            }
            if (((rb - r0) & 4) != 0 || rb + 4 >= r1) {   // fold fp32 partials into fp64 every 8 rows
                d[0] += acc.nfin; d[1] += acc.s1; d[2] += acc.s2; d[3] += acc.s3; d[4] += acc.s4;
                d[5] += acc.nzero; d[6] += acc.nsat; d[7] += acc.gx2; d[8] += acc.gy2;
                d[9] += acc.lap; d[10] += acc.lap2; d[11] += acc.nnan;
                acc = Acc();
            }
        }
    }

    __shared__ double sm[FR_WARPS][FR_NACC];
#pragma unroll
    for (int i = 0; i < FR_NACC; ++i) {
        double v = warp_sum(d[i]);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < FR_NACC) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < FR_WARPS; ++w) v += sm[w][threadIdx.x];
        a.partials[((size_t)t * gridDim.x + blockIdx.x) * FR_NACC + threadIdx.x] = v;
    }
}

// ---- finalize: fixed-order sum of the per-CTA partials, re-centre the moments ------------------
__global__ void __launch_bounds__(128) frame_finalize_kernel(const double* __restrict__ partials, int nblocks,
                                                            const float* __restrict__ pilot, double npix,
                                                            double* __restrict__ out) {
    const int64_t t = blockIdx.x;
    __shared__ double sm[4][FR_NACC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v[FR_NACC];
#pragma unroll
    for (int i = 0; i < FR_NACC; ++i) v[i] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        const double* p = partials + ((size_t)t * nblocks + b) * FR_NACC;
#pragma unroll
        for (int i = 0; i < FR_NACC; ++i) v[i] += p[i];
    }
#pragma unroll
    for (int i = 0; i < FR_NACC; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) sm[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[FR_NACC];
        for (int i = 0; i < FR_NACC; ++i) s[i] = sm[0][i] + sm[1][i] + sm[2][i] + sm[3][i];
        double* o = out + t * B4D_FR_NCOLS;
        const double n = s[0];
        const double K = (double)pilot[t];
        double mean = nan(""), m2 = nan(""), m3 = nan(""), m4 = nan("");
        if (n > 0) {
            const double a1 = s[1] / n, a2 = s[2] / n, a3 = s[3] / n, a4 = s[4] / n;   // raw moments of d
            mean = K + a1;
            m2 = a2 - a1 * a1;
            m3 = a3 - 3.0 * a1 * a2 + 2.0 * a1 * a1 * a1;
            m4 = a4 - 4.0 * a1 * a3 + 6.0 * a1 * a1 * a2 - 3.0 * a1 * a1 * a1 * a1;
            if (m2 < 0) m2 = 0;
        }
        o[B4D_FR_COUNT] = n; o[B4D_FR_MEAN] = mean; o[B4D_FR_M2] = m2; o[B4D_FR_M3] = m3; o[B4D_FR_M4] = m4;
        o[B4D_FR_NZERO] = s[5]; o[B4D_FR_NSAT] = s[6];
        o[B4D_FR_SGX2] = s[7]; o[B4D_FR_SGY2] = s[8]; o[B4D_FR_SLAP] = s[9]; o[B4D_FR_SLAP2] = s[10];
        o[B4D_FR_NPIX] = npix;
        o[B4D_FR_NNAN] = s[11];
    }
}

// largest float <= v  /  smallest float >= v  (so that float compares equal the reference's double compares)
float float_at_most(double v) {
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return f;
}
float float_at_least(double v) {
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}

}  // namespace

int b4d_frame_reductions_nolock(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                                const float* dark, double sat_value, double zero_eps, double* out);

int b4d_frame_pilot_launch(b4d_ctx* ctx, const float* stack, int64_t T, int64_t npix, const float* gain,
                           const float* dark, float* pilot) {
    ProfScope ps(ctx, KC_PILOT);
    pilot_kernel<<<(unsigned)T, 256, 0, ctx->stream>>>(stack, gain, dark, npix, pilot);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_frame_reductions(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                                    const float* gain, const float* dark, double sat_value, double zero_eps,
                                    double* out) {
    if (!ctx) return B4D_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->lock);
    return b4d_frame_reductions_nolock(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, out);
}

int b4d_frame_reductions_nolock(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                                const float* dark, double sat_value, double zero_eps, double* out) {
    if (!stack || !out || n_frames < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_frame_reductions: bad arguments (T=%lld ny=%d nx=%d)",
                        (long long)n_frames, ny, nx);
    if (dark && !gain) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_frame_reductions: dark given without gain");
    FrArgs a;
    a.stack = stack; a.gain = gain; a.dark = dark;
    a.ny = ny; a.nx = nx;
    a.nstrips = (nx + FR_STRIP - 1) / FR_STRIP;
    const int nbands = (ny + FR_BAND - 1) / FR_BAND;
    a.nitems = a.nstrips * nbands;
    const int nblocks = (a.nitems + FR_WARPS - 1) / FR_WARPS;
    a.has_sat = !(sat_value != sat_value);
    a.sat = a.has_sat ? float_at_least(sat_value) : 0.f;
    a.zeps = float_at_most(zero_eps);
    const bool vec = (nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(stack) & 15) == 0) &&
                     (!gain || (reinterpret_cast<uintptr_t>(gain) & 15) == 0) &&
                     (!dark || (reinterpret_cast<uintptr_t>(dark) & 15) == 0);

    void* p = nullptr;
    const int64_t chunk_max = 32768;   // gridDim.y limit is 65535
    int rc = b4d_scratch(ctx, SCR_PILOT, sizeof(float) * (size_t)n_frames, &p);
    if (rc) return rc;
    float* pilot = static_cast<float*>(p);
    rc = b4d_scratch(ctx, SCR_REDUCE, sizeof(double) * FR_NACC * (size_t)nblocks * (size_t)n_frames, &p);
    if (rc) return rc;
    a.partials = static_cast<double*>(p);
    a.pilot = pilot;

    for (int64_t t0 = 0; t0 < n_frames; t0 += chunk_max) {
        const int64_t tc = (n_frames - t0 < chunk_max) ? n_frames - t0 : chunk_max;
        const float* s0 = stack + (size_t)t0 * ny * nx;
        rc = b4d_frame_pilot_launch(ctx, s0, tc, (int64_t)ny * nx, gain, dark, pilot + t0);
        if (rc) return rc;
        FrArgs b = a;
        b.stack = s0;
        b.pilot = pilot + t0;
        b.partials = a.partials + (size_t)t0 * nblocks * FR_NACC;
        dim3 grid((unsigned)nblocks, (unsigned)tc);
        {
            ProfScope ps(ctx, KC_FRAME_REDUCE);
            if (vec) frame_reduce_kernel<true><<<grid, FR_WARPS * 32, 0, ctx->stream>>>(b);
            else frame_reduce_kernel<false><<<grid, FR_WARPS * 32, 0, ctx->stream>>>(b);
        }
        B4D_LAUNCH_CHECK(ctx);
        ProfScope ps2(ctx, KC_SMALL);
        frame_finalize_kernel<<<(unsigned)tc, 128, 0, ctx->stream>>>(b.partials, nblocks, b.pilot,
                                                                     (double)ny * (double)nx,
                                                                     out + t0 * B4D_FR_NCOLS);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}
