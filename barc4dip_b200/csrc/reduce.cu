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

// =================================================================================================
// v2 kernel (nx % 4 == 0, 16-byte aligned rows): overlapping strips, packed f32x2 arithmetic, fused tail collection
// =================================================================================================
// A warp marches down a strip of 30 x 4 = 120 own columns; lanes 0 and 31 load the 4 columns on either side and
// only serve as the halo of lanes 1 and 30, so every horizontal neighbour comes from one shuffle and no lane
// issues extra loads. Pixel pairs (c1,c2) and (c3,c4) of a lane live in float2 registers: the column-direction
// parts of the Sobel / Laplacian stencils and all accumulations are FADD2 / FMUL2 / FFMA2.
// Tail collection (the 0.05 / 99.95 percentiles of amplitude(), metrics/speckles.py:647): every pixel <= thr_lo
// or >= thr_hi (thresholds from a 16 K sample, ~0.2 % of the frame each) is staged in shared memory and flushed
// to a per-frame candidate list, from which tails_final_kernel reads the exact order statistics.
constexpr int F2_WARPS = 8;
constexpr int F2_STRIP = 120;
constexpr int F2_BAND = 64;
constexpr int F2_CAP = 1024;      // staged tail candidates per CTA and tail
constexpr int F2_QCAP = 1024;     // queued pixel quads (lanes whose min / max crossed a threshold) per CTA

struct Fr2Args {
    const float* stack;
    const float* gain;
    const float* dark;
    const float* pilot;       // per frame K
    const float* thr;         // per frame (thr_lo, thr_hi); nullable = no tail collection
    double* partials;         // (T, blocks_per_frame, FR_NACC)
    float* cand;              // (T, 2, gcap)
    unsigned* cand_cnt;       // (T, 2), zeroed by the caller
    unsigned* eq_cnt;         // (T, 2) pixels equal to thr_lo / thr_hi (ties at a threshold are counted, not listed), zeroed
    int* flag;                // (T) set when a staging buffer or a candidate list overflowed
    unsigned gcap;
    int ny, nx;
    int nstrips, nitems;
    float sat, zeps;
    int has_sat;
};

struct Win2 {
    float2 q0, q1;            // own pixels (c1,c2), (c3,c4), shifted by K
    float c0, c5;             // halo columns
    bool clean;
};

struct Acc2 {
    float2 s1, s2, s3, s4, gx2, gy2, lap, lap2;
    int nfin, nzero, nsat, nnan;
    __device__ __forceinline__ void clear() {
        s1 = s2 = s3 = s4 = gx2 = gy2 = lap = lap2 = make_float2(0.f, 0.f);
        nfin = nzero = nsat = nnan = 0;
    }
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

template <bool HAS_GAIN>
__device__ __forceinline__ float4 fetch2(const Fr2Args& a, const float* frame, int r, int j0, bool ld_ok) {
    const int rr = min(max(r, 0), a.ny - 1);          // "reflect" = duplicate the edge sample
    const unsigned off = (unsigned)rr * (unsigned)a.nx + (unsigned)j0;   // a frame has fewer than 2^32 pixels
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ld_ok) {
        v = __ldcs(reinterpret_cast<const float4*>(frame + off));
        if (HAS_GAIN) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(a.gain + off));
            const float4 dk = a.dark ? __ldg(reinterpret_cast<const float4*>(a.dark + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v.x = (v.x - dk.x) * g.x; v.y = (v.y - dk.y) * g.y;
            v.z = (v.z - dk.z) * g.z; v.w = (v.w - dk.w) * g.w;
        }
    }
    return v;
}

template <bool TAILS>
__device__ __forceinline__ Win2 finish2(const Fr2Args& a, const float4 raw, float K, bool edge_l, bool edge_r, bool count,
                                        float thr_lo, float thr_hi, Acc2& acc, float4* s_queue, unsigned* s_qn, unsigned* s_eq) {
    Win2 w;
    w.q0 = __fadd2_rn(f2(raw.x, raw.y), f2(-K, -K));
    w.q1 = __fadd2_rn(f2(raw.z, raw.w), f2(-K, -K));
    const float left = __shfl_up_sync(0xffffffffu, w.q1.y, 1);
    const float right = __shfl_down_sync(0xffffffffu, w.q0.x, 1);
    w.c0 = edge_l ? w.q0.x : left;
    w.c5 = edge_r ? w.q1.y : right;
    // lane-level screening so that ordinary pixels skip every per-pixel test
    const float sumabs = (fabsf(raw.x) + fabsf(raw.y)) + (fabsf(raw.z) + fabsf(raw.w));
    const float minabs = fminf(fminf(fabsf(raw.x), fabsf(raw.y)), fminf(fabsf(raw.z), fabsf(raw.w)));
    const float mx = fmaxf(fmaxf(raw.x, raw.y), fmaxf(raw.z, raw.w));
    bool clean = (sumabs <= 3.402823466e38f) && (minabs > a.zeps);
    if (a.has_sat) clean = clean && (mx < a.sat);
    w.clean = clean;
    if (count) {
        if (clean) {
            const float2 da = __fmul2_rn(w.q0, w.q0), db = __fmul2_rn(w.q1, w.q1);
            acc.s1 = __fadd2_rn(acc.s1, __fadd2_rn(w.q0, w.q1));
            acc.s2 = __fadd2_rn(acc.s2, __fadd2_rn(da, db));
            acc.s3 = __ffma2_rn(da, w.q0, acc.s3);
            acc.s3 = __ffma2_rn(db, w.q1, acc.s3);
            acc.s4 = __ffma2_rn(da, da, acc.s4);
            acc.s4 = __ffma2_rn(db, db, acc.s4);
            acc.nfin += 4;
        } else {
            const float xs[4] = {raw.x, raw.y, raw.z, raw.w};
            const float ds[4] = {w.q0.x, w.q0.y, w.q1.x, w.q1.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xv = xs[k];
                if (fabsf(xv) <= 3.402823466e38f) {          // false for NaN / inf
                    const float d = ds[k], d2 = d * d;
                    acc.s1.x += d;
                    acc.s2.x += d2;
                    acc.s3.x = fmaf(d2, d, acc.s3.x);
                    acc.s4.x = fmaf(d2, d2, acc.s4.x);
                    acc.nfin++;
                    acc.nzero += (fabsf(xv) <= a.zeps) ? 1 : 0;
                    acc.nsat += (a.has_sat && xv >= a.sat) ? 1 : 0;
                } else if (xv != xv) {
                    acc.nnan++;
                }
            }
        }
        if (TAILS) {
            // rare (~1.6 % of the lanes): park the quad, the per-pixel tests run densely after the main loop. Pixels EQUAL to
            // a threshold are only counted: integer detector frames tie by the thousand at 0 (masked / dark regions) or
            // at the saturation value, and the order statistics inside a run of ties is the tie itself.
            const float mn = fminf(fminf(raw.x, raw.y), fminf(raw.z, raw.w));
            if (mn <= thr_lo || mx >= thr_hi) {
                if (mn == thr_lo || mx == thr_hi) {
                    const unsigned el = (raw.x == thr_lo) + (raw.y == thr_lo) + (raw.z == thr_lo) + (raw.w == thr_lo);
                    const unsigned eh = (raw.x == thr_hi) + (raw.y == thr_hi) + (raw.z == thr_hi) + (raw.w == thr_hi);
                    if (el) atomicAdd(s_eq, el);
                    if (eh) atomicAdd(s_eq + 1, eh);
                }
                if (mn < thr_lo || mx > thr_hi) {
                    const unsigned p = atomicAdd(s_qn, 1u);
                    if (p < F2_QCAP) s_queue[p] = raw;
                }
            }
        }
    }
    return w;
}

__device__ __forceinline__ void stencil2(const Win2& up, const Win2& mid, const Win2& dn, Acc2& acc) {
    const float2 V0 = __fadd2_rn(up.q0, dn.q0), V1 = __fadd2_rn(up.q1, dn.q1);
    const float2 S0 = __ffma2_rn(f2(2.f, 2.f), mid.q0, V0), S1 = __ffma2_rn(f2(2.f, 2.f), mid.q1, V1);
    const float s0 = fmaf(2.f, mid.c0, up.c0 + dn.c0), s5 = fmaf(2.f, mid.c5, up.c5 + dn.c5);
    const float2 D0 = __fadd2_rn(dn.q0, f2(-up.q0.x, -up.q0.y)), D1 = __fadd2_rn(dn.q1, f2(-up.q1.x, -up.q1.y));
    const float d0 = dn.c0 - up.c0, d5 = dn.c5 - up.c5;
    const float2 GXa = f2(S0.y - s0, S1.x - S0.x), GXb = f2(S1.y - S0.y, s5 - S1.x);
    const float2 GYa = f2(fmaf(2.f, D0.x, d0 + D0.y), fmaf(2.f, D0.y, D0.x + D1.x));
    const float2 GYb = f2(fmaf(2.f, D1.x, D0.y + D1.y), fmaf(2.f, D1.y, D1.x + d5));
    const float2 Ha = f2(mid.c0 + mid.q0.y, mid.q0.x + mid.q1.x), Hb = f2(mid.q0.y + mid.q1.y, mid.q1.x + mid.c5);
    const float2 LPa = __ffma2_rn(f2(-4.f, -4.f), mid.q0, __fadd2_rn(V0, Ha));
    const float2 LPb = __ffma2_rn(f2(-4.f, -4.f), mid.q1, __fadd2_rn(V1, Hb));
    if (mid.clean) {
        acc.gx2 = __ffma2_rn(GXa, GXa, acc.gx2);
        acc.gx2 = __ffma2_rn(GXb, GXb, acc.gx2);
        acc.gy2 = __ffma2_rn(GYa, GYa, acc.gy2);
        acc.gy2 = __ffma2_rn(GYb, GYb, acc.gy2);
        acc.lap = __fadd2_rn(acc.lap, __fadd2_rn(LPa, LPb));
        acc.lap2 = __ffma2_rn(LPa, LPa, acc.lap2);
        acc.lap2 = __ffma2_rn(LPb, LPb, acc.lap2);
    } else {
        // the reference averages over pixels whose own value is finite; non-finite neighbours propagate into the
        // sums exactly as they do through scipy.ndimage.
        const float ctr[4] = {mid.q0.x, mid.q0.y, mid.q1.x, mid.q1.y};
        const float gx[4] = {GXa.x, GXa.y, GXb.x, GXb.y}, gy[4] = {GYa.x, GYa.y, GYb.x, GYb.y};
        const float lp[4] = {LPa.x, LPa.y, LPb.x, LPb.y};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (fabsf(ctr[k]) <= 3.402823466e38f) {
                acc.gx2.x = fmaf(gx[k], gx[k], acc.gx2.x);
                acc.gy2.x = fmaf(gy[k], gy[k], acc.gy2.x);
                acc.lap.x += lp[k];
                acc.lap2.x = fmaf(lp[k], lp[k], acc.lap2.x);
            }
        }
    }
}

template <bool HAS_GAIN, bool TAILS>
__global__ void __launch_bounds__(F2_WARPS * 32, 2) frame_reduce2_kernel(Fr2Args a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * F2_WARPS + warp;
    const int64_t t = blockIdx.y;
    const float* frame = a.stack + (size_t)t * a.ny * a.nx;
    const float K = __ldg(a.pilot + t);
    float thr_lo = -INFINITY, thr_hi = INFINITY;
    __shared__ float s_buf[TAILS ? 2 * F2_CAP : 2];
    __shared__ float4 s_queue[TAILS ? F2_QCAP : 1];
    __shared__ unsigned s_cnt[2];
    __shared__ unsigned s_eq[2];
    __shared__ unsigned s_base[2];
    __shared__ unsigned s_qn;
    if (TAILS) {
        thr_lo = __ldg(a.thr + 2 * t);
        thr_hi = __ldg(a.thr + 2 * t + 1);
        if (threadIdx.x < 2) { s_cnt[threadIdx.x] = 0u; s_eq[threadIdx.x] = 0u; }
        if (threadIdx.x == 2) s_qn = 0u;
        __syncthreads();
    }

    double d[FR_NACC];
#pragma unroll
    for (int i = 0; i < FR_NACC; ++i) d[i] = 0.0;

    if (item < a.nitems) {
        const int strip = item % a.nstrips, band = item / a.nstrips;
        const int j0 = strip * F2_STRIP + (lane - 1) * 4;
        const bool ld_ok = j0 >= 0 && j0 < a.nx;
        const bool own = lane >= 1 && lane <= 30 && j0 < a.nx;
        const bool edge_l = j0 == 0, edge_r = j0 + 4 == a.nx;
        const int r0 = band * F2_BAND;
        const int r1 = min(r0 + F2_BAND, a.ny);
        Acc2 acc;
        acc.clear();
        float4 raw[4];
        const float4 rawm = fetch2<HAS_GAIN>(a, frame, r0 - 1, j0, ld_ok), raw0 = fetch2<HAS_GAIN>(a, frame, r0, j0, ld_ok);
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[q] = fetch2<HAS_GAIN>(a, frame, r0 + 1 + q, j0, ld_ok);
        Win2 up = finish2<TAILS>(a, rawm, K, edge_l, edge_r, false, thr_lo, thr_hi, acc, s_queue, &s_qn, s_eq);
        Win2 mid = finish2<TAILS>(a, raw0, K, edge_l, edge_r, own, thr_lo, thr_hi, acc, s_queue, &s_qn, s_eq);
        for (int rb = r0; rb < r1; rb += 4) {
            // the four rows fetched one iteration ago are consumed while the next four are already in flight
            float4 cur[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) cur[q] = raw[q];
            if (rb + 4 < r1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) raw[q] = fetch2<HAS_GAIN>(a, frame, rb + 5 + q, j0, ld_ok);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const Win2 nxt = finish2<TAILS>(a, cur[q], K, edge_l, edge_r, own && (rb + 1 + q < r1), thr_lo, thr_hi, acc, s_queue, &s_qn, s_eq);
                if (own && rb + q < r1) stencil2(up, mid, nxt, acc);
                up = mid;
                mid = nxt;
            }
            if (((rb - r0) & 12) == 12 || rb + 4 >= r1) {   // fold fp32 partials into fp64 every 16 rows (32 values each)
                d[0] += acc.nfin; d[1] += (double)acc.s1.x + (double)acc.s1.y; d[2] += (double)acc.s2.x + (double)acc.s2.y;
                d[3] += (double)acc.s3.x + (double)acc.s3.y; d[4] += (double)acc.s4.x + (double)acc.s4.y;
                d[5] += acc.nzero; d[6] += acc.nsat;
                d[7] += (double)acc.gx2.x + (double)acc.gx2.y; d[8] += (double)acc.gy2.x + (double)acc.gy2.y;
                d[9] += (double)acc.lap.x + (double)acc.lap.y; d[10] += (double)acc.lap2.x + (double)acc.lap2.y;
                d[11] += acc.nnan;
                acc.clear();
            }
        }
    }

    __shared__ double sm[F2_WARPS][FR_NACC];
#pragma unroll
    for (int i = 0; i < FR_NACC; ++i) {
        double v = warp_sum(d[i]);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < FR_NACC) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < F2_WARPS; ++w) v += sm[w][threadIdx.x];
        a.partials[((size_t)t * gridDim.x + blockIdx.x) * FR_NACC + threadIdx.x] = v;
    }
    if (TAILS) {
        // per-pixel tests of the parked quads (the barrier above made the queue visible)
        const unsigned nq = min(s_qn, (unsigned)F2_QCAP);
        for (unsigned i = threadIdx.x; i < nq; i += blockDim.x) {
            const float4 v = s_queue[i];
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (xs[k] < thr_lo) { const unsigned p = atomicAdd(s_cnt, 1u); if (p < F2_CAP) s_buf[p] = xs[k]; }
                if (xs[k] > thr_hi) { const unsigned p = atomicAdd(s_cnt + 1, 1u); if (p < F2_CAP) s_buf[F2_CAP + p] = xs[k]; }
            }
        }
        __syncthreads();
        // flush the staged candidates: one reservation per CTA and tail
        if (threadIdx.x < 2) {
            if (s_qn > F2_QCAP) a.flag[t] = 1;
            if (s_eq[threadIdx.x]) atomicAdd(a.eq_cnt + 2 * t + threadIdx.x, s_eq[threadIdx.x]);
            const unsigned n = s_cnt[threadIdx.x];
            unsigned base = 0;
            if (n > F2_CAP) a.flag[t] = 1;
            else if (n) {
                base = atomicAdd(a.cand_cnt + 2 * t + threadIdx.x, n);
                if (base + n > a.gcap) a.flag[t] = 1;
            }
            s_base[threadIdx.x] = base;
        }
        __syncthreads();
#pragma unroll
        for (int tail = 0; tail < 2; ++tail) {
            const unsigned n = min(s_cnt[tail], (unsigned)F2_CAP), base = s_base[tail];
            float* dst = a.cand + ((size_t)t * 2 + tail) * a.gcap;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
                if (base + i < a.gcap) dst[base + i] = s_buf[tail * F2_CAP + i];
        }
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

#include "reduce.cuh"

unsigned b4d_tails_gcap();
int b4d_tails_probe_launch(b4d_ctx* ctx, const float* stack, int64_t T, int64_t npix, const float* gain, const float* dark,
                           double q_lo, double q_hi, float* thr);
int b4d_tails_final_launch(b4d_ctx* ctx, const float* cand, const unsigned* cnt, const unsigned* eq, const float* thr, const int* flag,
                           const double* fr, int64_t T, double q_lo, double q_hi, float* out, int64_t* nvalid_out);

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
    B4dCall g(ctx);
    return b4d_frame_reductions_nolock(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, out);
}

int b4d_frame_reductions_nolock(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                                const float* dark, double sat_value, double zero_eps, double* out) {
    return b4d_frame_reductions_ex(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, out, nullptr, nullptr);
}

extern "C" int b4d_frame_reductions_tails(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                                          const float* gain, const float* dark, double sat_value, double zero_eps,
                                          double q_lo, double q_hi, double* out, float* quant_out, int64_t* nvalid_out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!quant_out || !nvalid_out || !(q_lo >= 0.0 && q_lo < q_hi && q_hi <= 1.0))
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_frame_reductions_tails: need 0 <= q_lo < q_hi <= 1 and both outputs");
    FrTails tl = {q_lo, q_hi, quant_out, nvalid_out};
    return b4d_frame_reductions_ex(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, out, &tl, nullptr);
}

// The reduction pass in three steps so that the fused stack pipeline can run its streaming kernel a few frames at a
// time (the frames it has just read are then still in L2 when the forward row pass asks for them):
//   begin: scratch, pilot means and tail thresholds of every frame (strided samples only);
//   range: the streaming pass + finalize of frames [t0, t0 + tc) (tc <= 32768);
//   end:   exact tail order statistics of every frame.
int b4d_fr_begin(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain, const float* dark,
                 double sat_value, double zero_eps, double* out, const FrTails* tails, float* pilot_out, FrPlan* pl) {
    if (!stack || !out || n_frames < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_frame_reductions: bad arguments (T=%lld ny=%d nx=%d)",
                        (long long)n_frames, ny, nx);
    if (dark && !gain) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_frame_reductions: dark given without gain");
    pl->stack = stack; pl->gain = gain; pl->dark = dark; pl->n_frames = n_frames; pl->ny = ny; pl->nx = nx; pl->out = out;
    pl->has_tails = tails != nullptr;
    if (tails) pl->tails = *tails;
    pl->has_sat = !(sat_value != sat_value);
    pl->sat = pl->has_sat ? float_at_least(sat_value) : 0.f;
    pl->zeps = float_at_most(zero_eps);
    pl->vec = (nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(stack) & 15) == 0) &&
              (!gain || (reinterpret_cast<uintptr_t>(gain) & 15) == 0) &&
              (!dark || (reinterpret_cast<uintptr_t>(dark) & 15) == 0);
    const int64_t npix = (int64_t)ny * nx;
    // the fused tails need the v2 kernel; frames it does not cover are reported as unresolved
    pl->fuse_tails = tails && pl->vec && npix >= 65536;

    pl->nstrips = pl->vec ? (nx + F2_STRIP - 1) / F2_STRIP : (nx + FR_STRIP - 1) / FR_STRIP;
    const int nbands = (ny + FR_BAND - 1) / FR_BAND;
    pl->nitems = pl->nstrips * nbands;
    pl->nblocks = (pl->nitems + FR_WARPS - 1) / FR_WARPS;
    static_assert(F2_BAND == FR_BAND && F2_WARPS == FR_WARPS, "the two reduce kernels share the item decomposition");

    void* p = nullptr;
    int rc = b4d_scratch(ctx, SCR_PILOT, sizeof(float) * 3 * (size_t)n_frames, &p);
    if (rc) return rc;
    pl->pilot = pilot_out ? pilot_out : static_cast<float*>(p);
    pl->thr = static_cast<float*>(p) + n_frames;
    pl->gcap = b4d_tails_gcap();
    const size_t part_bytes = (sizeof(double) * FR_NACC * (size_t)pl->nblocks * (size_t)n_frames + 255) & ~size_t(255);
    const size_t cnt_bytes = pl->fuse_tails ? (((size_t)n_frames * 5 * sizeof(unsigned) + 255) & ~size_t(255)) : 0;
    const size_t cand_bytes = pl->fuse_tails ? (size_t)n_frames * 2 * pl->gcap * sizeof(float) : 0;
    rc = b4d_scratch(ctx, SCR_REDUCE, part_bytes + cnt_bytes + cand_bytes, &p);
    if (rc) return rc;
    pl->partials = static_cast<double*>(p);
    pl->cnt = reinterpret_cast<unsigned*>(static_cast<char*>(p) + part_bytes);
    pl->flag = reinterpret_cast<int*>(pl->cnt + 2 * n_frames);
    pl->eq = pl->cnt + 3 * n_frames;
    pl->cand = reinterpret_cast<float*>(static_cast<char*>(p) + part_bytes + cnt_bytes);
    if (pl->fuse_tails) B4D_CUDA(ctx, cudaMemsetAsync(pl->cnt, 0, cnt_bytes, ctx->stream));
    rc = b4d_frame_pilot_launch(ctx, stack, n_frames, npix, gain, dark, pl->pilot);
    if (rc) return rc;
    if (pl->fuse_tails && (rc = b4d_tails_probe_launch(ctx, stack, n_frames, npix, gain, dark, tails->q_lo, tails->q_hi, pl->thr))) return rc;
    return B4D_OK;
}

int b4d_fr_range(b4d_ctx* ctx, const FrPlan& pl, int64_t t0, int64_t tc) {
    if (t0 < 0 || tc < 1 || t0 + tc > pl.n_frames || tc > 32768)   // gridDim.y limit is 65535
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_frame_reductions: bad frame range");
    const int ny = pl.ny, nx = pl.nx;
    const float* s0 = pl.stack + (size_t)t0 * ny * nx;
    dim3 grid((unsigned)pl.nblocks, (unsigned)tc);
    double* part0 = pl.partials + (size_t)t0 * pl.nblocks * FR_NACC;
    if (pl.vec) {
        Fr2Args b;
        b.stack = s0; b.gain = pl.gain; b.dark = pl.dark; b.pilot = pl.pilot + t0; b.thr = pl.fuse_tails ? pl.thr + 2 * t0 : nullptr;
        b.partials = part0; b.cand = pl.cand + (size_t)t0 * 2 * pl.gcap; b.cand_cnt = pl.cnt + 2 * t0; b.eq_cnt = pl.eq + 2 * t0; b.flag = pl.flag + t0; b.gcap = pl.gcap;
        b.ny = ny; b.nx = nx; b.nstrips = pl.nstrips; b.nitems = pl.nitems; b.sat = pl.sat; b.zeps = pl.zeps; b.has_sat = pl.has_sat;
        ProfScope ps(ctx, KC_FRAME_REDUCE);
        if (pl.gain) {
            if (pl.fuse_tails) frame_reduce2_kernel<true, true><<<grid, F2_WARPS * 32, 0, ctx->stream>>>(b);
            else frame_reduce2_kernel<true, false><<<grid, F2_WARPS * 32, 0, ctx->stream>>>(b);
        } else {
            if (pl.fuse_tails) frame_reduce2_kernel<false, true><<<grid, F2_WARPS * 32, 0, ctx->stream>>>(b);
            else frame_reduce2_kernel<false, false><<<grid, F2_WARPS * 32, 0, ctx->stream>>>(b);
        }
    } else {
        FrArgs b;
        b.stack = s0; b.gain = pl.gain; b.dark = pl.dark; b.pilot = pl.pilot + t0; b.partials = part0;
        b.ny = ny; b.nx = nx; b.nstrips = pl.nstrips; b.nitems = pl.nitems; b.sat = pl.sat; b.zeps = pl.zeps; b.has_sat = pl.has_sat;
        ProfScope ps(ctx, KC_FRAME_REDUCE);
        frame_reduce_kernel<false><<<grid, FR_WARPS * 32, 0, ctx->stream>>>(b);
    }
    B4D_LAUNCH_CHECK(ctx);
    {
        ProfScope ps2(ctx, KC_SMALL);
        frame_finalize_kernel<<<(unsigned)tc, 128, 0, ctx->stream>>>(part0, pl.nblocks, pl.pilot + t0, (double)ny * (double)nx,
                                                                     pl.out + t0 * B4D_FR_NCOLS);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

int b4d_fr_end(b4d_ctx* ctx, const FrPlan& pl) {
    const int64_t T = pl.n_frames;
    if (pl.fuse_tails)
        return b4d_tails_final_launch(ctx, pl.cand, pl.cnt, pl.eq, pl.thr, pl.flag, pl.out, T, pl.tails.q_lo, pl.tails.q_hi,
                                      pl.tails.quant_out, pl.tails.nvalid_out);
    if (pl.has_tails) {
        B4D_CUDA(ctx, cudaMemsetAsync(pl.tails.nvalid_out, 0xff, sizeof(int64_t) * T, ctx->stream));   // -1: unresolved
        B4D_CUDA(ctx, cudaMemsetAsync(pl.tails.quant_out, 0xff, sizeof(float) * 4 * T, ctx->stream));   // NaN
    }
    return B4D_OK;
}

int b4d_frame_reductions_ex(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                            const float* dark, double sat_value, double zero_eps, double* out, const FrTails* tails,
                            float* pilot_out) {
    FrPlan pl;
    int rc = b4d_fr_begin(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, out, tails, pilot_out, &pl);
    if (rc) return rc;
    const int64_t chunk_max = 32768;
    for (int64_t t0 = 0; t0 < n_frames; t0 += chunk_max)
        if ((rc = b4d_fr_range(ctx, pl, t0, (n_frames - t0 < chunk_max) ? n_frames - t0 : chunk_max))) return rc;
    return b4d_fr_end(ctx, pl);
}
