// Per-pixel temporal moments over a (T, ny, nx) stack and the flat-field kernels.
//
// Temporal moments lift distribution_moments' definitions
// (metrics/statistics.py:75-81: mean, std ddof 0, m3/m2^1.5, m4/m2^2 - 3) along the time axis.
// The kernel streams the stack once: a thread owns four adjacent pixels, keeps their shifted
// power sums in registers (fp32 over 8 frames, folded into fp64), and adds them into the
// caller's (4, ny, nx) float64 accumulator -- which adds across GPUs because every rank uses the
// same shift map.
//
// Flat field: preprocessing/normalize.py:104-131.
#include "common.cuh"

namespace {

constexpr int TM_UNROLL = 8;

struct FF4 {
    float4 g, d;
    bool on;
};

__device__ __forceinline__ float4 ff4(float4 v, const FF4& f) {
    if (f.on) {
        v.x = (v.x - f.d.x) * f.g.x; v.y = (v.y - f.d.y) * f.g.y;
        v.z = (v.z - f.d.z) * f.g.z; v.w = (v.w - f.d.w) * f.g.w;
    }
    return v;
}

// grid.x covers pixel quads, grid.y splits the frame range (each slice atomically adds into sums)
__global__ void __launch_bounds__(256) temporal_accumulate_kernel(const float* __restrict__ stack, int64_t T,
                                                                  int64_t npix, int64_t frames_per_slice,
                                                                  const float* __restrict__ gain,
                                                                  const float* __restrict__ dark,
                                                                  const float* __restrict__ shift,
                                                                  double* __restrict__ sums) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // quad index
    const int64_t p = q * 4;
    if (p >= npix) return;
    const int64_t t0 = (int64_t)blockIdx.y * frames_per_slice;
    const int64_t t1 = min(T, t0 + frames_per_slice);
    if (t0 >= t1) return;

    FF4 f;
    f.on = gain != nullptr;
    f.g = f.on ? __ldg(reinterpret_cast<const float4*>(gain + p)) : make_float4(1.f, 1.f, 1.f, 1.f);
    f.d = (f.on && dark) ? __ldg(reinterpret_cast<const float4*>(dark + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 K = __ldg(reinterpret_cast<const float4*>(shift + p));

    double S[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) S[k][c] = 0.0;

    const float* base = stack + p;
    for (int64_t t = t0; t < t1; t += TM_UNROLL) {
        float4 v[TM_UNROLL];
#pragma unroll
        for (int u = 0; u < TM_UNROLL; ++u) {
            const int64_t tt = min(t + u, t1 - 1);
            v[u] = ldg_stream4(base + tt * npix);
        }
        float a[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[k][c] = 0.f;
#pragma unroll
        for (int u = 0; u < TM_UNROLL; ++u) {
            if (t + u < t1) {
                const float4 x = ff4(v[u], f);
                const float d[4] = {x.x - K.x, x.y - K.y, x.z - K.z, x.w - K.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float d2 = d[c] * d[c];
                    a[0][c] += d[c];
                    a[1][c] += d2;
                    a[2][c] = fmaf(d2, d[c], a[2][c]);
                    a[3][c] = fmaf(d2, d2, a[3][c]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) S[k][c] += (double)a[k][c];
    }

    if (gridDim.y == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double* o = sums + (size_t)k * npix + p;
            double2 lo = *reinterpret_cast<double2*>(o), hi = *reinterpret_cast<double2*>(o + 2);
            lo.x += S[k][0]; lo.y += S[k][1]; hi.x += S[k][2]; hi.y += S[k][3];
            *reinterpret_cast<double2*>(o) = lo;
            *reinterpret_cast<double2*>(o + 2) = hi;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(sums + (size_t)k * npix + p + c, S[k][c]);
    }
}

// scalar fallback for npix % 4 != 0 or unaligned pointers
__global__ void __launch_bounds__(256) temporal_accumulate_scalar_kernel(const float* __restrict__ stack, int64_t T,
                                                                         int64_t npix, const float* gain,
                                                                         const float* dark, const float* shift,
                                                                         double* sums) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float g = gain ? gain[p] : 1.f, dk = (gain && dark) ? dark[p] : 0.f, K = shift[p];
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    for (int64_t t = 0; t < T; ++t) {
        float x = stack[t * npix + p];
        if (gain) x = (x - dk) * g;
        const double d = (double)(x - K), d2 = d * d;
        s1 += d; s2 += d2; s3 += d2 * d; s4 += d2 * d2;
    }
    sums[p] += s1; sums[npix + p] += s2; sums[2 * npix + p] += s3; sums[3 * npix + p] += s4;
}

__global__ void __launch_bounds__(256) temporal_pilot_kernel(const float* __restrict__ stack, int64_t T, int64_t npix,
                                                             const float* gain, const float* dark,
                                                             float* __restrict__ shift) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float g = gain ? gain[p] : 1.f, dk = (gain && dark) ? dark[p] : 0.f;
    float s = 0.f;
    int n = 0;
    for (int64_t t = 0; t < T; ++t) {
        float x = stack[t * npix + p];
        if (gain) x = (x - dk) * g;
        if (isfinite(x)) { s += x; n++; }
    }
    shift[p] = n ? s / (float)n : 0.f;
}

__global__ void __launch_bounds__(256) temporal_finalize_kernel(const double* __restrict__ sums,
                                                                const float* __restrict__ shift, double n,
                                                                int64_t npix, double* __restrict__ maps) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const double a1 = sums[p] / n, a2 = sums[npix + p] / n, a3 = sums[2 * npix + p] / n, a4 = sums[3 * npix + p] / n;
    const double mean = (double)shift[p] + a1;
    double m2 = a2 - a1 * a1;
    const double m3 = a3 - 3.0 * a1 * a2 + 2.0 * a1 * a1 * a1;
    const double m4 = a4 - 4.0 * a1 * a3 + 6.0 * a1 * a1 * a2 - 3.0 * a1 * a1 * a1 * a1;
    if (m2 < 0) m2 = 0;
    const double sd = sqrt(m2);
    maps[p] = mean;
    maps[npix + p] = sd;
    maps[2 * npix + p] = sd * sd;
    maps[3 * npix + p] = m3 / (m2 * sd);          // m3 / m2^1.5  (NaN for constant pixels, as scipy)
    maps[4 * npix + p] = m4 / (m2 * m2) - 3.0;
}

// ---- flat field ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flat_field_kernel(const float* __restrict__ img, int64_t T, int64_t npix,
                                                         const float* __restrict__ flat,
                                                         const float* __restrict__ dark, float eps, float s,
                                                         int apply_scale, float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float dk = dark ? dark[p] : 0.f;
    const float den = flat[p] - dk;
    const bool bad = den <= eps;
    const float dsafe = bad ? 1.f : den;
    for (int64_t t = blockIdx.y; t < T; t += gridDim.y) {
        // same float32 operation order as the reference: (img - dark) / den_safe, then *= scale
        float v = __fdiv_rn(img[t * npix + p] - dk, dsafe);
        if (apply_scale) v = __fmul_rn(v, s);
        out[t * npix + p] = bad ? 0.f : v;
    }
}

// 3x3 median repair of the bad pixels (preprocessing/normalize.py:134-140): out[bad] = median_filter(out, size=3)[bad],
// scipy's default 'reflect' border (the edge sample repeated). The filter sees the UNREPAIRED frame, in which every bad
// pixel is 0 -- so a bad neighbour contributes 0 whether or not its own repair has already been written, and the repair
// can run in place.
__device__ __forceinline__ void cswap(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }

__global__ void __launch_bounds__(256) bad_pixel_repair_kernel(float* __restrict__ frames, int64_t T, int ny, int nx,
                                                               const float* __restrict__ flat, const float* __restrict__ dark,
                                                               float eps) {
    const int64_t npix = (int64_t)ny * nx;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    auto is_bad = [&](int64_t q) { return flat[q] - (dark ? dark[q] : 0.f) <= eps; };
    if (!is_bad(p)) return;
    const int y = (int)(p / nx), x = (int)(p % nx);
    int64_t nb[9];
    bool nbad[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = min(max(y + k / 3 - 1, 0), ny - 1), xx = min(max(x + k % 3 - 1, 0), nx - 1);
        nb[k] = (int64_t)yy * nx + xx;
        nbad[k] = is_bad(nb[k]);
    }
    for (int64_t t = blockIdx.y; t < T; t += gridDim.y) {
        float* f = frames + t * npix;
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = nbad[k] ? 0.f : f[nb[k]];
        // median of nine by the 19-exchange network
        cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]); cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
        cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]); cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
        cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]); cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
        cswap(v[4], v[2]);
        f[p] = v[4];
    }
}

__global__ void __launch_bounds__(256) flat_gain_kernel(const float* __restrict__ flat, const float* __restrict__ dark,
                                                        int64_t npix, float eps, float s, float* __restrict__ gain) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float den = flat[p] - (dark ? dark[p] : 0.f);
    gain[p] = (den <= eps) ? 0.f : s / den;
}

__global__ void __launch_bounds__(256) sub_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                  float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = a[p] - (b ? b[p] : 0.f);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int b4d_temporal_accumulate(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                                       const float* gain, const float* dark, const float* shift, double* sums) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!stack || !shift || !sums || n_frames < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_temporal_accumulate: bad arguments");
    if (dark && !gain) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_temporal_accumulate: dark given without gain");
    const int64_t npix = (int64_t)ny * nx;
    const bool vec = npix % 4 == 0 && aligned16(stack) && aligned16(shift) && aligned16(sums) &&
                     (!gain || aligned16(gain)) && (!dark || aligned16(dark));
    ProfScope ps(ctx, KC_TEMPORAL);
    if (vec) {
        const int64_t quads = npix / 4;
        const unsigned gx = (unsigned)((quads + 255) / 256);
        // enough CTAs to fill the machine a few times over: split the frame range when the image is small
        int64_t slices = 1;
        const int64_t want = (int64_t)ctx->sm_count * 8;
        if (gx < want) slices = (want + gx - 1) / gx;
        if (slices > (n_frames + 63) / 64) slices = (n_frames + 63) / 64;
        if (slices < 1) slices = 1;
        int64_t fps = (n_frames + slices - 1) / slices;
        fps = (fps + TM_UNROLL - 1) / TM_UNROLL * TM_UNROLL;
        slices = (n_frames + fps - 1) / fps;
        dim3 grid(gx, (unsigned)slices);
        temporal_accumulate_kernel<<<grid, 256, 0, ctx->stream>>>(stack, n_frames, npix, fps, gain, dark, shift, sums);
    } else {
        temporal_accumulate_scalar_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(
            stack, n_frames, npix, gain, dark, shift, sums);
    }
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_temporal_pilot(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                                  const float* gain, const float* dark, float* shift) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!stack || !shift || n_frames < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_temporal_pilot: bad arguments");
    const int64_t npix = (int64_t)ny * nx;
    temporal_pilot_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(stack, n_frames, npix, gain, dark, shift);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_temporal_finalize(b4d_ctx* ctx, const double* sums, const float* shift, int64_t n_total, int ny,
                                     int nx, double* maps) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!sums || !shift || !maps || n_total < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_temporal_finalize: bad arguments");
    const int64_t npix = (int64_t)ny * nx;
    temporal_finalize_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(sums, shift, (double)n_total, npix, maps);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_flat_field(b4d_ctx* ctx, const float* images, int64_t n_frames, int ny, int nx, const float* flat,
                              const float* dark, float eps, float scale_value, int apply_scale, float* out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!images || !flat || !out || n_frames < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_flat_field: bad arguments");
    const int64_t npix = (int64_t)ny * nx;
    const unsigned gx = (unsigned)((npix + 255) / 256);
    unsigned gy = (unsigned)(n_frames < 64 ? n_frames : 64);
    ProfScope ps(ctx, KC_FLATFIELD);
    flat_field_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(images, n_frames, npix, flat, dark, eps, scale_value,
                                                             apply_scale, out);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_bad_pixel_repair(b4d_ctx* ctx, float* frames, int64_t n_frames, int ny, int nx, const float* flat,
                                    const float* dark, float eps) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!frames || !flat || n_frames < 1 || ny < 1 || nx < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_bad_pixel_repair: bad arguments");
    const int64_t npix = (int64_t)ny * nx;
    const unsigned gy = (unsigned)(n_frames < 64 ? n_frames : 64);
    ProfScope ps(ctx, KC_FLATFIELD);
    bad_pixel_repair_kernel<<<dim3((unsigned)((npix + 255) / 256), gy), 256, 0, ctx->stream>>>(frames, n_frames, ny, nx, flat, dark, eps);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_flat_gain(b4d_ctx* ctx, const float* flat, const float* dark, int ny, int nx, float eps,
                             float scale_value, float* gain) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!flat || !gain || ny < 1 || nx < 1) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_flat_gain: bad arguments");
    const int64_t npix = (int64_t)ny * nx;
    flat_gain_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(flat, dark, npix, eps, scale_value, gain);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

extern "C" int b4d_sub(b4d_ctx* ctx, const float* a, const float* b, int64_t n, float* out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!a || !out || n < 1) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_sub: bad arguments");
    sub_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, b, n, out);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}
