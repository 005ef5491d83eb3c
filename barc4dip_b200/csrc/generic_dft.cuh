// Frames whose sides are not powers of two (the 227 / 228-pixel sub-tiles of the reference's 9 x 9 tiling executor,
// metrics/common.py:75-106, :278-378, or any frame up to 2048 pixels a side): 2-D DFTs by Bluestein's chirp-z
// algorithm on top of the power-of-two register FFT core (fft.cuh).
//
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w[k - n]),   w[n] = exp(-i pi n^2 / N)
//
// is a linear convolution, evaluated as a circular one of length M >= 2N - 1 (M = 512 ... 8192):
// FFT_M(x w) * B, inverse FFT_M, times w[k] / M, with B = FFT_M of the wrapped conj chirp (host, double precision,
// cached per N). One kernel performs the whole 1-D transform of a row: M/16 threads, two in-register FFTs, nothing
// but the row itself crosses HBM. A 2-D transform is rows, transpose, rows; inverse transforms conjugate on the way
// in and out. This is the correctness-first path of the tiles executor; the power-of-two kernels above stay the
// hot path. Included by spectral.cu inside its anonymous namespace.
#pragma once

struct GenPlan {
    int n = 0, M = 0;
    float2* w = nullptr;   // (n)  chirp
    float2* B = nullptr;   // (M)  spectrum of the wrapped conjugate chirp
};

struct GenCache {
    std::vector<GenPlan> plans;
};

constexpr int GEN_MAX = 4096;   // largest side: detector frames such as 2560 x 2160 (the reference is size-agnostic, signal/fft.py:236)
inline bool gen_size_ok(int n) { return n >= 1 && n <= GEN_MAX; }   // 1: the 1-D signals of signal/fft.py, corr.py as (1, n) frames
inline int gen_conv_len(int n) { return n <= 256 ? 512 : (n <= 512 ? 1024 : (n <= 1024 ? 2048 : (n <= 2048 ? 4096 : 8192))); }

// iterative radix-2 FFT in double precision (host; tables only)
inline void host_fft(std::vector<double>& re, std::vector<double>& im) {
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * 3.14159265358979323846 / (double)len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double wr = cos(ang * (double)k), wi = sin(ang * (double)k);
                const size_t a = i + k, b = i + k + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
    }
}

int gen_plan(b4d_ctx* ctx, GenCache*& cache, int n, const GenPlan** out) {
    if (!cache) cache = new GenCache();
    for (const GenPlan& p : cache->plans) if (p.n == n) { *out = &p; return B4D_OK; }
    GenPlan p;
    p.n = n; p.M = gen_conv_len(n);
    std::vector<float2> w(n);
    std::vector<double> br(p.M, 0.0), bi(p.M, 0.0);
    for (int k = 0; k < n; ++k) {
        const long long k2 = ((long long)k * k) % (2LL * n);          // exact argument reduction
        const double a = -3.14159265358979323846 * (double)k2 / (double)n;
        w[k] = make_float2((float)cos(a), (float)sin(a));
        br[k] = cos(a); bi[k] = -sin(a);                              // conj chirp at +k ...
        if (k) { br[p.M - k] = cos(a); bi[p.M - k] = -sin(a); }       // ... and at -k
    }
    host_fft(br, bi);
    std::vector<float2> B(p.M);
    for (int k = 0; k < p.M; ++k) B[k] = make_float2((float)br[k], (float)bi[k]);
    B4D_CUDA(ctx, cudaMalloc(&p.w, sizeof(float2) * n));
    B4D_CUDA(ctx, cudaMalloc(&p.B, sizeof(float2) * p.M));
    B4D_CUDA(ctx, cudaMemcpy(p.w, w.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    B4D_CUDA(ctx, cudaMemcpy(p.B, B.data(), sizeof(float2) * p.M, cudaMemcpyHostToDevice));
    cache->plans.reserve(64);
    cache->plans.push_back(p);
    *out = &cache->plans.back();
    return B4D_OK;
}

void gen_release(GenCache* cache) {
    if (!cache) return;
    for (GenPlan& p : cache->plans) { cudaFree(p.w); cudaFree(p.B); }
    delete cache;
}

struct BluArgs {
    const float2* in;        // (rows, n) complex rows, or
    const float* in_real;    // (rows, n) real rows, value - mean[row / rows_per_frame]
    const double* fr;        // frame-reduction table of the real input (mean column), nullable
    float2* out;             // (rows, n)
    const float2* tw;        // base-power twiddles of the length-M core
    const float2* w;
    const float2* B;
    int n;
    int64_t rows;
    int rows_per_frame;
    int inverse;             // conjugate on the way in and out: unnormalised inverse DFT
};

template <int M>
__global__ void __launch_bounds__(512) bluestein_rows_kernel(BluArgs a) {
    constexpr int T = M / 16, FPC = 512 / T, FS = padded_len(M) + 8;
    constexpr int GROUP = M >= 1024 ? 1 : 0;
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x, f = tid / T, j = tid % T;
    int64_t row = (int64_t)blockIdx.x * FPC + f;
    const bool valid = row < a.rows;
    if (!valid) row = a.rows - 1;                      // keeps the barriers whole; nothing is stored
    const int n = a.n;
    float mean = 0.f;
    if (a.in_real && a.fr) mean = (float)a.fr[(row / a.rows_per_frame) * B4D_FR_NCOLS + B4D_FR_MEAN];
    float2 x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int idx = j + m * T;
        float2 v = make_float2(0.f, 0.f);
        if (idx < n) {
            if (a.in_real) v.x = a.in_real[row * n + idx] - mean;
            else v = a.in[row * n + idx];
            if (a.inverse) v.y = -v.y;
            v = cmul(v, __ldg(a.w + idx));
        }
        x[m] = v;
    }
    float2* z = sm + f * FS;
    fft_regs<M, -1, 1, GROUP>(x, j, z, a.tw, f);
#pragma unroll
    for (int s = 0; s < 16; ++s) x[s] = cmul(x[s], __ldg(a.B + j + s * T));
    fft_regs<M, +1, 1, GROUP>(x, j, z, a.tw, f);
    const float inv_m = 1.f / (float)M;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int idx = j + s * T;
        if (idx < n && valid) {
            float2 v = cmul(x[s], __ldg(a.w + idx));
            v.x *= inv_m; v.y *= inv_m;
            if (a.inverse) v.y = -v.y;
            a.out[row * n + idx] = v;
        }
    }
}

// Sides in (2048, 4096]: M = 8192, one radix-2 step around the 4096-point register core (which tops out at three radix-16
// stages). One row per CTA; threads 0..255 carry the even samples, 256..511 the odd ones:
//   forward (decimation in time):      U[k] = E[k] + W^k O[k],  U[k + 4096] = E[k] - W^k O[k],   W = exp(-2 pi i / 8192)
//   inverse (decimation in frequency): y[2m] = IFFT_4096(Y[k] + Y[k + 4096]),  y[2m + 1] = IFFT_4096((Y[k] - Y[k + 4096]) conj(W^k))
// The 4096-point transform leaves element k = j + 256 s in slot s of thread j of EITHER half, so both halves hold the
// same k and swap their 16 values through shared memory.
__global__ void __launch_bounds__(512) bluestein_rows_8192_kernel(BluArgs a) {
    constexpr int H = 4096, T = H / 16, FS = padded_len(H) + 8;
    extern __shared__ float2 sm[];
    float2* X = sm + 2 * FS;                           // [2][H] exchange of the two halves
    const int tid = threadIdx.x, f = tid / T, j = tid % T;
    const int64_t row = blockIdx.x;
    const int n = a.n;
    float mean = 0.f;
    if (a.in_real && a.fr) mean = (float)a.fr[(row / a.rows_per_frame) * B4D_FR_NCOLS + B4D_FR_MEAN];
    float2 x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int idx = 2 * (j + m * T) + f;           // sample index inside the zero-padded 8192 sequence
        float2 v = make_float2(0.f, 0.f);
        if (idx < n) {
            if (a.in_real) v.x = a.in_real[row * n + idx] - mean;
            else v = a.in[row * n + idx];
            if (a.inverse) v.y = -v.y;
            v = cmul(v, __ldg(a.w + idx));
        }
        x[m] = v;
    }
    float2* z = sm + f * FS;
    fft_regs<H, -1, 1, 1>(x, j, z, a.tw, f);
    // W^k of this thread's sixteen k (both halves hold the same k)
    float2 wk[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        float sn, cs;
        sincospif((float)(j + s * T) * (1.f / 4096.f), &sn, &cs);
        wk[s] = make_float2(cs, -sn);
    }
    // forward combine, times the chirp spectrum B
#pragma unroll
    for (int s = 0; s < 16; ++s) X[f * H + j + s * T] = f ? cmul(x[s], wk[s]) : x[s];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int k = j + s * T;
        const float2 e = X[k], o = X[H + k];
        const float2 u = f ? csub(e, o) : cadd(e, o);
        x[s] = cmul(u, __ldg(a.B + f * H + k));
    }
    __syncthreads();
    // inverse split
#pragma unroll
    for (int s = 0; s < 16; ++s) X[f * H + j + s * T] = x[s];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int k = j + s * T;
        const float2 lo = X[k], hi = X[H + k];
        x[s] = f ? cmul(csub(lo, hi), cconj(wk[s])) : cadd(lo, hi);
    }
    fft_regs<H, +1, 1, 1>(x, j, z, a.tw, f);
    const float inv_m = 1.f / 8192.f;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int idx = 2 * (j + s * T) + f;
        if (idx < n) {
            float2 v = cmul(x[s], __ldg(a.w + idx));
            v.x *= inv_m; v.y *= inv_m;
            if (a.inverse) v.y = -v.y;
            a.out[row * n + idx] = v;
        }
    }
}

// (T, a, b) -> (T, b, a), complex
__global__ void __launch_bounds__(256) gen_transpose_kernel(const float2* __restrict__ in, float2* __restrict__ out, int na, int nb) {
    __shared__ float2 tile[32][33];
    const int64_t t = blockIdx.z;
    const float2* src = in + (size_t)t * na * nb;
    float2* dst = out + (size_t)t * na * nb;
    const int b0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int ia = a0 + r, ib = b0 + tx;
        if (ia < na && ib < nb) tile[r][tx] = src[(size_t)ia * nb + ib];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ib = b0 + r, ia = a0 + tx;
        if (ia < na && ib < nb) dst[(size_t)ib * na + ia] = tile[tx][r];
    }
}

// F (T, ny, nx) natural order, transform of (frame - mean). dc_mode: 0 add nx*ny*mean back to F[0,0], 1 set F[0,0] = 0.
// Writes, as asked: the shifted PSD map * scale, the shifted complex spectrum, |F|^2 in place (real part) for the
// autocorrelation, and per-CTA partial spectral sums (same six columns as the power-of-two column pass).
struct GenEpiArgs {
    float2* F;
    const double* fr;
    int ny, nx;
    int dc_mode;
    float scale;
    float* psd_out;          // nullable
    float2* cplx_out;        // nullable
    int sqmag_inplace;
    double* spec_partials;   // nullable: (T, gridDim.x, NSP)
};

__global__ void __launch_bounds__(256) gen_epilogue_kernel(GenEpiArgs a) {
    const int64_t t = blockIdx.y;
    const int ny = a.ny, nx = a.nx;
    const int64_t npix = (int64_t)ny * nx;
    float2* F = a.F + (size_t)t * npix;
    const double fxm = (double)(nx / 2) / (double)nx, fym = (double)(ny / 2) / (double)ny;
    const double fmax = fxm < fym ? fxm : fym, fmax2 = fmax * fmax;
    double s[NSP] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const int ky = (int)(i / nx), kx = (int)(i % nx);
        float2 v = F[i];
        if (i == 0) {
            if (a.dc_mode) v = make_float2(0.f, 0.f);
            else if (a.fr) v.x += (float)((double)npix * a.fr[t * B4D_FR_NCOLS + B4D_FR_MEAN]);
        }
        const float P = v.x * v.x + v.y * v.y;
        const int sy = (ky + ny / 2) % ny, sx = (kx + nx / 2) % nx;           // np.fft.fftshift
        if (a.psd_out) a.psd_out[(size_t)t * npix + (size_t)sy * nx + sx] = P * a.scale;
        if (a.cplx_out) a.cplx_out[(size_t)t * npix + (size_t)sy * nx + sx] = v;
        if (a.sqmag_inplace) F[i] = make_float2(P, 0.f);
        if (a.spec_partials && i != 0) {
            const int kys = ky < (ny + 1) / 2 ? ky : ky - ny, kxs = kx < (nx + 1) / 2 ? kx : kx - nx;   // np.fft.fftfreq
            const double fy = (double)kys / (double)ny, fx = (double)kxs / (double)nx;
            const double p = (double)(P * a.scale);
            s[4] += p;
            if (p > 0.0) s[5] += p * (double)logf(P * a.scale);
            if (fx * fx + fy * fy <= fmax2) { s[0] += p; s[1] += fx * fx * p; s[2] += fy * fy * p; s[3] += p * p; }
        }
    }
    if (a.spec_partials) {
        __shared__ double red[8][NSP];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < NSP; ++k) {
            const double v = warp_sum(s[k]);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < NSP) {
            double v = 0.0;
            for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
            a.spec_partials[((size_t)t * gridDim.x + blockIdx.x) * NSP + threadIdx.x] = v;
        }
    }
}

// C (T, ny, nx): unnormalised inverse transform of |F|^2. out = shifted Re(C) * scale_t, scale_t = norm_mult / C[0,0]
// (peak normalisation: the zero lag is the maximum of an autocorrelation) or plain_scale.
__global__ void __launch_bounds__(256) gen_autocorr_out_kernel(const float2* __restrict__ C, int ny, int nx, int use_norm,
                                                                double norm_mult, double plain_scale, float* __restrict__ out) {
    const int64_t t = blockIdx.y;
    const int64_t npix = (int64_t)ny * nx;
    const float2* c = C + (size_t)t * npix;
    const double r0 = (double)c[0].x;
    const float sc = (float)((use_norm && r0 > 0.0) ? norm_mult / r0 : plain_scale);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i % nx);
        const int sy = (y + ny / 2) % ny, sx = (x + nx / 2) % nx;
        out[(size_t)t * npix + (size_t)sy * nx + sx] = c[i].x * sc;
    }
}

// ifftshift of a complex spectrum (in) into natural order (out), or natural order -> same with a scale (shift = 0)
__global__ void __launch_bounds__(256) gen_unshift_kernel(const float2* __restrict__ in, float2* __restrict__ out, int ny, int nx,
                                                           int shift, float scale) {
    const int64_t t = blockIdx.y;
    const int64_t npix = (int64_t)ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i % nx);
        // np.fft.ifftshift: out[k] = in[(k + n//2) % n]
        const int sy = shift ? (y + ny / 2) % ny : y, sx = shift ? (x + nx / 2) % nx : x;
        float2 v = in[(size_t)t * npix + (size_t)sy * nx + sx];
        v.x *= scale; v.y *= scale;
        out[(size_t)t * npix + i] = v;
    }
}

// A <- A * conj(B), both (T, ny, nx) natural order spectra of mean-removed frames; DC handling as in the epilogue:
// dc_mode 0 adds nx*ny*mean back to both DC bins first (no mean removal asked), 1 clears the product's DC bin.
__global__ void __launch_bounds__(256) gen_cross_kernel(float2* __restrict__ A, const float2* __restrict__ Bs, const double* __restrict__ fra,
                                                         const double* __restrict__ frb, int ny, int nx, int dc_mode) {
    const int64_t t = blockIdx.y;
    const int64_t npix = (int64_t)ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        float2 a = A[(size_t)t * npix + i], b = Bs[(size_t)t * npix + i];
        if (i == 0) {
            if (dc_mode) a = make_float2(0.f, 0.f);
            else {
                a.x += (float)((double)npix * fra[t * B4D_FR_NCOLS + B4D_FR_MEAN]);
                b.x += (float)((double)npix * frb[t * B4D_FR_NCOLS + B4D_FR_MEAN]);
            }
        }
        A[(size_t)t * npix + i] = make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
    }
}

// out = shifted Re(C) * scale (xcorr2d: real output, signal/corr.py:240-251)
__global__ void __launch_bounds__(256) gen_real_out_kernel(const float2* __restrict__ C, int ny, int nx, float scale, float* __restrict__ out) {
    const int64_t t = blockIdx.y;
    const int64_t npix = (int64_t)ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i % nx);
        out[(size_t)t * npix + (size_t)((y + ny / 2) % ny) * nx + (x + nx / 2) % nx] = C[(size_t)t * npix + i].x * scale;
    }
}

// Phase correlation (signal/tracking.py:277-285) on natural-order spectra: A <- whiten(A * inv_s * Rc) with Rc the
// conjugate spectrum of the embedded template (shared by all frames) and inv_s = 1 / (std + eps) of the frame (the frame's
// mean was removed before the transform); DC: the exact residue of the mean removal.
__global__ void __launch_bounds__(256) gen_phase_product_kernel(float2* __restrict__ A, const float2* __restrict__ Rc,
                                                                 const double* __restrict__ fr, int64_t npix, float eps) {
    const int64_t t = blockIdx.y;
    const float inv_s = (float)(1.0 / (sqrt(fr[t * B4D_FR_NCOLS + B4D_FR_M2]) + (double)eps));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        float2 f = A[(size_t)t * npix + i];
        f.x *= inv_s; f.y *= inv_s;
        const float2 r = Rc[i];
        float2 g = make_float2(f.x * r.x - f.y * r.y, f.x * r.y + f.y * r.x);
        const float s2 = fmaf(g.x, g.x, g.y * g.y);
        const float mag = s2 > 0.f ? sqrtf(s2) : 0.f;
        const float inv = 1.f / (mag + eps);
        A[(size_t)t * npix + i] = make_float2(g.x * inv, g.y * inv);
    }
}

// A <- A * Rc (Rc: conjugate template spectra, one for all frames (r_stride 0) or one per frame)
__global__ void __launch_bounds__(256) gen_mul_kernel(float2* __restrict__ A, const float2* __restrict__ Rc, size_t r_stride, int64_t npix) {
    const int64_t t = blockIdx.y;
    const float2* r = Rc + (size_t)t * r_stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 f = A[(size_t)t * npix + i], q = r[i];
        A[(size_t)t * npix + i] = make_float2(f.x * q.x - f.y * q.y, f.x * q.y + f.y * q.x);
    }
}

// out = (raw - dark) * gain, the per-pixel affine map of the fused loaders, materialised (generic pipeline only)
__global__ void __launch_bounds__(256) gen_apply_gain_kernel(const float* __restrict__ raw, const float* __restrict__ gain,
                                                              const float* __restrict__ dark, int64_t npix, float* __restrict__ out) {
    const int64_t t = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x)
        out[(size_t)t * npix + i] = (raw[(size_t)t * npix + i] - (dark ? dark[i] : 0.f)) * gain[i];
}

// conj in place (the template spectrum is stored conjugated)
__global__ void __launch_bounds__(256) gen_conj_kernel(float2* __restrict__ A, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) A[i].y = -A[i].y;
}

// out = shifted |C| * scale (the tracker's magnitude map) or shifted Re(C) * scale
__global__ void __launch_bounds__(256) gen_shift_out_kernel(const float2* __restrict__ C, int ny, int nx, float scale, int magnitude,
                                                             float* __restrict__ out) {
    const int64_t t = blockIdx.y;
    const int64_t npix = (int64_t)ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i % nx);
        const float2 c = C[(size_t)t * npix + i];
        const float v = magnitude ? sqrtf(c.x * c.x + c.y * c.y) : c.x;
        out[(size_t)t * npix + (size_t)((y + ny / 2) % ny) * nx + (x + nx / 2) % nx] = v * scale;
    }
}

// first-occurrence argmax of a (T, n) float map, one CTA per frame
__global__ void __launch_bounds__(1024) gen_argmax_kernel(const float* __restrict__ map, int64_t n, unsigned* __restrict__ idx_out) {
    const int64_t t = blockIdx.x;
    const float* m = map + (size_t)t * n;
    ArgBest b = {-INFINITY, 0xffffffffu};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) best_update(b, m[i], (unsigned)i);
    __shared__ ArgBest sb[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgBest ob = {__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.idx, o)};
        best_update(b, ob.v, ob.idx);
    }
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) best_update(b, sb[w].v, sb[w].idx);
        idx_out[t] = b.idx == 0xffffffffu ? 0u : b.idx;
    }
}

template <int M>
int launch_bluestein(b4d_ctx* ctx, const BluArgs& a) {
    constexpr int T = M / 16, FPC = 512 / T;
    constexpr size_t smem = (size_t)FPC * (padded_len(M) + 8) * sizeof(float2);
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) { B4D_CUDA(ctx, cudaFuncSetAttribute(bluestein_rows_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[ctx->device] = true; }
    const int64_t blocks = (a.rows + FPC - 1) / FPC;
    if (blocks > 0x7fffffffLL) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "generic DFT: too many rows");
    ProfScope ps(ctx, KC_GENERIC);
    bluestein_rows_kernel<M><<<(unsigned)blocks, 512, smem, ctx->stream>>>(a);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int launch_bluestein_8192(b4d_ctx* ctx, const BluArgs& a) {
    constexpr size_t smem = ((size_t)2 * (padded_len(4096) + 8) + 2 * 4096) * sizeof(float2);
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) { B4D_CUDA(ctx, cudaFuncSetAttribute(bluestein_rows_8192_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[ctx->device] = true; }
    if (a.rows > 0x7fffffffLL) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "generic DFT: too many rows");
    ProfScope ps(ctx, KC_GENERIC);
    bluestein_rows_8192_kernel<<<(unsigned)a.rows, 512, smem, ctx->stream>>>(a);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int gen_rows(b4d_ctx* ctx, GenCache*& cache, BluArgs a) {
    const GenPlan* p = nullptr;
    int rc = gen_plan(ctx, cache, a.n, &p);
    if (rc) return rc;
    a.w = p->w; a.B = p->B;
    if ((rc = get_twiddle_bases(ctx, p->M > 4096 ? 4096 : p->M, &a.tw))) return rc;
    if (p->M == 8192) return launch_bluestein_8192(ctx, a);
    switch (p->M) {
        case 512: return launch_bluestein<512>(ctx, a);
        case 1024: return launch_bluestein<1024>(ctx, a);
        case 2048: return launch_bluestein<2048>(ctx, a);
        default: return launch_bluestein<4096>(ctx, a);
    }
}

int gen_transpose(b4d_ctx* ctx, const float2* in, float2* out, int64_t T, int na, int nb) {
    ProfScope ps(ctx, KC_GENERIC);
    gen_transpose_kernel<<<dim3((nb + 31) / 32, (na + 31) / 32, (unsigned)T), 256, 0, ctx->stream>>>(in, out, na, nb);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

// forward 2-D DFT of (frame - mean): A (T, ny, nx) natural order on return; Bf is a same-sized work buffer
int gen_forward(b4d_ctx* ctx, GenCache*& cache, const float* stack, const double* fr, int64_t T, int ny, int nx, float2* A, float2* Bf) {
    BluArgs r;
    memset(&r, 0, sizeof(r));
    r.in_real = stack; r.fr = fr; r.out = A; r.n = nx; r.rows = T * ny; r.rows_per_frame = ny;
    int rc = gen_rows(ctx, cache, r);
    if (rc) return rc;
    if ((rc = gen_transpose(ctx, A, Bf, T, ny, nx))) return rc;
    memset(&r, 0, sizeof(r));
    r.in = Bf; r.out = Bf; r.n = ny; r.rows = T * nx; r.rows_per_frame = nx;
    if ((rc = gen_rows(ctx, cache, r))) return rc;
    return gen_transpose(ctx, Bf, A, T, nx, ny);
}

// unnormalised inverse 2-D DFT, in place on A (T, ny, nx)
int gen_inverse(b4d_ctx* ctx, GenCache*& cache, int64_t T, int ny, int nx, float2* A, float2* Bf) {
    BluArgs r;
    memset(&r, 0, sizeof(r));
    r.in = A; r.out = A; r.n = nx; r.rows = T * ny; r.rows_per_frame = ny; r.inverse = 1;
    int rc = gen_rows(ctx, cache, r);
    if (rc) return rc;
    if ((rc = gen_transpose(ctx, A, Bf, T, ny, nx))) return rc;
    memset(&r, 0, sizeof(r));
    r.in = Bf; r.out = Bf; r.n = ny; r.rows = T * nx; r.rows_per_frame = nx; r.inverse = 1;
    if ((rc = gen_rows(ctx, cache, r))) return rc;
    return gen_transpose(ctx, Bf, A, T, nx, ny);
}
