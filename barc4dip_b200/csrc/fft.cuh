// In-CTA power-of-two complex FFTs for sm_100a: register radix-16/8/4/2 butterflies, Stockham
// autosort exchanges through padded shared memory.
//
// A length-N transform is owned by N/16 threads; every thread keeps 16 complex values in
// registers per stage.  Stage s with accumulated length LS and radix R (virtual thread v):
//     k = v % LS
//     x[m] = z[v + m*N/R] * exp(DIR * 2 pi i * m * k / (LS*R)),   m = 0..R-1
//     X    = DFT_R(x)
//     z[(v - k)*R + k + q*LS] = X[q]
// Reads are unit-stride in v for every stage; writes are made conflict-free by padding the
// array by one element every 16 (PAD).  Twiddles come from per-stage tables laid out [m-1][k]
// so that a warp reads consecutive entries.
#pragma once

#include <cuda_runtime.h>

namespace b4dfft {

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kCos1_16 = 0.92387953251128675613f;   // cos(pi/8)
constexpr float kSin1_16 = 0.38268343236508977173f;   // sin(pi/8)

__host__ __device__ constexpr int pad16(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by DIR * i  (DIR = -1: by -i, the forward W4; DIR = +1: by +i)
template <int DIR>
__device__ __forceinline__ float2 rot90(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// multiply by exp(DIR * i pi/4) and exp(DIR * 3 i pi/4)
template <int DIR>
__device__ __forceinline__ float2 rot45(float2 a) {
    return DIR < 0 ? make_float2((a.x + a.y) * kSqrtHalf, (a.y - a.x) * kSqrtHalf)
                   : make_float2((a.x - a.y) * kSqrtHalf, (a.y + a.x) * kSqrtHalf);
}
template <int DIR>
__device__ __forceinline__ float2 rot135(float2 a) {
    return DIR < 0 ? make_float2((a.y - a.x) * kSqrtHalf, -(a.x + a.y) * kSqrtHalf)
                   : make_float2(-(a.x + a.y) * kSqrtHalf, (a.x - a.y) * kSqrtHalf);
}
// multiply by (c, DIR*s)
template <int DIR>
__device__ __forceinline__ float2 rotcs(float2 a, float c, float s) {
    const float ss = DIR < 0 ? -s : s;
    return make_float2(fmaf(a.x, c, -a.y * ss), fmaf(a.x, ss, a.y * c));
}

template <int DIR>
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// in-place 4-point DFT of (a0,a1,a2,a3) -> outputs X0..X3 in the same slots
template <int DIR>
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), s3 = rot90<DIR>(csub(a1, a3));
    a0 = cadd(s0, s2);
    a2 = csub(s0, s2);
    a1 = cadd(s1, s3);
    a3 = csub(s1, s3);
}

// 8-point DFT: inner radix-2 over (x[a], x[a+4]), twiddle W8^(a*q1), outer radix-4 over a.
// Input x[0..7] natural order, output X[q] natural order in the same array.
template <int DIR>
__device__ __forceinline__ void bfly8(float2* x) {
    bfly2<DIR>(x[0], x[4]);
    bfly2<DIR>(x[1], x[5]);
    bfly2<DIR>(x[2], x[6]);
    bfly2<DIR>(x[3], x[7]);
    // t[a][q1]: q1 = 0 -> x[a], q1 = 1 -> x[a+4]; twiddle W8^a on the q1 = 1 branch
    x[5] = rot45<DIR>(x[5]);
    x[6] = rot90<DIR>(x[6]);
    x[7] = rot135<DIR>(x[7]);
    bfly4<DIR>(x[0], x[1], x[2], x[3]);   // q1 = 0: X[0 + 2 q2] in slots 0..3
    bfly4<DIR>(x[4], x[5], x[6], x[7]);   // q1 = 1: X[1 + 2 q2] in slots 4..7
    // reorder to natural: X[2 q2] = slot q2, X[1 + 2 q2] = slot 4 + q2
    const float2 t1 = x[1], t2 = x[2], t3 = x[3], t4 = x[4], t5 = x[5], t6 = x[6];
    x[1] = t4; x[2] = t1; x[3] = t5; x[4] = t2; x[5] = t6; x[6] = t3;
}

// 16-point DFT as 4 x 4: inner radix-4 over (x[a], x[a+4], x[a+8], x[a+12]) -> t[a][q1],
// twiddle W16^(a*q1), outer radix-4 over a -> X[q1 + 4 q2].
template <int DIR>
__device__ __forceinline__ void bfly16(float2* x) {
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4<DIR>(x[a], x[a + 4], x[a + 8], x[a + 12]);
    // now x[a + 4 q1] = t[a][q1]
    x[5] = rotcs<DIR>(x[5], kCos1_16, kSin1_16);      // a=1,q1=1: W16^1
    x[9] = rot45<DIR>(x[9]);                           // a=1,q1=2: W16^2
    x[13] = rotcs<DIR>(x[13], kSin1_16, kCos1_16);     // a=1,q1=3: W16^3 = (cos 3pi/8, sin 3pi/8) = (sin pi/8, cos pi/8)
    x[6] = rot45<DIR>(x[6]);                           // a=2,q1=1: W16^2
    x[10] = rot90<DIR>(x[10]);                         // a=2,q1=2: W16^4
    x[14] = rot135<DIR>(x[14]);                        // a=2,q1=3: W16^6
    x[7] = rotcs<DIR>(x[7], kSin1_16, kCos1_16);       // a=3,q1=1: W16^3
    x[11] = rot135<DIR>(x[11]);                        // a=3,q1=2: W16^6
    x[15] = rotcs<DIR>(x[15], -kCos1_16, -kSin1_16);   // a=3,q1=3: W16^9 = -W16^1
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) bfly4<DIR>(x[4 * q1], x[4 * q1 + 1], x[4 * q1 + 2], x[4 * q1 + 3]);
    // slot 4 q1 + q2 holds X[q1 + 4 q2]: transpose the 4x4 to natural order
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = r + 1; c < 4; ++c) {
            const float2 t = x[4 * r + c];
            x[4 * r + c] = x[4 * c + r];
            x[4 * c + r] = t;
        }
}

template <int R, int DIR>
__device__ __forceinline__ void bfly(float2* x) {
    if (R == 16) bfly16<DIR>(x);
    else if (R == 8) bfly8<DIR>(x);
    else if (R == 4) bfly4<DIR>(x[0], x[1], x[2], x[3]);
    else bfly2<DIR>(x[0], x[1]);
}

// ---- radix plan -------------------------------------------------------------------------------
template <int N> struct Plan;
template <> struct Plan<2048> { static constexpr int R1 = 16, R2 = 16, R3 = 8; };
template <> struct Plan<1024> { static constexpr int R1 = 16, R2 = 16, R3 = 4; };
template <> struct Plan<512>  { static constexpr int R1 = 16, R2 = 8,  R3 = 4; };
template <> struct Plan<256>  { static constexpr int R1 = 16, R2 = 16, R3 = 1; };
template <> struct Plan<128>  { static constexpr int R1 = 16, R2 = 8,  R3 = 1; };

// Twiddle tables for one N: stage 2 at tw, stage 3 at tw + (R2-1)*R1. Entry [(m-1)*LS + k] =
// exp(-2 pi i m k / (LS*R)) (forward sign; inverse transforms conjugate on the fly).
template <int N>
__host__ __device__ constexpr int twiddle_count() {
    return (Plan<N>::R2 - 1) * Plan<N>::R1 + (Plan<N>::R3 > 1 ? (Plan<N>::R3 - 1) * Plan<N>::R1 * Plan<N>::R2 : 0);
}

// One Stockham stage on registers. x holds (16/R) butterflies: x[b*R + m] is input m of
// butterfly b whose virtual thread is v = j + b*(N/16).
//   in_index(b, m)  = v + m*(N/R)
//   out_index(b, q) = (v - k)*R + k + q*LS,  k = v % LS
template <int N, int R, int LS>
struct StageIdx {
    static constexpr int NB = 16 / R;           // butterflies per thread
    __device__ static __forceinline__ int v(int j, int b) { return j + b * (N / 16); }
    __device__ static __forceinline__ int in(int j, int b, int m) { return v(j, b) + m * (N / R); }
    __device__ static __forceinline__ int out(int j, int b, int q) {
        const int vv = v(j, b), k = vv & (LS - 1);
        return (vv - k) * R + k + q * LS;
    }
};

template <int N, int R, int LS, int DIR>
__device__ __forceinline__ void stage_compute(float2* x, int j, const float2* __restrict__ tw) {
    using I = StageIdx<N, R, LS>;
#pragma unroll
    for (int b = 0; b < I::NB; ++b) {
        if (LS > 1) {
            const int k = I::v(j, b) & (LS - 1);
#pragma unroll
            for (int m = 1; m < R; ++m) {
                float2 w = __ldg(tw + (m - 1) * LS + k);
                if (DIR > 0) w.y = -w.y;
                x[b * R + m] = cmul(x[b * R + m], w);
            }
        }
        bfly<R, DIR>(x + b * R);
    }
}

// Full transform of data already sitting in registers for stage 1 (x[m] = z[j + m*N/16]).
// S is a functor giving the shared-memory slot of logical element i: float2& S(i).
// On return the transform is in shared memory in natural order (a __syncthreads() has been
// issued after the last store).
template <int N, int DIR, class Slot>
__device__ __forceinline__ void fft_from_regs(float2* x, int j, Slot S, const float2* __restrict__ tw) {
    using P = Plan<N>;
    {   // stage 1: LS = 1, radix R1 = 16
        using I = StageIdx<N, P::R1, 1>;
        stage_compute<N, P::R1, 1, DIR>(x, j, nullptr);
#pragma unroll
        for (int q = 0; q < 16; ++q) S(I::out(j, 0, q)) = x[q];
    }
    __syncthreads();
    {   // stage 2
        using I = StageIdx<N, P::R2, P::R1>;
#pragma unroll
        for (int b = 0; b < I::NB; ++b)
#pragma unroll
            for (int m = 0; m < P::R2; ++m) x[b * P::R2 + m] = S(I::in(j, b, m));
        stage_compute<N, P::R2, P::R1, DIR>(x, j, tw);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < I::NB; ++b)
#pragma unroll
            for (int q = 0; q < P::R2; ++q) S(I::out(j, b, q)) = x[b * P::R2 + q];
    }
    __syncthreads();
    if (P::R3 > 1) {   // stage 3
        constexpr int R3 = P::R3 > 1 ? P::R3 : 2;
        using I = StageIdx<N, R3, P::R1 * P::R2>;
#pragma unroll
        for (int b = 0; b < I::NB; ++b)
#pragma unroll
            for (int m = 0; m < R3; ++m) x[b * R3 + m] = S(I::in(j, b, m));
        stage_compute<N, R3, P::R1 * P::R2, DIR>(x, j, tw + (P::R2 - 1) * P::R1);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < I::NB; ++b)
#pragma unroll
            for (int q = 0; q < R3; ++q) S(I::out(j, b, q)) = x[b * R3 + q];
        __syncthreads();
    }
}

}  // namespace b4dfft
