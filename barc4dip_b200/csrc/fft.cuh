// In-CTA power-of-two complex FFTs for sm_100a: register radix-16/8/4/2 butterflies on packed
// f32x2 arithmetic (FADD2 / FMUL2 / FFMA2 with the half-swap and sign-pattern operand modifiers),
// Stockham autosort exchanges through padded shared memory.
//
// A length-N transform is owned by N/16 threads; every thread keeps 16 complex values in
// registers per stage.  Stage s with accumulated length LS and radix R (virtual thread v):
//     k = v % LS
//     x[m] = z[v + m*N/R] * exp(DIR * 2 pi i * m * k / (LS*R)),   m = 0..R-1
//     X    = DFT_R(x)
//     z[(v - k)*R + k + q*LS] = X[q]
// Reads are unit-stride in v for every stage; writes are made conflict-free by padding the
// array by one element every 16 (pad16).  For N >= 256 every index above is "a per-thread base
// + a compile-time constant" (N/16 is a multiple of 16, so pad16(j + C) = pad16(j) + 17C/16), which
// turns the exchanges into LDS/STS with immediate offsets.
#pragma once

#include <cuda_runtime.h>

namespace b4dfft {

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kCos1_16 = 0.92387953251128675613f;   // cos(pi/8)
constexpr float kSin1_16 = 0.38268343236508977173f;   // sin(pi/8)

__host__ __device__ constexpr int pad16(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4); }

// ---- packed complex helpers (one instruction each on sm_100a) -------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// a * w  (FMUL2 with a broadcast operand + FFMA2 with a half-swapped operand)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    const float2 t = __fmul2_rn(make_float2(a.x, a.x), w);
    return __ffma2_rn(make_float2(a.y, a.y), make_float2(-w.y, w.x), t);
}
// The same product written as a * w.x + (half-swapped, (-, +) signed) a * w.y: the compiler contracts it into FMUL2 +
// FFMA2 with the swap / sign pattern on the addend and builds no (-w.y, w.x) register pair (one FADD + one or two MOV per
// twiddle in the form above). Same rounding class (one rounded product, one fused step). Measured per kernel: the column
// pass gains 3 % (2.50 -> 2.43 ms per 128 frames), the row passes lose 3 %, so the variant is a template argument.
template <int CM>
__device__ __forceinline__ float2 cmulv(float2 a, float2 w) {
    if (CM == 0) return cmul(a, w);
    const float2 u = __fmul2_rn(a, make_float2(w.x, w.x)), v = __fmul2_rn(a, make_float2(w.y, w.y));
    return __fadd2_rn(u, make_float2(-v.y, v.x));
}
// a + (DIR*i) * b   and   a - (DIR*i) * b      (DIR = -1: -i, the forward W4)
template <int DIR>
__device__ __forceinline__ float2 add_rot90(float2 a, float2 b) {
    return DIR < 0 ? __fadd2_rn(a, make_float2(b.y, -b.x)) : __fadd2_rn(a, make_float2(-b.y, b.x));
}
template <int DIR>
__device__ __forceinline__ float2 sub_rot90(float2 a, float2 b) {
    return DIR < 0 ? __fadd2_rn(a, make_float2(-b.y, b.x)) : __fadd2_rn(a, make_float2(b.y, -b.x));
}
// multiply by DIR * i
template <int DIR>
__device__ __forceinline__ float2 rot90(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// multiply by exp(DIR * i pi/4) and exp(DIR * 3 i pi/4)
template <int DIR>
__device__ __forceinline__ float2 rot45(float2 a) {
    const float2 s = DIR < 0 ? __fadd2_rn(make_float2(a.x, a.y), make_float2(a.y, -a.x))
                             : __fadd2_rn(make_float2(a.x, a.y), make_float2(-a.y, a.x));
    return __fmul2_rn(s, make_float2(kSqrtHalf, kSqrtHalf));
}
template <int DIR>
__device__ __forceinline__ float2 rot135(float2 a) {
    const float2 s = DIR < 0 ? __fadd2_rn(make_float2(-a.x, -a.y), make_float2(a.y, -a.x))
                             : __fadd2_rn(make_float2(-a.x, -a.y), make_float2(-a.y, a.x));
    return __fmul2_rn(s, make_float2(kSqrtHalf, kSqrtHalf));
}
// multiply by (c, DIR*s)
template <int DIR>
__device__ __forceinline__ float2 rotcs(float2 a, float c, float s) {
    return cmul(a, make_float2(c, DIR < 0 ? -s : s));
}

template <int DIR>
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// in-place 4-point DFT of (a0,a1,a2,a3) -> outputs X0..X3 in the same slots (8 packed instructions)
template <int DIR>
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), t = csub(a1, a3);
    a0 = cadd(s0, s2);
    a2 = csub(s0, s2);
    a1 = add_rot90<DIR>(s1, t);
    a3 = sub_rot90<DIR>(s1, t);
}

// 8-point DFT: inner radix-2 over (x[a], x[a+4]), twiddle W8^(a*q1), outer radix-4 over a.
// Input x[0..7] natural order, output X[q] natural order in the same array.
template <int DIR>
__device__ __forceinline__ void bfly8(float2* x) {
    bfly2<DIR>(x[0], x[4]);
    bfly2<DIR>(x[1], x[5]);
    bfly2<DIR>(x[2], x[6]);
    bfly2<DIR>(x[3], x[7]);
    x[5] = rot45<DIR>(x[5]);
    x[6] = rot90<DIR>(x[6]);
    x[7] = rot135<DIR>(x[7]);
    bfly4<DIR>(x[0], x[1], x[2], x[3]);   // q1 = 0: X[0 + 2 q2] in slots 0..3
    bfly4<DIR>(x[4], x[5], x[6], x[7]);   // q1 = 1: X[1 + 2 q2] in slots 4..7
    const float2 t1 = x[1], t2 = x[2], t3 = x[3], t4 = x[4], t5 = x[5], t6 = x[6];
    x[1] = t4; x[2] = t1; x[3] = t5; x[4] = t2; x[5] = t6; x[6] = t3;
}

// 16-point DFT as 4 x 4: inner radix-4 over (x[a], x[a+4], x[a+8], x[a+12]) -> t[a][q1],
// twiddle W16^(a*q1), outer radix-4 over a -> X[q1 + 4 q2].
template <int DIR>
__device__ __forceinline__ void bfly16(float2* x) {
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4<DIR>(x[a], x[a + 4], x[a + 8], x[a + 12]);
    x[5] = rotcs<DIR>(x[5], kCos1_16, kSin1_16);      // a=1,q1=1: W16^1
    x[9] = rot45<DIR>(x[9]);                           // a=1,q1=2: W16^2
    x[13] = rotcs<DIR>(x[13], kSin1_16, kCos1_16);     // a=1,q1=3: W16^3
    x[6] = rot45<DIR>(x[6]);                           // a=2,q1=1: W16^2
    x[10] = rot90<DIR>(x[10]);                         // a=2,q1=2: W16^4
    x[14] = rot135<DIR>(x[14]);                        // a=2,q1=3: W16^6
    x[7] = rotcs<DIR>(x[7], kSin1_16, kCos1_16);       // a=3,q1=1: W16^3
    x[11] = rot135<DIR>(x[11]);                        // a=3,q1=2: W16^6
    x[15] = rotcs<DIR>(x[15], -kCos1_16, -kSin1_16);   // a=3,q1=3: W16^9 = -W16^1
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) bfly4<DIR>(x[4 * q1], x[4 * q1 + 1], x[4 * q1 + 2], x[4 * q1 + 3]);
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
template <> struct Plan<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 16; };   // Bluestein convolutions only
template <> struct Plan<2048> { static constexpr int R1 = 16, R2 = 16, R3 = 8; };
template <> struct Plan<1024> { static constexpr int R1 = 16, R2 = 16, R3 = 4; };
template <> struct Plan<512>  { static constexpr int R1 = 16, R2 = 8,  R3 = 4; };
template <> struct Plan<256>  { static constexpr int R1 = 16, R2 = 16, R3 = 1; };
template <> struct Plan<128>  { static constexpr int R1 = 16, R2 = 8,  R3 = 1; };

// ================================================================================================
// Register-to-register transform (v2 core).
//
// Input : x[m] = data[j + m*T], T = N/16 (thread j of the transform, natural strided ownership).
// Output: x[s] = DFT(data)[j + s*T]  -- the SAME ownership, so epilogues and the next transform of a chain work
//         straight from registers: only the two inner exchanges go through shared memory (the last stage of a
//         Stockham pass writes element v + q*LS with LS*R = N, i.e. thread-local slots).
// Twiddles: per stage the thread loads the base powers w^1, w^2, w^4 (, w^8) of its own k from a small table
//         (two LDG.128 issued BEFORE the exchange barrier, while x[] is dead) and forms the other powers by
//         at most two complex multiplications (error <= 7 ulp).
// Table layout (float2): [0, 32)  stage 2: k in [0,16) -> (w1, w2);  [32, 64)  -> (w4, w8), base 16*R2
//                        [64, 64 + 2*LS3)  stage 3: k in [0,LS3) -> (w1, w2);  [64 + 2*LS3, ...) -> (w4, w8), or bare w4 for a radix <= 8, base N
//         (two arrays per stage: the 16-byte fetches of a warp's consecutive k are then contiguous -- 4 wavefronts of
//          the load/store pipe per fetch instead of the 8 an interleaved (w1, w2, w4, w8) record costs; w4 alone is an
//          8-byte fetch when the radix needs no w8)
// Synchronisation: a __syncthreads() is issued on entry (z may still be read by a previous user); on return
//         other threads may still be reading z, so the caller synchronises before writing z itself.
// ================================================================================================
template <int N>
__host__ __device__ constexpr int twiddle_base_count() {
    return 64 + (Plan<N>::R3 > 1 ? 4 * Plan<N>::R1 * Plan<N>::R2 : 0);
}

template <int DIR>
__device__ __forceinline__ float2 tw_dir(float2 w) { return DIR > 0 ? make_float2(w.x, -w.y) : w; }

// x[m] *= w^m for m = 1..R-1 given the base powers (already conjugated for DIR > 0)
template <int R, int CM = 0>
__device__ __forceinline__ void apply_twiddle_powers(float2* x, float2 w1, float2 w2, float2 w4, float2 w8) {
    x[1] = cmulv<CM>(x[1], w1);
    if (R > 2) {
        const float2 w3 = cmulv<CM>(w1, w2);
        x[2] = cmulv<CM>(x[2], w2);
        x[3] = cmulv<CM>(x[3], w3);
        if (R > 4) {
            x[4] = cmulv<CM>(x[4], w4);
            x[5] = cmulv<CM>(x[5], cmulv<CM>(w4, w1));
            x[6] = cmulv<CM>(x[6], cmulv<CM>(w4, w2));
            const float2 w7 = cmulv<CM>(w4, w3);
            x[7] = cmulv<CM>(x[7], w7);
            if (R > 8) {
                x[8] = cmulv<CM>(x[8], w8);
                x[9] = cmulv<CM>(x[9], cmulv<CM>(w8, w1));
                x[10] = cmulv<CM>(x[10], cmulv<CM>(w8, w2));
                x[11] = cmulv<CM>(x[11], cmulv<CM>(w8, w3));
                x[12] = cmulv<CM>(x[12], cmulv<CM>(w8, w4));
                x[13] = cmulv<CM>(x[13], cmulv<CM>(w8, cmulv<CM>(w4, w1)));
                x[14] = cmulv<CM>(x[14], cmulv<CM>(w8, cmulv<CM>(w4, w2)));
                x[15] = cmulv<CM>(x[15], cmulv<CM>(w8, w7));
            }
        }
    }
}

// Base powers (w, w^2, w^4, w^8) of a stage's twiddle from the table. DERIVE: only (w, w^2) are fetched -- one LDG.128 --
// and w^4, w^8 come from squaring (two packed instructions each, and only the powers radix R needs): a table load costs
// the load/store pipe four wavefronts per warp, and the six of a 2048-point transform were a sixth of the pass's exchange
// traffic. The squared powers carry ~3 (w^4) and ~7 ulp (w^8) instead of 0.5, which is immaterial for the inverse
// transforms (their inputs are whitened unit phasors or |F|^2, their outputs are compared at 1e-5 of the peak) but not for
// the forward ones: on noise-free band-limited frames the whitening turns the rounding noise of the forward transform in
// the empty bins into unit phasors, and the tracker's sub-pixel answer on such a frame moved by 0.1 px with derived
// forward twiddles (the reference's own float32 / float64 paths differ by 0.03 px there). Hence: derived powers in the
// inverse transforms (DIR > 0), table powers in the forward ones. Measured: all transforms derived 6.645 -> 6.55 ms per
// 128 frames.
#ifndef B4D_TW_DERIVE
#define B4D_TW_DERIVE 1           // 0: never derive, 1: inverse transforms only, 3: every transform (experiments)
#endif
template <int R, bool DERIVE>
__device__ __forceinline__ void ldg_tw4(const float2* pa, const float2* pb, float2& a, float2& b, float2& c, float2& d) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(pa));
    a = make_float2(u.x, u.y); b = make_float2(u.z, u.w);
    c = d = make_float2(0.f, 0.f);
    if (DERIVE) {
        if (R > 4) c = cmulv<1>(b, b);
        if (R > 8) d = cmulv<1>(c, c);
    } else if (R > 8) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pb));
        c = make_float2(v.x, v.y); d = make_float2(v.z, v.w);
    } else if (R > 4) {
        c = __ldg(pb);
    }
}

// Barrier of the threads that share an exchange buffer. GROUP = 0: the whole CTA (__syncthreads). GROUP = 1: the N/16
// threads of one transform only (named barrier 1 + group), legal when they are whole warps (BATCH == 1; for N < 512 the
// caller pads the group to one warp with threads that repeat the work of the first N/16): the independent transforms
// of a CTA then stop waiting for one another at every exchange. GROUP = 2: the first
// BATCH * N/16 threads of the CTA (named barrier 15), the others having left for good or for other work.
template <int N, int GROUP, int BATCH>
__device__ __forceinline__ void fft_sync(int group) {
    if (GROUP == 1) asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(N / 16 < 32 ? 32 : N / 16) : "memory");
    else if (GROUP == 2) asm volatile("bar.sync 15, %0;" ::"n"(BATCH * (N / 16)) : "memory");
    else __syncthreads();
}

template <int N, int DIR, int BATCH, int GROUP = 0, int CM = 0>
__device__ __forceinline__ void fft_regs(float2* x, int j, float2* __restrict__ z, const float2* __restrict__ twb, int group = 0) {
    static_assert(GROUP != 1 || BATCH == 1, "group barriers: one transform per group (N < 512: the caller pads the group to a warp)");
    static_assert(GROUP != 2 || (BATCH * (N / 16)) % 32 == 0, "sub-CTA barriers need whole warps");
    using P = Plan<N>;
    constexpr int T = N / 16;
    constexpr bool FAST = (T % 16) == 0;
    constexpr bool TWD = (B4D_TW_DERIVE == 3) || (B4D_TW_DERIVE == 1 && DIR > 0);
    constexpr int R2 = P::R2, NB2 = 16 / R2;
    constexpr bool HAS3 = P::R3 > 1;
    constexpr int R3 = HAS3 ? P::R3 : 2, NB3 = 16 / R3, LS3 = 16 * R2;
    const int pj = pad16(j);

    // ---- stage 1: radix 16, LS = 1 -> element 16 j + q at padded index 17 j + q
    bfly16<DIR>(x);
    fft_sync<N, GROUP, BATCH>(group);
    {
        float2* w = z + 17 * j * BATCH;
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q * BATCH] = x[q];
    }
    // stage-2 base twiddles of every block of this thread, in flight across the barrier
    float2 b1[NB2], b2[NB2], b4[NB2], b8[NB2];
#pragma unroll
    for (int b = 0; b < NB2; ++b) {
        const int k = FAST ? (j & 15) : ((j + b * T) & 15);
        if (FAST && b > 0) { b1[b] = b1[0]; b2[b] = b2[0]; b4[b] = b4[0]; b8[b] = b8[0]; }
        else ldg_tw4<R2, TWD>(twb + 2 * k, twb + 32 + (R2 > 8 ? 2 : 1) * k, b1[b], b2[b], b4[b], b8[b]);
    }
    fft_sync<N, GROUP, BATCH>(group);

    // ---- stage 2: radix R2, LS = 16
    {
        if (FAST) {
            const float2* r = z + pj * BATCH;
#pragma unroll
            for (int b = 0; b < NB2; ++b)
#pragma unroll
                for (int m = 0; m < R2; ++m) x[b * R2 + m] = r[((b * T + m * (N / R2)) / 16 * 17) * BATCH];
        } else {
#pragma unroll
            for (int b = 0; b < NB2; ++b)
#pragma unroll
                for (int m = 0; m < R2; ++m) x[b * R2 + m] = z[pad16(j + b * T + m * (N / R2)) * BATCH];
        }
#pragma unroll
        for (int b = 0; b < NB2; ++b) {
            apply_twiddle_powers<R2, CM>(x + b * R2, tw_dir<DIR>(b1[b]), tw_dir<DIR>(b2[b]), tw_dir<DIR>(b4[b]), tw_dir<DIR>(b8[b]));
            bfly<R2, DIR>(x + b * R2);
        }
    }
    if constexpr (!HAS3) {
        // last stage: block b, output q is element j + T (b + NB2 q)
        float2 y[16];
#pragma unroll
        for (int b = 0; b < NB2; ++b)
#pragma unroll
            for (int q = 0; q < R2; ++q) y[b + NB2 * q] = x[b * R2 + q];
#pragma unroll
        for (int s = 0; s < 16; ++s) x[s] = y[s];
    } else {
        fft_sync<N, GROUP, BATCH>(group);
        if (FAST) {
            const int k = j & 15;
            float2* w = z + (((j - k) / 16) * 17 * R2 + k) * BATCH;
#pragma unroll
            for (int b = 0; b < NB2; ++b)
#pragma unroll
                for (int q = 0; q < R2; ++q) w[((b * T / 16) * 17 * R2 + 17 * q) * BATCH] = x[b * R2 + q];
        } else {
#pragma unroll
            for (int b = 0; b < NB2; ++b) {
                const int v = j + b * T, kk = v & 15;
#pragma unroll
                for (int q = 0; q < R2; ++q) z[pad16((v - kk) * R2 + kk + q * 16) * BATCH] = x[b * R2 + q];
            }
        }
        float2 c1[NB3], c2[NB3], c4[NB3], c8[NB3];
#pragma unroll
        for (int b = 0; b < NB3; ++b)
            ldg_tw4<R3, TWD>(twb + 64 + 2 * (j + b * T), twb + 64 + 2 * LS3 + (R3 > 8 ? 2 : 1) * (j + b * T), c1[b], c2[b], c4[b], c8[b]);
        fft_sync<N, GROUP, BATCH>(group);

        // ---- stage 3: radix R3, LS = 16 R2 = N / R3: k = v = j + b T
        if (FAST) {
            const float2* r = z + pj * BATCH;
#pragma unroll
            for (int b = 0; b < NB3; ++b)
#pragma unroll
                for (int m = 0; m < R3; ++m) x[b * R3 + m] = r[((b * T + m * (N / R3)) / 16 * 17) * BATCH];
        } else {
#pragma unroll
            for (int b = 0; b < NB3; ++b)
#pragma unroll
                for (int m = 0; m < R3; ++m) x[b * R3 + m] = z[pad16(j + b * T + m * (N / R3)) * BATCH];
        }
#pragma unroll
        for (int b = 0; b < NB3; ++b) {
            apply_twiddle_powers<R3, CM>(x + b * R3, tw_dir<DIR>(c1[b]), tw_dir<DIR>(c2[b]), tw_dir<DIR>(c4[b]), tw_dir<DIR>(c8[b]));
            bfly<R3, DIR>(x + b * R3);
        }
        // block b, output q is element v + q LS3 = j + T (b + NB3 q)
        float2 y[16];
#pragma unroll
        for (int b = 0; b < NB3; ++b)
#pragma unroll
            for (int q = 0; q < R3; ++q) y[b + NB3 * q] = x[b * R3 + q];
#pragma unroll
        for (int s = 0; s < 16; ++s) x[s] = y[s];
        (void)LS3;
    }
}

// v2 core + write-back: on return the transform sits in shared memory in natural order (element i at
// z[pad16(i) * BATCH]) and a __syncthreads() has been issued.
template <int N, int DIR, int BATCH, int GROUP = 0>
__device__ __forceinline__ void fft_regs_to_smem(float2* x, int j, float2* __restrict__ z, const float2* __restrict__ twb, int group = 0) {
    constexpr int T = N / 16;
    fft_regs<N, DIR, BATCH, GROUP>(x, j, z, twb, group);
    fft_sync<N, GROUP, BATCH>(group);
    if ((T % 16) == 0) {
        float2* w = z + pad16(j) * BATCH;
#pragma unroll
        for (int s = 0; s < 16; ++s) w[(s * T / 16 * 17) * BATCH] = x[s];
    } else {
#pragma unroll
        for (int s = 0; s < 16; ++s) z[pad16(j + s * T) * BATCH] = x[s];
    }
    __syncthreads();
}

}  // namespace b4dfft
