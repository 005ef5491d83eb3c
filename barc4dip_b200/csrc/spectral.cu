// The 2-D FFT family: batched real 2-D FFT with fused PSD / autocorrelation / phase-correlation
// epilogues (signal/fft.py, signal/corr.py, signal/tracking.py of the reference).
//
// Decomposition of one (ny, nx) real frame (all transforms in-CTA, see fft.cuh):
//   K1 rows_fwd : rows are paired (z = row_a + i row_b), one complex FFT per pair, split into the
//                 two half spectra (kx = 0..nx/2-1, Nyquist packed into Im of kx = 0), written to
//                 a blocked intermediate H[kx/8][y][kx%8] so that K2 reads contiguous 64 B runs.
//   K2 cols     : a CTA owns 8 adjacent kx columns (all ky): FFT along y, then fused epilogues
//                 - |F|^2 * scale with fftshift + Hermitian mirror stores (psd2d), spectral sums;
//                 - complex spectrum out (fft2d) or conj spectrum to the blocked layout (reference
//                   spectrum of the tracker / second operand of xcorr2d);
//                 - |F|^2 (autocorr) or whitened F * R (phase correlation) followed, in the same
//                   CTA, by the inverse FFT along y, written to blocked intermediates.
//   K3 rows_inv : two half-spectrum rows (two rows of one map, or one row of two maps) form one
//                 complex inverse FFT whose real / imaginary parts are the two real output rows;
//                 fftshifted stores, peak normalisation, argmax partials.
// Every frame is shifted by a pilot mean K before the FFT (K*nx*ny is added back to the DC bin),
// so fp32 rounding is relative to the fluctuations, not to the pedestal.
#include <cuda.h>            // CUtensorMap (the driver entry point itself is fetched through the runtime)

#include "common.cuh"
#include "fft.cuh"
#include "select.cuh"
#include "reduce.cuh"

using namespace b4dfft;

int b4d_select_impl(b4d_ctx* ctx, const float* stack, int64_t T, int64_t n, const double* q_dev, int n_q,
                    int use_abs, float* out, int64_t* n_valid);
int b4d_put_doubles(b4d_ctx* ctx, double* dst, const double* src_host, int n);

struct RefBuffers {                   // device buffers of one tracker reference (see b4d_ref)
    float2* ref = nullptr;
    float2* ref_nyq = nullptr;
    float2* gref = nullptr;
    size_t gref_elems = 0;
    int ny = 0, nx = 0;
};

struct FftPlanCache {
    // released reference buffers, kept for the next b4d_phase_reference_create of the same frame shape (a tracker per
    // stack, or per frame in incremental tracking, would otherwise pay a cudaMalloc / cudaFree pair of 17 MB each time)
    std::vector<RefBuffers> ref_pool;
    float2* twb[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // base-power twiddle tables: 128 ... 4096
    // tracker reference: conj spectrum of the embedded z-scored template
    float2* ref = nullptr;       // blocked (nx/2/8, ny, 8)
    float2* ref_nyq = nullptr;   // (ny)
    int ref_ny = 0, ref_nx = 0;
    double* theta = nullptr;     // 1130 x (sin, cos)
    int* blk_list = nullptr;     // fused median: sample row blocks first, then all the others
    int blk_n = 0, blk_ns = 0;
    void* gen = nullptr;         // GenCache*: chirp tables of the arbitrary-length path (generic_dft.cuh)
    float2* gref = nullptr;      // tracker reference for frames that are not powers of two: conj spectrum, natural order
    size_t gref_elems = 0;
};

namespace {

constexpr int NSP = 6;    // spectral partial sums per CTA
constexpr int NTHETA = 1130;   // int(2*pi*180), maths/radial.py:150

inline int log2i(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
inline bool fft_size_ok(int n) { return n >= 128 && n <= 2048 && (n & (n - 1)) == 0; }

// base powers (w, w^2, w^4, w^8) per stage and k, split into a (w, w^2) array and a (w^4, w^8) array per stage so that a
// warp's 16-byte fetches are contiguous (layout in fft.cuh, v2 core)
template <int N>
void fill_twiddle_bases(std::vector<float2>& h) {
    using P = Plan<N>;
    h.assign(twiddle_base_count<N>(), make_float2(0.f, 0.f));
    // (the second array holds (w^4, w^8) records for a radix-16 stage and bare w^4 for a radix-8 one: contiguous either way)
    auto put = [&](size_t a_at, size_t b_at, int k, double base, int radix) {
        const int npow = radix > 8 ? 4 : 3, rec = radix > 8 ? 2 : 1;
        for (int p = 0; p < npow; ++p) {
            const double a = -2.0 * 3.14159265358979323846 * (double)(1 << p) * (double)k / base;
            const size_t at = p < 2 ? a_at + 2 * (size_t)k + p : b_at + (size_t)rec * k + (p - 2);
            h[at] = make_float2((float)cos(a), (float)sin(a));
        }
    };
    for (int k = 0; k < 16; ++k) put(0, 32, k, 16.0 * P::R2, P::R2);
    if (P::R3 > 1) {
        const int ls3 = 16 * P::R2;
        for (int k = 0; k < ls3; ++k) put(64, 64 + 2 * (size_t)ls3, k, (double)N, P::R3);
    }
}

int get_twiddle_bases(b4d_ctx* ctx, int n, const float2** out) {
    if (!ctx->fft) ctx->fft = new FftPlanCache();
    const int slot = log2i(n) - 7;
    if (!ctx->fft->twb[slot]) {
        std::vector<float2> h;
        switch (n) {
            case 128: fill_twiddle_bases<128>(h); break;
            case 256: fill_twiddle_bases<256>(h); break;
            case 512: fill_twiddle_bases<512>(h); break;
            case 1024: fill_twiddle_bases<1024>(h); break;
            case 4096: fill_twiddle_bases<4096>(h); break;
            default: fill_twiddle_bases<2048>(h); break;
        }
        B4D_CUDA(ctx, cudaMalloc(&ctx->fft->twb[slot], h.size() * sizeof(float2)));
        B4D_CUDA(ctx, cudaMemcpy(ctx->fft->twb[slot], h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    *out = ctx->fft->twb[slot];
    return B4D_OK;
}

// =================================================================================================
// K1: rows forward
// =================================================================================================
constexpr int TC = 8;     // width of a tile of the blocked intermediates [kx/TC][y][kx%TC]
#ifndef B4D_PAIR_DEFAULT
#define B4D_PAIR_DEFAULT 0        // frames per reduce / forward-rows pair of the whole-batch schedule (0: off)
#endif
#ifndef B4D_SCHED_SUB_DEFAULT
#define B4D_SCHED_SUB_DEFAULT 0   // frames per step of the frame-pipelined schedule (0: whole batches, one kernel after the other)
#endif
#ifndef B4D_COLS_CW_2048
#define B4D_COLS_CW_2048 4   // columns per K2 CTA at ny = 2048 (4: two CTAs per SM; 8: one)
#endif

// Cache policy of the row <-> column intermediates (bits of the `keep` argument of the four FFT kernels):
//   0  streaming both ways (evict-first): the intermediates of a large batch go through HBM anyway and must not push the
//      reference spectrum and the prefetched tiles out of L2;
//   1  stores stay in L2 (st.global.cg): a few frames at a time, the consumer finds them there;
//   2  loads leave the line at normal priority instead of marking it evict-first;
//   4  the column pass drops its input tile from L2 once every thread has read it (discard.global.L2: the dead, dirty
//      lines are neither kept nor written back to HBM).
// The bits are honoured only by experiment builds (python -m barc4dip_b200.build --variant keep -DB4D_KEEP_POLICY=1): a
// run-time policy switch compiles into two predicated memory instructions per access -- one of them always off, each
// taking an issue slot -- in kernels that are bound by instruction issue and by the load/store pipe. The default build
// streams both ways (policy 0, the whole-batch schedule's).
#ifndef B4D_KEEP_POLICY
#define B4D_KEEP_POLICY 0
#endif
__device__ __forceinline__ void st_inter(float2* p, float2 v, int keep) {
    if (B4D_KEEP_POLICY && (keep & 1)) __stcg(p, v); else __stcs(p, v);
}
__device__ __forceinline__ float2 ld_inter(const float2* p, int keep) {
    return (B4D_KEEP_POLICY && (keep & 2)) ? __ldcg(p) : __ldcs(p);
}

struct RowsFwdArgs {
    const float* stack;
    const float* gain;
    const float* dark;
    const float* pilot;   // nullable
    float2* H;
    const float2* tw;
    int ny;
    double* mom;          // nullable: (T, gridDim.x, 2) per-CTA sums of d = x - K and of d^2 (z-score of the tracker)
    int pf_dist;          // L2 prefetch distance in CTAs (0 = off), set by the launcher
    int keep;             // cache policy of the intermediate (see st_inter)
};

template <int NX>
__global__ void __launch_bounds__(512) rows_fwd_kernel(RowsFwdArgs a) {
    __shared__ double s_mom[16][2];
    constexpr int TPF = NX / 16;          // threads per transform
    constexpr int FPC = 512 / TPF;        // transforms per CTA
    constexpr int ROWS = 2 * FPC;
    constexpr int FS = padded_len(NX) + 8;
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x, f = tid / TPF, j = tid % TPF;
    const int64_t t = blockIdx.y;
    const int y0 = blockIdx.x * ROWS;
    const float K = a.pilot ? __ldg(a.pilot + t) : 0.f;
    const size_t fo = (size_t)t * a.ny * NX;
    const float* ra = a.stack + fo + (size_t)(y0 + 2 * f) * NX;
    const float* rb = ra + NX;
    float2 x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int col = j + m * TPF;
        float va = __ldcs(ra + col), vb = __ldcs(rb + col);
        if (a.gain) {
            const size_t pa = (size_t)(y0 + 2 * f) * NX + col, pb = pa + NX;
            va = (va - (a.dark ? __ldg(a.dark + pa) : 0.f)) * __ldg(a.gain + pa);
            vb = (vb - (a.dark ? __ldg(a.dark + pb) : 0.f)) * __ldg(a.gain + pb);
        }
        x[m] = make_float2(va - K, vb - K);
    }
    // the rows of the CTA that takes this SM next (pf_dist CTAs ahead in launch order; a CTA's rows are one contiguous
    // run of 16384 floats) are pulled into L2 now, one 128-byte line per thread
    if (a.pf_dist > 0) {
        const size_t next = (size_t)t * gridDim.x + blockIdx.x + (size_t)a.pf_dist;
        if (next < (size_t)gridDim.x * gridDim.y)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.stack + next * (size_t)(ROWS * NX) + (size_t)tid * 32));
    }
    if (a.mom) {
        // the tracker z-scores the frame (tracking.py:308-311): its mean and standard deviation come from these sums,
        // 32 pixels in fp32 per thread, fp64 from there on, fixed order
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < 16; ++m) { s1 = __fadd2_rn(s1, x[m]); s2 = __ffma2_rn(x[m], x[m], s2); }
        const double w1 = warp_sum((double)s1.x + (double)s1.y), w2 = warp_sum((double)s2.x + (double)s2.y);
        if ((tid & 31) == 0) { s_mom[tid >> 5][0] = w1; s_mom[tid >> 5][1] = w2; }
    }
    float2* z = sm + f * FS;
    fft_regs<NX, -1, 1, (NX >= 1024)>(x, j, z, a.tw, f);
    // Split Z = FFT(a + i b) into the half spectra of rows a and b: H_a[k] = (Z[k] + conj Z[NX - k]) / 2, H_b[k] = (Z[k] -
    // conj Z[NX - k]) / (2i), k < NX/2. Thread j holds Z[j + s TPF], s = 0..15: its slots 0..7 ARE the Z[k] of its own
    // eight outputs and stay in registers; the Z[NX - k] are slots 8..15 of thread TPF - j of the same transform, so only
    // the upper halves cross shared memory (8 stores + 8 loads per thread, barriers of the transform's own threads).
    fft_sync<NX, (NX >= 1024), 1>(f);                 // the other threads of the transform are done reading their exchanges
    constexpr bool LIN = (TPF % 16) == 0;
    constexpr int PS = TPF + TPF / 16;
    if (LIN) {
        float2* w = z + pad16(j);
#pragma unroll
        for (int s2 = 8; s2 < 16; ++s2) w[PS * s2] = x[s2];
    } else {
#pragma unroll
        for (int s2 = 8; s2 < 16; ++s2) z[pad16(j + s2 * TPF)] = x[s2];
    }
    if (a.mom) __syncthreads(); else fft_sync<NX, (NX >= 1024), 1>(f);
    if (a.mom && tid < 2) {                           // (the barrier above made every warp's partial sums visible)
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 16; ++w) v += s_mom[w][tid];
        a.mom[((size_t)t * gridDim.x + blockIdx.x) * 2 + tid] = v;
    }
    {
        // pad16(NX - j - s TPF) = padded_len(NX) - j - ceil(j / 16) - s (TPF + TPF / 16) when TPF is a multiple of 16
        const float2* zm = z + (padded_len(NX) - j - ((j + 15) >> 4));
        // blocked store: k = j + s TPF lives in tile j/8 + s TPF/8, rows y0 + 2f (a) and y0 + 2f + 1 (b) are 8 elements apart
        float2* Hf = a.H + (size_t)t * a.ny * (NX / 2) + ((size_t)(j / TC) * a.ny + y0 + 2 * f) * TC + (j % TC);
        const size_t tstride = (size_t)(TPF / TC) * a.ny * TC;
        const float2 hh = make_float2(0.5f, 0.5f);
#pragma unroll
        for (int s2 = 0; s2 < 8; ++s2) {
            const float2 Z = x[s2];
            float2 va, vb;
            if (s2 == 0) {
                // thread 0: k = 0 carries the two real samples DC and Nyquist (Z[NX/2] is its own slot 8, just written)
                const float2 Zm = j == 0 ? z[pad16(NX / 2)] : (LIN ? zm[0] : z[pad16(NX - j)]);
                va = __fmul2_rn(__fadd2_rn(Z, make_float2(Zm.x, -Zm.y)), hh);
                vb = __fmul2_rn(__fadd2_rn(make_float2(Z.y, -Z.x), make_float2(Zm.y, Zm.x)), hh);
                if (j == 0) { va = make_float2(Z.x, Zm.x); vb = make_float2(Z.y, Zm.y); }
            } else {
                const float2 Zm = LIN ? zm[-PS * s2] : z[pad16(NX - j - s2 * TPF)];
                va = __fmul2_rn(__fadd2_rn(Z, make_float2(Zm.x, -Zm.y)), hh);
                vb = __fmul2_rn(__fadd2_rn(make_float2(Z.y, -Z.x), make_float2(Zm.y, Zm.x)), hh);
            }
            st_inter(Hf + s2 * tstride, va, a.keep);  // (streaming unless the consumer follows closely)
            st_inter(Hf + s2 * tstride + TC, vb, a.keep);
        }
    }
}

// =================================================================================================
// K2: columns forward + epilogues (+ inverse along y)
// =================================================================================================
struct ColsArgs {
    const float2* H;            // blocked input, per frame ny*nx/2
    const float2* tw;
    const float* pilot;         // nullable: DC += nx*ny*K
    int nx;
    int zero_dc;                // clear F[0,0] (mean removal) for every output
    int ac_zero_dc;             // clear it for the autocorrelation branch only (fused pipeline)
    int pf_dist;                // L2 prefetch distance in CTAs (resident CTAs of the launch; 0 = off), set by the launcher
    // plain outputs
    float2* cplx_out;           // (T, ny, nx) shifted complex spectrum (fft2d)
    float* psd_out;             // (T, ny, nx) shifted |F|^2 * psd_scale
    float psd_scale;
    double* spec_partials;      // (T, ntiles, NSP)
    float2* conj_out;           // blocked conj(F) (T, nx/2/TC, ny, TC)
    float2* conj_nyq_out;       // (T, ny)
    // autocorrelation branch
    float2* i2_ac;              // inverse-along-y of |F|^2, column pairs packed (see the end of cols_kernel): (T, nx/4/TC, ny, TC)
    float2* i2_ac_nyq;          // (T, ny) inverse-along-y of the Nyquist column's |F|^2 (packed variant only)
    double* ac_partials;        // (T, ntiles) sum of the full-spectrum |F|^2 owned by the tile
    // product branch: G = F * R (optionally whitened), inverse along y -> i2_pc
    float2* i2_pc;              // blocked, row pairs interleaved ([kx/8][y/2][kx%8][y%2], see cols_body)
    const float2* R;            // blocked, shared by all frames (ref_stride 0) or per frame
    const float2* Rnyq;
    size_t r_stride, rnyq_stride;
    int whiten;
    float eps;
    const double* fr;           // frame-reduction table (mean, m2) for the z-score; nullable
    int fr_stride;
    int keep;                   // cache policy of the intermediates (see st_inter)
    int psd_tma;                // the PSD map leaves through TMA tensor stores (set by the launcher, see cols_body)
};

// ---- TMA (cp.async.bulk.tensor) helpers of the column pass --------------------------------------------------------
// The PSD map is row-major (T, ny, nx) and a CTA owns 8 adjacent columns: written from registers that is one 32-byte
// piece per row and thread, twice (the Hermitian mirror), 0.52 M of the pass's 2.4 M load/store wavefronts per frame. With
// psd_tma the CTA instead stages the scaled |F|^2 tile in shared memory -- the exchange buffer is idle between the
// forward transform and the first exchange of the inverse one -- as a dense [ny][8] float tile in OUTPUT row order, and
// one thread hands it to the TMA unit as ny/256 boxes of 256 rows x 8 columns
// (cp.async.bulk.tensor.2d.global.shared::cta): for that half of the map the load/store pipe sees 16 conflict-free STS
// instead of 16 four-wavefront STG per thread, and the row-strided global writes are generated by the copy engine while
// the CTA goes on with the product / inverse transform. The mirrored half stays on STG: the TMA unit wants the innermost
// coordinate of a store 16-byte aligned (measured on B200: an unaligned one raises "illegal instruction"), the main
// columns hx + 8 tile .. + 7 are, their mirrors hx - 8 tile - 7 .. hx - 8 tile start one float past a 16-byte boundary
// for every tile.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int x, int y) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(s) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
constexpr int TMA_BOX_ROWS = 256;
constexpr int COLS_CM = 1;    // complex-product form of the column pass (fft.cuh: cmulv)

struct SpecAcc {
    double total = 0, fx2 = 0, fy2 = 0, p2 = 0, all = 0, plogp = 0;
};

template <int NY>
__device__ __forceinline__ void spec_accumulate(SpecAcc& s, float P, int ky, int kx, int nx, double w, bool is_dc) {
    if (is_dc) return;
    const double p = (double)P;
    const int kys = ky <= NY / 2 ? ky : ky - NY;
    const double fy = (double)kys / (double)NY, fx = (double)kx / (double)nx;
    s.all += w * p;
    if (P > 0.f) s.plogp += w * p * (double)logf(P);
    if (fx * fx + fy * fy <= 0.25) {
        s.total += w * p;
        s.fx2 += w * fx * fx * p;
        s.fy2 += w * fy * fy * p;
        s.p2 += w * p * p;
    }
}

__device__ __forceinline__ float2 whiten_if(float2 G, int whiten, float eps) {
    if (whiten) {
        // 1 / (|G| + eps) from MUFU.RSQ and MUFU.RCP (a few ulp; the reference divides in float32 as well)
        const float s2 = fmaf(G.x, G.x, G.y * G.y);
        float rs, inv;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(s2, 1e-37f)));   // |G| = s2 * rsqrt(s2); 0 stays 0
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(s2, rs, eps)));
        G.x *= inv;
        G.y *= inv;
    }
    return G;
}

// Thread (j, c) owns the 16 elements ky = j + s*T (T = NY/16) of column c of the CTA's CW columns, in every phase:
// the loads, the forward transform (register to register, fft.cuh v2 core), the epilogues, the inverse transforms
// and the stores. Shared memory carries only the two inner exchanges of each transform, the |F|^2 copy that
// feeds the second inverse transform, and (tile 0) the mirror exchange that unpacks the DC / Nyquist column.
// CW (4 or 8) columns per CTA out of a layout tile of TC = 8: CW = 4 halves the footprint so that two CTAs
// share an SM and one's loads overlap the other's butterflies.
// The body is instantiated twice: TILE0 = true for the one CTA per frame that owns the packed DC / Nyquist column
// (thread-dependent unpacking branches at every element), false for the other tiles, whose code then carries no
// branch at all inside the element loops.
template <int NY, int CW, bool SPEC, bool AC, bool PC, bool TILE0>
__device__ __forceinline__ void cols_body(const ColsArgs& a, float2* sm, double* red, const CUtensorMap* tmap) {
    constexpr int T = NY / 16;
    constexpr int NT = T * CW;
    constexpr int PL = padded_len(NY);
    constexpr int GS = T * TC;                        // global distance (float2) between a thread's consecutive elements
    float2* A = sm;                                   // [PL][CW] exchange buffer
    float* Bp = reinterpret_cast<float*>(sm + PL * CW);   // [NY][CW] |F|^2 (AC)
    float* Pns = Bp + NY * CW;                        // [NY] |F_nyquist|^2 (AC, tile 0)

    const int tid = threadIdx.x, c = tid % CW, j = tid / CW;
    const int tile = blockIdx.x, ntiles = gridDim.x;
    const int64_t t = blockIdx.y;
    const int nx = a.nx, hx = nx / 2, kx = tile * CW + c;
    const bool nyq_owner = TILE0 && c == 0;           // TILE0: the CTA owns column 0
    // |F|^2 copy for the packed inverse transforms: the columns c and c + CW/2 that share a transform are adjacent
    // ([ky][c % (CW/2)][c / (CW/2)]), so that the transform's threads fetch a pair with one conflict-free 8-byte load
    const int bpw = j * CW + ((c % (CW / 2)) << 1) + c / (CW / 2);
    // element s of this thread lives at g0 + s*GS inside a frame's blocked half spectrum
    const size_t g0 = (size_t)t * NY * hx + (size_t)(kx / TC) * NY * TC + (size_t)j * TC + (kx % TC);

    float2 x[16];
    {
        const float2* Hin = a.H + g0;
#pragma unroll
        for (int s = 0; s < 16; ++s) x[s] = ld_inter(Hin + s * GS, a.keep);
    }
    // CTAs start in linear order and a tile is one contiguous NY*TC run of the blocked intermediate, so the tile of the
    // CTA that will take this SM next (pf_dist CTAs ahead) is known now: one 128-byte line per thread is pulled into L2
    // while this CTA works, and that CTA's loads wait for an L2 hit instead of for HBM.
    // (CW < TC: TC / CW CTAs share a layout tile and each pulls its share of the lines.)
    if (a.pf_dist > 0) {
        constexpr int R = TC / CW;
        const size_t next = (size_t)t * ntiles + tile + (size_t)a.pf_dist;
        if (next < (size_t)gridDim.y * ntiles)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.H + (next / R) * ((size_t)NY * TC) + ((next % R) * NT + (size_t)tid) * 16));
    }
    fft_regs<NY, -1, CW, 0, COLS_CM>(x, j, A + c, a.tw);
    // (the transform's barriers lie between every thread's tile loads and this point) the tile is dead now: drop its lines
    // from L2 instead of letting them be written back, one 128-byte line per thread
    if (B4D_KEEP_POLICY && CW == TC && (a.keep & 4))
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.H + (size_t)t * NY * hx + (size_t)tile * NY * TC + (size_t)tid * 16) : "memory");
    // TMA path of the PSD map (every tile but tile 0, whose packed DC / Nyquist column has its own rules): the exchange
    // buffer becomes the staging area of the main tile Bs[ky'][c], indexed by OUTPUT row ky' = (ky + NY/2) % NY
    const bool tma = !TILE0 && CW == TC && NY >= TMA_BOX_ROWS && a.psd_tma && a.psd_out;
    float* Bs = reinterpret_cast<float*>(A);
    if (tma) __syncthreads();                         // (other threads may still be reading the last exchange)

    // ---- tile 0: column 0 carries C = F[:,0] + i F[:,nx/2]; publish it so that its owners can read C[-ky]
    float2* A0 = A;                                   // natural order, [NY]
    if (TILE0) {
        __syncthreads();
        if (c == 0) {
#pragma unroll
            for (int s = 0; s < 16; ++s) A0[j + s * T] = x[s];
        }
        __syncthreads();
    }

    // ---- epilogue, two sweeps over the thread's own 16 elements so that the output pointers of the first and the
    //      reference values of the second never compete for the 64 registers a thread has:
    //      (1) plain outputs, partial sums, |F|^2 for the autocorrelation branch; (2) product with the reference.
    SpecAcc sp;
    float acsum = 0.f;                                // 16 terms in fp32, tiles and frames are summed in fp64
    // column 0 of tile 0: F0[ky] = (C + conj(Cm))/2, Fn[ky] = (C - conj(Cm))/(2i) with Cm = C[-ky]
    auto unpack0 = [&](int s, float2& F, float2& Fn, float& dcshift) {
        const int ky = j + s * T;
        const float2 Cm = A0[(NY - ky) & (NY - 1)];
        Fn = make_float2(0.5f * (F.y + Cm.y), -0.5f * (F.x - Cm.x));
        F = make_float2(0.5f * (F.x + Cm.x), 0.5f * (F.y - Cm.y));
        if (ky == 0) {
            dcshift = F.x;
            if (a.pilot) F.x += (float)((double)nx * (double)NY * (double)__ldg(a.pilot + t));
            if (a.zero_dc) F = make_float2(0.f, 0.f);
        }
    };
    {
        // shifted output rows: main (ky + NY/2) mod NY = j + T ((s + 8) & 15); mirror (NY/2 - ky) mod NY = (NY - j) +
        // T (((8 - s) & 15) - 16), except s = 8 where it is (NY - j) mod NY. Both are "thread base + constant * T rows".
        const size_t rstride = (size_t)T * nx;        // elements between output rows T apart
        float* psd = a.psd_out ? a.psd_out + (size_t)t * NY * nx + (size_t)j * nx + (kx + hx) : nullptr;
        float* psdm = a.psd_out ? a.psd_out + (size_t)t * NY * nx + (size_t)(NY - j) * nx + (hx - kx) : nullptr;
        // the complex / conjugate spectra are outputs of the plain forward launches only (fft2d, reference spectra): with an
        // inverse branch compiled in they are never asked for, and their per-element tests leave the loops
        constexpr bool PLAIN = !AC && !PC;
        float2* cpl = (PLAIN && a.cplx_out) ? a.cplx_out + (size_t)t * NY * nx + (size_t)j * nx + (kx + hx) : nullptr;
        float2* cplm = (PLAIN && a.cplx_out) ? a.cplx_out + (size_t)t * NY * nx + (size_t)(NY - j) * nx + (hx - kx) : nullptr;
        float2* cj = (PLAIN && a.conj_out) ? a.conj_out + g0 : nullptr;
        float2* cjn = (PLAIN && a.conj_nyq_out) ? a.conj_nyq_out + (size_t)t * NY : nullptr;
        const bool mirror = !TILE0 || kx >= 1;
        const float wgt = mirror ? 2.f : 1.f;
        const float ps = a.psd_scale;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int ky = j + s * T;
            const int um = (s + 8) & 15;                                                  // relative to row j
            const int ur = s == 8 ? (j == 0 ? -16 : 0) : ((8 - s) & 15) - 16;             // relative to row NY - j
            float2 F = x[s];
            float2 Fn = make_float2(0.f, 0.f);
            float dcshift = 0.f;
            if (nyq_owner) unpack0(s, F, Fn, dcshift);
            const float P = F.x * F.x + F.y * F.y;
            if (psd) {
                if (tma) Bs[((ky + NY / 2) & (NY - 1)) * CW + c] = P * ps;
                else __stcs(psd + (ptrdiff_t)um * (ptrdiff_t)rstride, P * ps);
                if (mirror) __stcs(psdm + (ptrdiff_t)ur * (ptrdiff_t)rstride, P * ps);
            }
            if (cpl) {
                __stcs(cpl + (ptrdiff_t)um * (ptrdiff_t)rstride, F);
                if (mirror) __stcs(cplm + (ptrdiff_t)ur * (ptrdiff_t)rstride, cconj(F));
            }
            if (cj) cj[s * GS] = cconj(F);
            if (SPEC) spec_accumulate<NY>(sp, P * ps, ky, kx, nx, (double)wgt, TILE0 && ky == 0 && kx == 0);
            float Pa = P, Pn = 0.f;
            if (nyq_owner) {                          // the Nyquist column kx = nx/2 lands in shifted column 0
                Pn = Fn.x * Fn.x + Fn.y * Fn.y;
                const size_t rs = (size_t)((ky + NY / 2) & (NY - 1)) * nx;
                if (a.psd_out) __stcs(a.psd_out + (size_t)t * NY * nx + rs, Pn * ps);
                if (PLAIN && a.cplx_out) __stcs(a.cplx_out + (size_t)t * NY * nx + rs, Fn);
                if (cjn) cjn[ky] = cconj(Fn);
                if (SPEC) spec_accumulate<NY>(sp, Pn * ps, ky, hx, nx, 1.0, false);
            }
            if (AC) {
                if (a.ac_zero_dc && nyq_owner && ky == 0) Pa = 0.f;
                acsum += fmaf(wgt, Pa, Pn);
                Bp[bpw + s * NT] = Pa;
                if (nyq_owner) Pns[ky] = Pn;
            }
        }
    }
    // The first half of the thread's reference-spectrum values is requested now: the barriers of the TMA hand-over and of
    // the block reductions below are spent waiting anyway, and one 1024-thread CTA per SM has no other warps to hide an L2
    // round trip behind. (The second half follows the first product sweep: sixteen values in flight next to x[] would spill.)
    float2 Rv0[8];
    const float2* R = nullptr;
    if (PC) {
        R = a.R + (size_t)t * a.r_stride + (g0 - (size_t)t * NY * hx);
#pragma unroll
        for (int i = 0; i < 8; ++i) Rv0[i] = __ldg(R + i * GS);
    }
    if (tma) {
        fence_async_smem();                           // the staged tiles become visible to the async proxy
        __syncthreads();
        if (tid == 0) {
            const int x_main = hx + tile * CW;
            const int y0 = (int)t * NY;
#pragma unroll
            for (int b = 0; b < NY / TMA_BOX_ROWS; ++b) tma_store_2d(tmap, Bs + b * TMA_BOX_ROWS * CW, x_main, y0 + b * TMA_BOX_ROWS);
            tma_commit();
        }
    }
    // ---- block reductions of the scalar partials ------------------------------------------------
    if (SPEC || AC) {
        double v[NSP + 1] = {sp.total, sp.fx2, sp.fy2, sp.p2, sp.all, sp.plogp, (double)acsum};
        const int warp = tid >> 5, lane = tid & 31, nw = (NT + 31) / 32;
#pragma unroll
        for (int i = SPEC ? 0 : NSP; i < NSP + 1; ++i) {
            if (i == NSP && !AC) break;
            double r = warp_sum(v[i]);
            __syncthreads();
            if (lane == 0) red[warp] = r;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < nw; ++w) tot += red[w];
                if (i < NSP) {
                    if (a.spec_partials) a.spec_partials[((size_t)t * ntiles + tile) * NSP + i] = tot;
                } else {
                    a.ac_partials[(size_t)t * ntiles + tile] = tot;
                }
            }
        }
    }
    if (PC) {
        float inv_s = 1.f;
        if (a.fr) inv_s = (float)(1.0 / (sqrt(a.fr[(size_t)t * a.fr_stride + B4D_FR_M2]) + (double)a.eps));
        const float2* Rn = a.Rnyq + (size_t)t * a.rnyq_stride;
        // the reference spectrum is fetched eight elements at a time (a compiler barrier keeps the second half from
        // being hoisted over the first, which would spill)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 Rv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) Rv[i] = h == 0 ? Rv0[i] : __ldg(R + (8 + i) * GS);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int s = 8 * h + i;
                const int ky = j + s * T;
                float2 F = x[s];
                float2 Fn = make_float2(0.f, 0.f);
                float dcshift = 0.f;
                if (nyq_owner) {
                    unpack0(s, F, Fn, dcshift);
                    if (ky == 0 && a.fr) {
                        // DC of the mean-removed frame: sum(x - K) + n (K - mean), formed without cancellation
                        const double K = a.pilot ? (double)__ldg(a.pilot + t) : 0.0;
                        const double n = (double)nx * (double)NY;
                        F.x = (float)((double)dcshift + n * (K - a.fr[(size_t)t * a.fr_stride + B4D_FR_MEAN]));
                    }
                }
                F.x *= inv_s; F.y *= inv_s;
                float2 G = whiten_if(cmulv<COLS_CM>(F, Rv[i]), a.whiten, a.eps);
                if (nyq_owner) {                      // pack: column 0 <- G[:,0] + i G[:,nx/2]
                    Fn.x *= inv_s; Fn.y *= inv_s;
                    const float2 Gn = whiten_if(cmul(Fn, __ldg(Rn + ky)), a.whiten, a.eps);
                    G = make_float2(G.x - Gn.y, G.y + Gn.x);
                }
                x[s] = G;
            }
            asm volatile("" ::: "memory");
        }
    }

    // the copy engine must have read the staged tiles before the exchange buffer is written again (or the CTA leaves):
    // the issuing thread waits here, everybody else meets it at the next barrier (entry of the inverse transform)
    if (tma && tid == 0) tma_wait_read_all();
    if (!PC && !AC) return;

    // ---- inverse along y of the product ---------------------------------------------------------
    if (PC) {
        fft_regs<NY, +1, CW, 0, COLS_CM>(x, j, A + c, a.tw);
        // rows 2p and 2p + 1 of a column sit next to each other in this intermediate ([kx/8][y/2][kx%8][y%2]): the row
        // pass, which transforms such a pair in one complex transform, fetches both with one 16-byte load and every
        // line it touches whole. (ky = j + s T with T even: the pair interleave moves the thread's base only.)
        float2* o = a.i2_pc + (size_t)t * NY * hx + (size_t)(kx / TC) * NY * TC + ((size_t)(j >> 1) * TC + (kx % TC)) * 2 + (j & 1);
#pragma unroll
        for (int s = 0; s < 16; ++s) st_inter(o + s * GS, x[s], a.keep);
    }
    if (AC) {
        // Inverse along y of |F|^2, kept in Bp. Its columns are real, so two of them (c2 and c2 + CW/2) share one complex
        // transform, z = IFFT_y(P_a) + i IFFT_y(P_b); each part is Hermitian in y and rows_inv_ac_kernel separates them
        // from rows y and -y. Only the first half of the CTA (whole warps) carries these CW/2 transforms; in tile 0 four
        // more warps transform the Nyquist column. The packed intermediate has nx/4 columns: tile-major, column
        // tile * CW/2 + c2 in the usual blocked layout.
        constexpr int CH = CW / 2, NTH = T * CH;
        asm volatile("" ::: "memory");                // the loads below must not be hoisted over the first inverse
        __syncthreads();                              // everybody is done with A
        if (tid < NTH) {
            const int c2 = tid % CH, j2 = tid / CH;
#pragma unroll
            for (int s = 0; s < 16; ++s) x[s] = *reinterpret_cast<const float2*>(Bp + j2 * CW + 2 * c2 + s * NT);
            fft_regs<NY, +1, CH, 2, COLS_CM>(x, j2, A + c2, a.tw);
            const int pc = tile * CH + c2;
            // (row pairs interleaved like i2_pc: [pc/8][y/2][pc%8][y%2])
            float2* o = a.i2_ac + (size_t)t * NY * (hx / 2) + (size_t)(pc / TC) * NY * TC + ((size_t)(j2 >> 1) * TC + (pc % TC)) * 2 + (j2 & 1);
#pragma unroll
            for (int s = 0; s < 16; ++s) st_inter(o + s * GS, x[s], a.keep);
        } else if (TILE0 && tid < NTH + (T < 32 ? 32 : T)) {
            // (T < 32: the group is padded to a warp, the extra threads repeat the work of the first T)
            const int j3 = (tid - NTH) % T;
#pragma unroll
            for (int s = 0; s < 16; ++s) x[s] = make_float2(Pns[j3 + s * T], 0.f);
            fft_regs<NY, +1, 1, 1, COLS_CM>(x, j3, A + PL * CH, a.tw, 1);
            if (tid - NTH < T) {
                float2* o = a.i2_ac_nyq + (size_t)t * NY + j3;
#pragma unroll
                for (int s = 0; s < 16; ++s) o[s * T] = x[s];
            }
        }
    }
}

template <int NY, int CW, bool SPEC, bool AC, bool PC>
__global__ void __launch_bounds__(NY / 16 * CW, 1024 / (NY / 16 * CW)) cols_kernel(ColsArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) float2 sm[];
    __shared__ double red[32];
    if (blockIdx.x == 0) cols_body<NY, CW, SPEC, AC, PC, true>(a, sm, red, &tmap);
    else cols_body<NY, CW, SPEC, AC, PC, false>(a, sm, red, &tmap);
}

// =================================================================================================
// K3: rows inverse (two half-spectrum rows -> two real rows)
// =================================================================================================
struct ArgBest {
    float v;
    unsigned idx;
};
__device__ __forceinline__ void best_update(ArgBest& b, float v, unsigned idx) {
    if (v > b.v || (v == b.v && idx < b.idx)) { b.v = v; b.idx = idx; }
}

struct RowsInvArgs {
    const float2* Ia;       // blocked intermediate (per frame ny*nx/2), row pairs interleaved: [kx/8][y/2][kx%8][y%2]
    const float2* tw;
    int ny;
    float* outA;            // shifted real output: (T, ny, nx) full map, or compact rows (mag_mode 1); nullable
    int kindA;              // 0: signed value * scale; 1: |value| * scale
    const double* normA;    // per-frame partials (T, n_normA) whose sum is the peak (nullable -> 1)
    int n_normA;
    double norm_mult;       // value = raw * norm_mult / sum(normA)
    double scaleA;          // extra factor (e.g. 1/(nx*ny))
    ArgBest* bestA;         // (T, nblk) argmax partials (nullable)
    // row-block selection: blockIdx.x -> row block of the frame (identity when both are null)
    int nblk;               // row blocks per frame (stride of the argmax partials)
    const int* blk_map;     // shared by all frames (gridDim.x entries)
    const int* blk_map_pf;  // per frame (T, gridDim.x)
    // destination of the map:
    //   0 full map; 1 compact rows: (T, gridDim.x * rows_per_cta, nx) in launch order; 2 none, census of the |.| values
    //   against the frame's bracket instead (fused median, select.cuh)
    int mag_mode;
    const SelFast* sel;     // per frame bracket
    unsigned* cand;         // (T, regions * FM_REGION + FM_SAMPLE_CAP) region store
    unsigned* cnt3;         // (T, regions, 3)
    unsigned* bhist;        // (T, SEL_BINS)
    int regions;            // CTA regions per frame = row blocks
    int pf_dist;            // L2 prefetch distance in CTAs (0 = off), set by the launcher
    int keep;               // cache policy of the intermediate (see st_inter)
};

// One complex inverse transform yields rows 2p and 2p + 1 of the map. Two thread mappings: the gather uses lanes
// (c = lane & 7, fl = lane >> 3): 8 adjacent kx of 4 transforms, i.e. whole 64-byte runs of the blocked intermediate
// (16-byte loads: both rows at once would need them adjacent, they are 64 B apart). The transform and everything
// after it use the natural mapping (f = tid / TPF, j = tid % TPF): the v2 core leaves x-position j + TPF*s in slot s,
// so the two real output rows are stored straight from registers, 128 contiguous bytes per warp and slot.
// MODE = a.mag_mode and ABS = a.kindA are compile-time copies of the two switches the epilogue turns on.
template <int NX, int MODE, bool ABS>
__global__ void __launch_bounds__(512, 2) rows_inv_kernel(RowsInvArgs a) {
    constexpr int TPF = NX / 16;
    constexpr int WPG = TPF / 8;          // warps per group of 4 transforms (gather mapping)
    constexpr int GPC = 16 / WPG;         // groups per CTA (512 threads)
    constexpr int FPC = 4 * GPC;
    constexpr int FS = padded_len(NX) + 8;
    constexpr int HX = NX / 2;
    constexpr int RPC = 2 * FPC;          // rows per CTA
    extern __shared__ float2 sm[];
    __shared__ float s_scale;
    __shared__ float s_max[16];
    __shared__ unsigned s_idx;
    __shared__ unsigned s_cen[2];                     // fused median: values below the bracket, valid values of the CTA
    __shared__ unsigned s_wcnt[16];                   // candidates per warp
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t t = blockIdx.y;
    const int NY = a.ny;
    const int blk = a.blk_map_pf ? a.blk_map_pf[(size_t)t * gridDim.x + blockIdx.x] : (a.blk_map ? a.blk_map[blockIdx.x] : (int)blockIdx.x);
    const int y0 = blk * RPC;
    const float2* Ia = a.Ia + (size_t)t * NY * HX;
    unsigned Lk = 0u, Uk = 0xfffffffeu;               // fused median: the frame's bracket, fetched early
    if (MODE == 2) { Lk = a.sel[t].L[0]; Uk = a.sel[t].U[0]; }
    if (tid == 0) s_idx = 0xffffffffu;
    if (MODE == 2 && tid < 2) s_cen[tid] = 0u;

    // ---- inputs of the transform, natural mapping (f = tid / TPF, j = tid % TPF): thread j wants Z[j + m TPF], m = 0..15.
    //      Z[k] = Ga[k] + i Gb[k] for k < NX/2 and Z[NX - k] = conj(Ga[k]) + i conj(Gb[k]). The thread fetches the pairs
    //      k = j + m TPF, m = 0..7: their Z[k] ARE its slots 0..7 and stay in registers; their Z[NX - k] are slots 8..15 of
    //      thread TPF - j of the same transform and cross through shared memory -- half of what a gather of all sixteen
    //      slots would move, and only the transform's own threads meet at the barrier.
    const int f = tid / TPF, j = tid % TPF;
    float2* z = sm + f * FS;
    float2 x[16];
    {
        const int ya = y0 + 2 * f;
        float2 ga[8], gb[8];
        {
            // k = j + m TPF lives in tile j/8 + m TPF/8: a fixed stride of (TPF/8) tiles between the thread's loads; rows
            // ya and ya + 1 of a column are adjacent (pair-interleaved layout written by the column pass): one 16-byte load
            const size_t tstride = (size_t)(TPF / TC) * NY * TC;
            const float2* pa = Ia + (((size_t)(j / TC) * (NY / 2) + (ya >> 1)) * TC + (j % TC)) * 2;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float4 g = __ldcs(reinterpret_cast<const float4*>(pa + m * tstride));
                ga[m] = make_float2(g.x, g.y);
                gb[m] = make_float2(g.z, g.w);
            }
        }
        // the rows of the CTA that takes this SM next (pf_dist CTAs ahead in launch order): RPC rows x 64 bytes in each of
        // the nx/16 column tiles, 512 lines in all, one per thread, pulled into L2 while this CTA works
        if (a.pf_dist > 0) {
            const unsigned nb = blockIdx.x + (unsigned)a.pf_dist;
            const int64_t tn = t + nb / gridDim.x;
            const unsigned bn = nb % gridDim.x;
            if (tn < (int64_t)gridDim.y) {
                const int blkn = a.blk_map_pf ? a.blk_map_pf[(size_t)tn * gridDim.x + bn] : (a.blk_map ? a.blk_map[bn] : (int)bn);
                const int q = tid / (RPC / 2), l = tid % (RPC / 2);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Ia + (size_t)tn * NY * HX + ((size_t)q * NY + (size_t)blkn * RPC) * TC + (size_t)l * 16));
            }
        }
        float sA = (float)a.scaleA;
        if (a.normA) {
            if (warp == 0) {
                // fixed-order reduction of the per-tile partials: lane-strided chunks, then a shuffle tree
                double part = 0.0;
                for (int i = lane; i < a.n_normA; i += 32) part += a.normA[(size_t)t * a.n_normA + i];
                part = warp_sum(part);
                if (lane == 0) s_scale = (float)(part > 0.0 ? a.norm_mult / part : a.scaleA);
            }
            __syncthreads();
            sA = s_scale;
        }
        const float2 sA2 = make_float2(sA, sA);
        // With TPF a multiple of 16 the padded index of Z[NX - k] is a per-thread base plus a compile-time multiple of m:
        // pad16(NX - j - m TPF) = padded_len(NX) - j - ceil(j / 16) - m (TPF + TPF / 16).
        constexpr bool LIN = (TPF % 16) == 0;
        constexpr int PS = TPF + TPF / 16;
        float2* zm = z + (padded_len(NX) - j - ((j + 15) >> 4));
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float2 g1 = __fmul2_rn(ga[m], sA2), g2 = __fmul2_rn(gb[m], sA2);
            const float2 A = __fadd2_rn(g1, make_float2(-g2.y, g2.x));
            const float2 B = __fadd2_rn(make_float2(g1.x, -g1.y), make_float2(g2.y, g2.x));
            if (m == 0) {
                // pair 0 of thread 0 is the packed slot (DC, Nyquist), both real: Z[0] and Z[NX/2]
                x[0] = j == 0 ? make_float2(g1.x, g2.x) : A;
                if (j == 0) z[pad16(HX)] = make_float2(g1.y, g2.y);
                else if (LIN) zm[0] = B;
                else z[pad16(NX - j)] = B;
            } else {
                x[m] = A;
                if (LIN) zm[-PS * m] = B;
                else z[pad16(NX - j - m * TPF)] = B;
            }
        }
        fft_sync<NX, (NX >= 1024), 1>(f);
#pragma unroll
        for (int m = 8; m < 16; ++m) x[m] = z[pad16(j + m * TPF)];
    }
    fft_regs<NX, +1, 1, (NX >= 1024)>(x, j, z, a.tw, f);

    // ---- stores: real part -> row ya, imaginary part -> row ya + 1. Slot s holds x-position j + TPF s, i.e. shifted
    //      column j + TPF ((s + 8) & 15): a per-thread pointer plus a compile-time offset.
    const int ra_ = y0 + 2 * f, rb_ = ra_ + 1;
    const unsigned rowA = (unsigned)((ra_ + NY / 2) & (NY - 1)) * (unsigned)NX;
    const unsigned rowB = (unsigned)((rb_ + NY / 2) & (NY - 1)) * (unsigned)NX;
    float* pA = nullptr;
    float* pB = nullptr;
    if (MODE == 0 && a.outA) { pA = a.outA + (size_t)t * NY * NX + rowA + j; pB = a.outA + (size_t)t * NY * NX + rowB + j; }
    if (MODE == 1) {
        // compact rows in launch order (sample rows of the fused median, 3x3 window of the peak)
        pA = a.outA + ((size_t)t * gridDim.x * RPC + (size_t)blockIdx.x * RPC + 2 * f) * NX + j;
        pB = pA + NX;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        const int s = (o + 8) & 15;
        float va = x[s].x, vb = x[s].y;
        if (ABS) { va = fabsf(va); vb = fabsf(vb); }
        x[s] = make_float2(va, vb);
        if (MODE != 2 && pA) { pA[TPF * o] = va; pB[TPF * o] = vb; }
        mx = fmaxf(mx, fmaxf(va, vb));
    }
    // ---- argmax partial: block maximum first, then the smallest linear index that holds it (first occurrence in
    //      row-major order of the shifted map wins ties)
    if (a.bestA) {
        float wm = mx;
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, q));
        if (lane == 0) s_max[warp] = wm;
        __syncthreads();
        float bm = s_max[0];
#pragma unroll
        for (int w = 1; w < 16; ++w) bm = fmaxf(bm, s_max[w]);
        if (mx == bm) {
            unsigned best = 0xffffffffu;
#pragma unroll
            for (int o = 15; o >= 0; --o) {
                const int s = (o + 8) & 15;
                if (x[s].y == bm) best = rowB + (unsigned)(j + TPF * o);
            }
#pragma unroll
            for (int o = 15; o >= 0; --o) {
                const int s = (o + 8) & 15;
                if (x[s].x == bm) best = min(best, rowA + (unsigned)(j + TPF * o));
            }
            atomicMin(&s_idx, best);
        }
        __syncthreads();
        if (tid == 0) {
            ArgBest b = {bm, s_idx == 0xffffffffu ? 0u : s_idx};     // (an all-NaN block has no maximum)
            a.bestA[(size_t)t * a.nblk + blk] = b;
        }
    }
    if (MODE == 2) {
        // fused median: census of the |.| values against the frame's bracket, straight from the registers. The exchange
        // buffer is dead (the outputs sit in registers) and becomes the stage: 32 slot rows of 512 words, a column per
        // thread, so that every value of a thread has a slot.
        static_assert(FPC * FS * sizeof(float2) >= 32 * 512 * sizeof(unsigned), "stage does not fit the exchange buffer");
        const unsigned W = min(Uk, 0x7fffffffu) - Lk;
        __syncthreads();                              // the other transforms of the CTA may still be reading their exchanges
        const unsigned* stage = reinterpret_cast<const unsigned*>(sm);
        const unsigned ad0 = (unsigned)__cvta_generic_to_shared(stage + tid);
        unsigned ad = ad0, below = 0;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            census_stage_value<2048>(x[s].x, Lk, W, below, ad);
            census_stage_value<2048>(x[s].y, Lk, W, below, ad);
        }
        // A NaN anywhere in the frame reaches every output of its inverse transform, so one value per lane and row tells
        // whether the warp's values are valid.
        const unsigned nvalid = (unsigned)(__popc(__ballot_sync(0xffffffffu, x[0].x == x[0].x)) +
                                           __popc(__ballot_sync(0xffffffffu, x[0].y == x[0].y))) * 16u;
        below = __reduce_add_sync(0xffffffffu, below);
        if (lane == 0) { atomicAdd(&s_cen[0], below); atomicAdd(&s_cen[1], nvalid); }
        unsigned* store = a.cand + (size_t)t * ((size_t)a.regions * FM_REGION + FM_SAMPLE_CAP);
        // (the flush's barrier lies between the atomics above and thread 0 reading the totals)
        census_stage_flush<512>(stage, (ad - ad0) >> 11, s_wcnt, 0u, 0u, Lk, bracket_shift(Lk, Uk), store + (size_t)blk * FM_REGION,
                                a.cnt3 + ((size_t)t * a.regions + blk) * 3, a.bhist + (size_t)t * SEL_BINS);
        if (tid == 0) {
            unsigned* c3 = a.cnt3 + ((size_t)t * a.regions + blk) * 3;
            c3[1] = s_cen[0]; c3[2] = s_cen[1];
        }
    }
}

// =================================================================================================
// K3b: rows inverse of the packed autocorrelation intermediate (fused pipeline)
// =================================================================================================
// cols_kernel<.., AC, PC> leaves z = IFFT_y(P_a) + i IFFT_y(P_b) for the column pairs (a, b = a + cw/2) of every tile
// of cw columns. P is real, so g_a(y) = IFFT_y(P_a) is Hermitian in y and
//     g_a(y) = (z(y) + conj z(-y)) / 2,      g_b(y) = (z(y) - conj z(-y)) / (2i).
// The autocorrelation is point symmetric, ac(-y, -x) = ac(y, x): only the rows y = 0 .. ny/2 are transformed (two rows
// per complex transform, as in rows_inv_kernel) and every row is stored twice, as it is and mirrored.
struct RowsInvAcArgs {
    const float2* Iz;       // packed blocked intermediate (T, nx/4/TC, ny/2, TC, 2): row pairs interleaved
    const float2* Inyq;     // (T, ny) inverse-along-y of the Nyquist column
    const float2* tw;
    int ny;
    int ch_log2;            // log2(cw / 2): pairing distance of the column pass
    float* out;             // (T, ny, nx) shifted map
    const double* norm;     // per-frame partials whose sum is the peak (nullable -> scale)
    int n_norm;
    double norm_mult;
    double scale;
    ArgBest* best;          // (T, gridDim.x) argmax partials (nullable)
    int keep;               // cache policy of the intermediate (see st_inter)
};

template <int NX>
__global__ void __launch_bounds__(512, 2) rows_inv_ac_kernel(RowsInvAcArgs a) {
    constexpr int TPF = NX / 16;
    constexpr int WPG = TPF / 8;
    constexpr int GPC = 16 / WPG;
    constexpr int FPC = 4 * GPC;
    constexpr int FS = padded_len(NX) + 4;            // transforms 8 words apart in bank space: a lane group of the gather writes
                                                      // elements {0..3, 8..11} (+4 for the partner column) of its transform, so the four
                                                      // transforms of a warp land on disjoint banks, half-warp by half-warp
    constexpr int HX = NX / 2;
    extern __shared__ float2 sm[];
    __shared__ float s_scale;
    __shared__ float s_max[16];
    __shared__ unsigned s_idx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t t = blockIdx.y;
    const int NY = a.ny;
    const int y0 = blockIdx.x * 2 * FPC;
    const float2* Iz = a.Iz + (size_t)t * NY * (HX / 2);
    if (tid == 0) s_idx = 0xffffffffu;

    // ---- gather: each thread fetches z(ya), z(-ya), z(yb), z(-yb) of 4 packed columns, i.e. 8 values of k
    {
        const int c = lane & 7, fl = lane >> 3, jt = warp % WPG, grp = warp / WPG;
        const int f = grp * 4 + fl, jg = jt * 8 + c;
        const int ya = y0 + 2 * f, yb = ya + 1;
        const int ma = (NY - ya) & (NY - 1), mb = NY - yb;
        float2 za[4], zam[4], zb[4], zbm[4];
        {
            // element (y, c) of a tile sits at ((y / 2) * 8 + c) * 2 + y % 2 (row pairs interleaved by the column pass): rows
            // ya (even) and yb = ya + 1 come with one 16-byte load; their mirrors -ya (even) and -yb (odd) belong to two pairs
            const size_t tstride = (size_t)(TPF / TC) * NY * TC;
            const float2* p = Iz + (size_t)(jg / TC) * NY * TC + (size_t)(jg % TC) * 2;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float4 g = __ldcs(reinterpret_cast<const float4*>(p + m * tstride + (size_t)(ya >> 1) * (2 * TC)));
                za[m] = make_float2(g.x, g.y);
                zb[m] = make_float2(g.z, g.w);
                zam[m] = ld_inter(p + m * tstride + (size_t)(ma >> 1) * (2 * TC), a.keep);
                zbm[m] = ld_inter(p + m * tstride + (size_t)(mb >> 1) * (2 * TC) + 1, a.keep);
            }
        }
        float nya = 0.f, nyb = 0.f;
        if (jg == 0) {
            nya = a.Inyq[(size_t)t * NY + ya].x;
            nyb = a.Inyq[(size_t)t * NY + yb].x;
        }
        if (warp == 0) {
            double sc = a.scale;
            if (a.norm) {
                double part = 0.0;
                for (int i = lane; i < a.n_norm; i += 32) part += a.norm[(size_t)t * a.n_norm + i];
                part = warp_sum(part);
                sc = part > 0.0 ? a.norm_mult / part : a.scale;
            }
            if (lane == 0) s_scale = (float)sc;
        }
        __syncthreads();
        const float sA = s_scale, sH = 0.5f * sA;
        const float2 sH2 = make_float2(sH, sH);
        float2* z = sm + f * FS;
        // Column pc = jg + m TPF of the packed intermediate carries the spectrum columns ka = ka0 + 2 TPF m and kb = ka + cw/2
        // (ka0 = jg with a zero bit inserted at position ch_log2). 2 TPF is a multiple of 16, so the four padded indices
        // pad16(ka), pad16(kb), pad16(NX - ka), pad16(NX - kb) are per-thread bases plus compile-time multiples of m:
        // kb = ka + cw/2 never crosses a 16-boundary, NX - kb = (NX - ka) - cw/2 crosses one iff (NX - ka) % 16 < cw/2.
        static_assert(TPF % 8 == 0, "2 TPF must be a multiple of 16");
        constexpr int PS2 = 2 * TPF + 2 * TPF / 16;
        const int chm = (1 << a.ch_log2) - 1, c1 = chm + 1;
        const int ka0 = ((jg >> a.ch_log2) << (a.ch_log2 + 1)) + (jg & chm);
        float2* zka = z + pad16(ka0);
        float2* zkb = zka + c1;
        float2* zma = z + (padded_len(NX) - ka0 - ((ka0 + 15) >> 4));
        float2* zmb = zma - c1 - ((((16 - (ka0 & 15)) & 15) < c1) ? 1 : 0);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            // rows ya (g1) and yb (g2) of the two columns: g_a = (z + conj zm) / 2, g_b = (z - conj zm) / (2i)
            const float2 g1a = __fmul2_rn(__fadd2_rn(za[m], make_float2(zam[m].x, -zam[m].y)), sH2);
            const float2 g1b = __fmul2_rn(__fadd2_rn(make_float2(za[m].y, -za[m].x), make_float2(zam[m].y, zam[m].x)), sH2);
            const float2 g2a = __fmul2_rn(__fadd2_rn(zb[m], make_float2(zbm[m].x, -zbm[m].y)), sH2);
            const float2 g2b = __fmul2_rn(__fadd2_rn(make_float2(zb[m].y, -zb[m].x), make_float2(zbm[m].y, zbm[m].x)), sH2);
            if (m == 0 && ka0 == 0) {
                z[pad16(0)] = make_float2(g1a.x, g2a.x);
                z[pad16(HX)] = make_float2(sA * nya, sA * nyb);
            } else {
                zka[PS2 * m] = __fadd2_rn(g1a, make_float2(-g2a.y, g2a.x));
                zma[-PS2 * m] = __fadd2_rn(make_float2(g1a.x, -g1a.y), make_float2(g2a.y, g2a.x));
            }
            zkb[PS2 * m] = __fadd2_rn(g1b, make_float2(-g2b.y, g2b.x));
            zmb[-PS2 * m] = __fadd2_rn(make_float2(g1b.x, -g1b.y), make_float2(g2b.y, g2b.x));
        }
        __syncthreads();
    }

    // ---- transform (natural mapping), outputs stay in registers
    const int f = tid / TPF, j = tid % TPF;
    float2* z = sm + f * FS;
    float2 x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) x[m] = z[pad16(j + m * TPF)];
    fft_regs<NX, +1, 1, (NX >= 1024)>(x, j, z, a.tw, f);

    // ---- stores: real part -> row ya, imaginary part -> row yb, each also mirrored through the centre of the shifted
    //      map: (r, c) -> ((ny - r) % ny, (nx - c) % nx). Rows 0 and ny/2 (shifted ny/2 and 0) are their own mirrors.
    const int ya = y0 + 2 * f, yb = ya + 1;
    const bool mainA = ya <= NY / 2, mirA = ya >= 1 && ya < NY / 2, mainB = yb <= NY / 2, mirB = yb < NY / 2;
    const unsigned rA = (unsigned)((ya + NY / 2) & (NY - 1)), rB = (unsigned)((yb + NY / 2) & (NY - 1));
    const unsigned rAm = (unsigned)((NY - rA) & (NY - 1)), rBm = (unsigned)((NY - rB) & (NY - 1));
    float* o = a.out + (size_t)t * NY * NX;
    float* pA = o + (size_t)rA * NX + j;
    float* pB = o + (size_t)rB * NX + j;
    float* pAm = o + (size_t)rAm * NX + (NX - j);     // slot o >= 1 lands at pAm[-TPF o]; slot 0 at column (NX - j) % NX
    float* pBm = o + (size_t)rBm * NX + (NX - j);
    const int wrap0 = j == 0 ? -NX : 0;
    float mx = -INFINITY;
#pragma unroll
    for (int oo = 0; oo < 16; ++oo) {
        const int s = (oo + 8) & 15;
        const float va = x[s].x, vb = x[s].y;
        const int mo = oo == 0 ? wrap0 : -TPF * oo;
        if (mainA) { __stcs(pA + TPF * oo, va); mx = fmaxf(mx, va); }          // (the map is written once: streaming stores)
        if (mirA) __stcs(pAm + mo, va);
        if (mainB) { __stcs(pB + TPF * oo, vb); mx = fmaxf(mx, vb); }
        if (mirB) __stcs(pBm + mo, vb);
    }
    // ---- argmax partial: block maximum first, then the smallest linear index (of either copy) that holds it
    if (a.best) {
        float wm = mx;
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, q));
        if (lane == 0) s_max[warp] = wm;
        __syncthreads();
        float bm = s_max[0];
#pragma unroll
        for (int w = 1; w < 16; ++w) bm = fmaxf(bm, s_max[w]);
        if (mx == bm) {
            unsigned best = 0xffffffffu;
#pragma unroll
            for (int oo = 0; oo < 16; ++oo) {
                const int s = (oo + 8) & 15;
                const unsigned col = (unsigned)(j + TPF * oo), colm = (unsigned)((NX - (int)col) & (NX - 1));
                if (mainA && x[s].x == bm) {
                    best = min(best, rA * (unsigned)NX + col);
                    if (mirA) best = min(best, rAm * (unsigned)NX + colm);
                }
                if (mainB && x[s].y == bm) {
                    best = min(best, rB * (unsigned)NX + col);
                    if (mirB) best = min(best, rBm * (unsigned)NX + colm);
                }
            }
            atomicMin(&s_idx, best);
        }
        __syncthreads();
        if (tid == 0) {
            ArgBest b = {bm, s_idx == 0xffffffffu ? 0u : s_idx};     // (an all-NaN block has no maximum)
            a.best[(size_t)t * gridDim.x + blockIdx.x] = b;
        }
    }
}

// =================================================================================================
// small kernels
// =================================================================================================
// mean and variance of a frame from the per-CTA sums of rows_fwd_kernel, into the frame-reduction table columns the
// column pass reads (B4D_FR_MEAN, B4D_FR_M2); fixed summation order
__global__ void __launch_bounds__(32) rows_moments_finalize_kernel(const double* __restrict__ mom, int nblk, const float* __restrict__ pilot,
                                                                   double npix, double* __restrict__ fr) {
    const int64_t t = blockIdx.x;
    const int lane = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int b = lane; b < nblk; b += 32) { s1 += mom[((size_t)t * nblk + b) * 2]; s2 += mom[((size_t)t * nblk + b) * 2 + 1]; }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
        const double a1 = s1 / npix, a2 = s2 / npix;
        double m2 = a2 - a1 * a1;
        if (m2 < 0) m2 = 0;
        double* o = fr + t * B4D_FR_NCOLS;
        for (int i = 0; i < B4D_FR_NCOLS; ++i) o[i] = nan("");
        o[B4D_FR_COUNT] = npix; o[B4D_FR_NPIX] = npix;
        o[B4D_FR_MEAN] = (double)pilot[t] + a1;
        o[B4D_FR_M2] = m2;
    }
}


// peak location from the argmax partials: out_idx[t] = shifted linear index, out_val[t] = value
__global__ void __launch_bounds__(128) argmax_reduce_kernel(const ArgBest* __restrict__ part, int n,
                                                            unsigned* __restrict__ out_idx, float* __restrict__ out_val) {
    const int64_t t = blockIdx.x;
    ArgBest b = {-INFINITY, 0xffffffffu};
    for (int i = threadIdx.x; i < n; i += blockDim.x) best_update(b, part[t * n + i].v, part[t * n + i].idx);
    __shared__ ArgBest sb[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best_update(b, __shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.idx, o));
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w) best_update(b, sb[w].v, sb[w].idx);
        out_idx[t] = b.idx == 0xffffffffu ? 0u : b.idx;          // never hand an out-of-range index to the consumers
        out_val[t] = b.v;
    }
}

// z-score the (h, w) template over its own pixels and embed it at (y0, x0) in a zero (ny, nx) frame
// (signal/tracking.py:251-260). mean / variance come from the frame-reduction table of the template.
// blockIdx.y = template index (one padded frame each); the templates are h*w apart, their reduction rows B4D_FR_NCOLS
__global__ void __launch_bounds__(256) embed_template_kernel(const float* __restrict__ tpl, int h, int w, int ny, int nx,
                                                             int y0, int x0, float eps, const double* __restrict__ fr,
                                                             float* __restrict__ out) {
    tpl += (size_t)blockIdx.y * h * w;
    fr += (size_t)blockIdx.y * B4D_FR_NCOLS;
    out += (size_t)blockIdx.y * ny * nx;
    const float m = (float)fr[B4D_FR_MEAN];
    const float inv = 1.f / ((float)sqrt(fr[B4D_FR_M2]) + eps);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ny * nx; i += gridDim.x * blockDim.x) {
        const int y = i / nx, xq = i - y * nx;
        const int ty = y - y0, tx = xq - x0;
        float v = 0.f;
        if (ty >= 0 && ty < h && tx >= 0 && tx < w) v = (tpl[ty * w + tx] - m) * inv;
        out[i] = v;
    }
}

// (dy, dx, peak, snr) per frame from the |corr| map (signal/tracking.py:283-297, 314-375)
__global__ void phase_finalize_kernel(const float* __restrict__ mag, const unsigned* __restrict__ peak_idx, int ny, int nx,
                                      const float* __restrict__ med2, const long long* __restrict__ nvalid, int subpixel,
                                      double eps, double* __restrict__ out, int64_t T) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* c = mag + (size_t)t * ny * nx;
    const int i = (int)(peak_idx[t] / (unsigned)nx), j = (int)(peak_idx[t] % (unsigned)nx);
    const float pk = c[(size_t)i * nx + j];
    double dy = (double)(i - ny / 2), dx = (double)(j - nx / 2);
    if (subpixel && i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
        auto C = [&](int a, int b) { return c[(size_t)a * nx + b]; };
        // float32 scalar arithmetic, one rounding per operation, as numpy evaluates it
        const float gy = __fdiv_rn(__fsub_rn(C(i + 1, j), C(i - 1, j)), 2.f);
        const float gyy = __fsub_rn(__fadd_rn(C(i + 1, j), C(i - 1, j)), __fmul_rn(2.f, C(i, j)));
        const float gx = __fdiv_rn(__fsub_rn(C(i, j + 1), C(i, j - 1)), 2.f);
        const float gxx = __fsub_rn(__fadd_rn(C(i, j + 1), C(i, j - 1)), __fmul_rn(2.f, C(i, j)));
        const float gxy = __fdiv_rn(__fadd_rn(__fsub_rn(__fsub_rn(C(i + 1, j + 1), C(i + 1, j - 1)), C(i - 1, j + 1)), C(i - 1, j - 1)), 4.f);
        const float det = __fsub_rn(__fmul_rn(gxx, gyy), __fmul_rn(gxy, gxy));
        if (det != 0.f) {
            const float inv = __fdiv_rn(1.f, det);
            // NOTE: the reference returns the x step first and adds it to dy
            const float di = __fmul_rn(-__fsub_rn(__fmul_rn(gyy, gx), __fmul_rn(gxy, gy)), inv);
            const float dj = __fmul_rn(-__fsub_rn(__fmul_rn(gxx, gy), __fmul_rn(gxy, gx)), inv);
            dy += (double)di;
            dx += (double)dj;
        }
    }
    // np.median of a float32 array: mean of the two middle values, rounded to float32
    const float med = (nvalid[t] & 1) ? med2[2 * t] : __fmul_rn(__fadd_rn(med2[2 * t], med2[2 * t + 1]), 0.5f);
    double* o = out + t * 4;
    o[0] = dy; o[1] = dx; o[2] = (double)pk; o[3] = fabs((double)pk) / ((double)med + eps);
}

// Row blocks (of rpc rows, unshifted order) that hold the shifted rows i-1, i, i+1 around every frame's peak
__global__ void window_blocks_kernel(const unsigned* __restrict__ peak_idx, int ny, int nx, int rpc, int* __restrict__ blk3,
                                     int64_t T) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int i = (int)(peak_idx[t] / (unsigned)nx);
    for (int d = -1; d <= 1; ++d) {
        const int rs = min(max(i + d, 0), ny - 1);
        const int y = (rs + ny / 2) & (ny - 1);
        blk3[t * 3 + d + 1] = y / rpc;
    }
}

// phase_finalize_kernel for the fused median: the 3x3 neighbourhood of the peak comes from the compact window
// (T, 3 * rpc, nx) that rows_inv_kernel recomputed for the blocks named in blk3; frames whose median bracket missed
// (need[t]) report snr = NaN and are redone by the caller through the map-based path.
__global__ void phase_finalize_window_kernel(const float* __restrict__ win, const int* __restrict__ blk3, int rpc,
                                             const unsigned* __restrict__ peak_idx, int ny, int nx,
                                             const float* __restrict__ med2, const long long* __restrict__ nvalid,
                                             const int* __restrict__ need, int subpixel, double eps,
                                             double* __restrict__ out, int64_t T) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* wt = win + (size_t)t * 3 * rpc * nx;
    const int* b3 = blk3 + t * 3;
    auto C = [&](int a, int b) {
        const int y = (a + ny / 2) & (ny - 1), blk = y / rpc, r = y % rpc;
        const int slot = b3[0] == blk ? 0 : (b3[1] == blk ? 1 : 2);
        return wt[((size_t)slot * rpc + r) * nx + b];
    };
    const int i = (int)(peak_idx[t] / (unsigned)nx), j = (int)(peak_idx[t] % (unsigned)nx);
    const float pk = C(i, j);
    double dy = (double)(i - ny / 2), dx = (double)(j - nx / 2);
    if (subpixel && i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
        // float32 scalar arithmetic, one rounding per operation, as numpy evaluates it
        const float gy = __fdiv_rn(__fsub_rn(C(i + 1, j), C(i - 1, j)), 2.f);
        const float gyy = __fsub_rn(__fadd_rn(C(i + 1, j), C(i - 1, j)), __fmul_rn(2.f, C(i, j)));
        const float gx = __fdiv_rn(__fsub_rn(C(i, j + 1), C(i, j - 1)), 2.f);
        const float gxx = __fsub_rn(__fadd_rn(C(i, j + 1), C(i, j - 1)), __fmul_rn(2.f, C(i, j)));
        const float gxy = __fdiv_rn(__fadd_rn(__fsub_rn(__fsub_rn(C(i + 1, j + 1), C(i + 1, j - 1)), C(i - 1, j + 1)), C(i - 1, j - 1)), 4.f);
        const float det = __fsub_rn(__fmul_rn(gxx, gyy), __fmul_rn(gxy, gxy));
        if (det != 0.f) {
            const float inv = __fdiv_rn(1.f, det);
            // NOTE: the reference returns the x step first and adds it to dy
            const float di = __fmul_rn(-__fsub_rn(__fmul_rn(gyy, gx), __fmul_rn(gxy, gy)), inv);
            const float dj = __fmul_rn(-__fsub_rn(__fmul_rn(gxx, gy), __fmul_rn(gxy, gx)), inv);
            dy += (double)di;
            dx += (double)dj;
        }
    }
    const float med = (nvalid[t] & 1) ? med2[2 * t] : __fmul_rn(__fadd_rn(med2[2 * t], med2[2 * t + 1]), 0.5f);
    double* o = out + t * 4;
    o[0] = dy; o[1] = dx; o[2] = (double)pk;
    o[3] = need[t] ? nan("") : fabs((double)pk) / ((double)med + eps);
}

// grain(): widths of the peak-normalised autocorrelation (metrics/speckles.py:546-575)
//   lx, ly : width_at_fraction of the row / column through the argmax (maths/stats.py:45-89)
//   leq    : 2 * distance_at_fraction_from_peak(radial_mean_interpolated(ac)) (maths/stats.py:127-155,
//            maths/radial.py:132-169); radii are evaluated in increasing order until the first crossing,
//            which is all the first-crossing rule ever looks at.
__device__ double cut_width(const float* __restrict__ p, int stride, int n, int cidx, double fraction, double* sh_i) {
    // parallel search of the first sample below thr walking left and right from cidx
    const double peak = (double)p[(size_t)cidx * stride], thr = peak * fraction;
    int il = -1, ir = n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool below = (double)p[(size_t)i * stride] < thr;
        if (below && i <= cidx) il = max(il, i);
        if (below && i >= cidx) ir = min(ir, i);
    }
    __shared__ int s_l[32], s_r[32];
    for (int o = 16; o > 0; o >>= 1) { il = max(il, __shfl_xor_sync(0xffffffffu, il, o)); ir = min(ir, __shfl_xor_sync(0xffffffffu, ir, o)); }
    if ((threadIdx.x & 31) == 0) { s_l[threadIdx.x >> 5] = il; s_r[threadIdx.x >> 5] = ir; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { il = max(il, s_l[w]); ir = min(ir, s_r[w]); }
        double width;
        if (il < 0 || ir >= n) width = (double)n;
        else {
            double ya = p[(size_t)il * stride], yb = p[(size_t)(il + 1) * stride];
            const double xl = (yb == ya) ? (double)il : il + (thr - ya) / (yb - ya);
            ya = p[(size_t)(ir - 1) * stride]; yb = p[(size_t)ir * stride];
            const double xr = (yb == ya) ? (double)ir : (ir - 1) + (thr - ya) / (yb - ya);
            width = xr - xl;
        }
        *sh_i = width;
    }
    __syncthreads();
    return *sh_i;
}

__global__ void __launch_bounds__(1024) grain_kernel(const float* __restrict__ ac, int n, const unsigned* __restrict__ peak_idx,
                                                     const double* __restrict__ theta, double fraction,
                                                     double* __restrict__ out) {
    const int64_t t = blockIdx.x;
    const float* a = ac + (size_t)t * n * n;
    const int iy = (int)(peak_idx[t] / (unsigned)n), ix = (int)(peak_idx[t] % (unsigned)n);
    __shared__ double s_w;
    __shared__ double rad[32];
    const double ly = cut_width(a + ix, n, n, iy, fraction, &s_w);
    __syncthreads();
    const double lx = cut_width(a + (size_t)iy * n, 1, n, ix, fraction, &s_w);
    __syncthreads();

    // radial profile about (n//2, n//2): r_k = k (r_max = n/2, nr = n/2 + 1 -> dr = 1 exactly for even n)
    const int cy = n / 2, cx = n / 2;
    const double r_max = (double)(n / 2);
    const int nr = n / 2 + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double peak0 = 0.0, prev = 0.0, dist = (double)nr;
    bool found = false;
    for (int k0 = 0; k0 < nr && !found; k0 += 32) {
        // warp w evaluates radius k0 + w
        const int k = k0 + warp;
        double s = 0.0;
        if (k < nr) {
            const double r = r_max * (double)k / (double)(nr - 1);
            for (int m = lane; m < NTHETA; m += 32) {
                const double yy = r * theta[2 * m] + cy, xx = r * theta[2 * m + 1] + cx;
                if (yy >= 0.0 && yy <= (double)(n - 1) && xx >= 0.0 && xx <= (double)(n - 1)) {
                    int y0 = min((int)floor(yy), n - 2), x0 = min((int)floor(xx), n - 2);
                    const double ty = yy - y0, tx = xx - x0;
                    const float* q = a + (size_t)y0 * n + x0;
                    s += (double)q[0] * (1.0 - ty) * (1.0 - tx) + (double)q[1] * (1.0 - ty) * tx +
                         (double)q[n] * ty * (1.0 - tx) + (double)q[n + 1] * ty * tx;
                }
            }
            s = warp_sum(s) / (double)NTHETA;
        }
        __syncthreads();
        if (lane == 0) rad[warp] = s;
        __syncthreads();
        // every thread scans the 32 new radii identically (cheap, keeps control flow uniform)
        for (int w = 0; w < 32 && k0 + w < nr && !found; ++w) {
            const double v = rad[w];
            const int kk = k0 + w;
            if (kk == 0) { peak0 = v; prev = v; continue; }
            const double thr = peak0 * fraction;
            if (v < thr) {
                const double xc = (v == prev) ? (double)kk : (kk - 1) + (thr - prev) / (v - prev);
                dist = xc;
                found = true;
            }
            prev = v;
        }
    }
    if (threadIdx.x == 0) {
        const double dr = r_max / (double)(nr - 1);
        double* o = out + t * 4;
        o[0] = lx; o[1] = ly; o[2] = 2.0 * dist * dr; o[3] = ly != 0.0 ? lx / ly : INFINITY;
    }
}

// out *= 1 / max|out| per frame (xcorr2d "peak"), max taken from the argmax partial of |values|
__global__ void __launch_bounds__(256) scale_by_kernel(float* __restrict__ data, int64_t n, const float* __restrict__ peak,
                                                       int use_reciprocal) {
    const int64_t t = blockIdx.y;
    const float m = peak[t];
    if (!(m > 0.f)) return;
    float* d = data + t * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[i] = use_reciprocal ? d[i] / m : d[i] * m;
}

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ data, int64_t n, float* __restrict__ out) {
    const int64_t t = blockIdx.x;
    const float* d = data + t * n;
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(d[i]));
    m = warp_max(m);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]); out[t] = m; }
}

// spectral table (T, B4D_SP_NCOLS) from the per-tile partials, fixed summation order
__global__ void spec_finalize_kernel(const double* __restrict__ part, int ntiles, double* __restrict__ out, int64_t T) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double s[NSP] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ntiles; ++i)
        for (int k = 0; k < NSP; ++k) s[k] += part[((size_t)t * ntiles + i) * NSP + k];
    double* o = out + t * B4D_SP_NCOLS;
    o[B4D_SP_TOTAL] = s[0]; o[B4D_SP_FX2] = s[1]; o[B4D_SP_FY2] = s[2]; o[B4D_SP_P2] = s[3];
    o[B4D_SP_ALL] = s[4]; o[B4D_SP_PLOGP] = s[5];
    o[B4D_SP_F95] = nan(""); o[B4D_SP_NCOLS - 1] = 0.0;
}

// ---- f95: radius at which the radius-sorted cumulative PSD first reaches 0.95 (speckles.py:783-790) ----
// Two levels over integer keys r2 = kx^2 + ky^2 (square frames): coarse bins of 1024 keys, then the
// exact key inside the crossing bin.
__global__ void __launch_bounds__(256) f95_hist_kernel(const float* __restrict__ psd, int n, int level,
                                                       const int* __restrict__ coarse_bin, double* __restrict__ hist, int nb) {
    const int64_t t = blockIdx.y;
    const float* p = psd + (size_t)t * n * n;
    extern __shared__ double shh[];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) shh[i] = 0.0;
    __syncthreads();
    const int h = n / 2;
    const int cb = level ? coarse_bin[t] : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / n), xq = (int)(i % n);
        const int ky = y - h, kx = xq - h;
        const long long r2 = (long long)ky * ky + (long long)kx * kx;
        if (r2 == 0 || r2 > (long long)h * h) continue;
        const double v = (double)p[i];
        if (!level) atomicAdd(&shh[(int)(r2 >> 10)], v);
        else if ((int)(r2 >> 10) == cb) atomicAdd(&shh[(int)(r2 & 1023)], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (shh[i] != 0.0) atomicAdd(hist + (size_t)t * nb + i, shh[i]);
}

__global__ void f95_scan_kernel(double* __restrict__ hist, int nb, int level, const double* __restrict__ spec_tab,
                                int* __restrict__ coarse_bin, double* __restrict__ below, int n, double* __restrict__ out,
                                int64_t T) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double* h = hist + (size_t)t * nb;
    const double total = spec_tab[t * B4D_SP_NCOLS + B4D_SP_TOTAL];
    double cum = level ? below[t] : 0.0;
    int hit = nb - 1;
    bool found = false;
    for (int i = 0; i < nb; ++i) {
        if (!found && (cum + h[i]) / total >= 0.95 && h[i] != 0.0) { hit = i; found = true; }
        if (!found) cum += h[i];
        h[i] = 0.0;   // ready for the next use
    }
    if (!level) { coarse_bin[t] = hit; below[t] = cum; }
    else {
        const long long r2 = ((long long)coarse_bin[t] << 10) + hit;
        out[t * B4D_SP_NCOLS + B4D_SP_F95] = sqrt((double)r2) / (double)n;
    }
}

// -------------------------------------------------------------------------------------------------
// host-side launch helpers
// -------------------------------------------------------------------------------------------------
int rows_prefetch_dist(b4d_ctx* ctx) {               // experiment knob: B4D_ROWS_PREFETCH = distance in % of the resident CTAs (0 = off)
    static int pct = -1;
    if (pct < 0) { const char* e = getenv("B4D_ROWS_PREFETCH"); pct = e ? atoi(e) : 100; }
    return ctx->sm_count * 2 * pct / 100;
}

template <int NX>
int launch_rows_fwd(b4d_ctx* ctx, const RowsFwdArgs& a, int64_t T) {
    constexpr int TPF = NX / 16, FPC = 512 / TPF, ROWS = 2 * FPC;
    constexpr size_t smem = (size_t)FPC * (padded_len(NX) + 8) * sizeof(float2);
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) { B4D_CUDA(ctx, cudaFuncSetAttribute(rows_fwd_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[ctx->device] = true; }
    if (a.ny % ROWS) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "ny=%d is not a multiple of %d rows per CTA", a.ny, ROWS);
    ProfScope ps(ctx, KC_ROWS_FWD);
    RowsFwdArgs b = a;
    b.pf_dist = rows_prefetch_dist(ctx);
    b.keep = ctx->keep_mode;
    rows_fwd_kernel<NX><<<dim3(a.ny / ROWS, (unsigned)T), 512, smem, ctx->stream>>>(b);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

// Tensor map of a (T, ny, nx) float32 map stack viewed as a 2-D tensor [T * ny rows][nx columns], boxes of 256 rows x cw
// columns (cuTensorMapEncodeTiled, fetched from the driver through the runtime: no link against libcuda).
bool psd_tma_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("B4D_PSD_TMA"); on = (e && atoi(e) == 0) ? 0 : 1; }
    return on != 0;
}

bool make_psd_tmap(float* base, int64_t T, int ny, int nx, int cw, CUtensorMap* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeFn>(fn);
        else
            cudaGetLastError();
    }
    if (!encode || (reinterpret_cast<uintptr_t>(base) & 15) || (nx % 4)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)T * (cuuint64_t)ny};
    const cuuint64_t strides[1] = {(cuuint64_t)nx * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)TMA_BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NY, int CW, bool SPEC, bool AC, bool PC>
int launch_cols_inst(b4d_ctx* ctx, const ColsArgs& a, int64_t T) {
    constexpr size_t smem = (size_t)padded_len(NY) * CW * sizeof(float2) +
                            (AC ? (size_t)NY * CW * sizeof(float) + (size_t)NY * sizeof(float) : 0);
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) {
        B4D_CUDA(ctx, cudaFuncSetAttribute(cols_kernel<NY, CW, SPEC, AC, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[ctx->device] = true;
    }
    ProfScope ps(ctx, KC_COLS);
    static int pf_pct = -1;                           // experiment knob: B4D_COLS_PREFETCH = distance in % of the resident CTAs (0 = off)
    if (pf_pct < 0) { const char* e = getenv("B4D_COLS_PREFETCH"); pf_pct = e ? atoi(e) : 20; }
    ColsArgs b = a;
    b.pf_dist = ctx->sm_count * (1024 / (NY / 16 * CW)) * pf_pct / 100;
    b.keep = ctx->keep_mode;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    b.psd_tma = 0;
    if (CW == TC && NY >= TMA_BOX_ROWS && a.psd_out && psd_tma_enabled() && make_psd_tmap(a.psd_out, T, NY, a.nx, CW, &tmap)) b.psd_tma = 1;
    cols_kernel<NY, CW, SPEC, AC, PC><<<dim3(a.nx / 2 / CW, (unsigned)T), NY / 16 * CW, smem, ctx->stream>>>(b, tmap);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

// Columns per CTA of the column pass at ny = 2048. A CTA's row-major map stores (PSD, complex spectrum) are one
// contiguous piece of CW elements per row, and the load/store pipe pays per piece: with a map to write, 8 columns (one
// 1024-thread CTA per SM, 32-byte pieces) beat 4 (two 512-thread CTAs, 16-byte pieces) -- 2.87 vs 3.14 ms per 128 frames
// in the fused pipeline, 1.33 vs 1.72 ms PSD only; without one, two independent CTAs per SM overlap their phases better
// (tracker only: 1.71 vs 1.89 ms). `wide_out`: the launch writes such a map. Experiment knob: B4D_COLS_CW=4|8.
int cols_cw(int ny, bool wide_out) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("B4D_COLS_CW"); forced = (e && atoi(e) == 8) ? 8 : (e && atoi(e) == 4 ? 4 : 0); }
    if (ny < 2048) return TC;
    return forced ? forced : (wide_out ? 8 : B4D_COLS_CW_2048);
}
int cols_tiles(int ny, int nx, bool wide_out) { return nx / 2 / cols_cw(ny, wide_out); }
bool cols_wide_out(const ColsArgs& a) { return a.psd_out != nullptr || a.cplx_out != nullptr; }

template <int NY, int CW>
int launch_cols_cw(b4d_ctx* ctx, ColsArgs& a, int64_t T);

template <int NY>
int launch_cols(b4d_ctx* ctx, ColsArgs& a, int64_t T) {
    if (NY >= 2048 && cols_cw(NY, cols_wide_out(a)) == 4) return launch_cols_cw<NY, (NY >= 2048 ? 4 : TC)>(ctx, a, T);
    return launch_cols_cw<NY, TC>(ctx, a, T);
}

template <int NY, int CW>
int launch_cols_cw(b4d_ctx* ctx, ColsArgs& a, int64_t T) {
    const bool spec = a.spec_partials != nullptr, ac = a.i2_ac != nullptr, pc = a.i2_pc != nullptr;
    if (spec && pc && !ac) return b4d_fail(ctx, B4D_ERR_INVALID, "cols: spectral sums next to the product branch need the autocorrelation branch too");
    if (spec && ac && pc) return launch_cols_inst<NY, CW, true, true, true>(ctx, a, T);
    if (spec && ac) return launch_cols_inst<NY, CW, true, true, false>(ctx, a, T);
    if (spec) return launch_cols_inst<NY, CW, true, false, false>(ctx, a, T);
    if (ac && pc) return launch_cols_inst<NY, CW, false, true, true>(ctx, a, T);
    if (ac) return launch_cols_inst<NY, CW, false, true, false>(ctx, a, T);
    if (pc) return launch_cols_inst<NY, CW, false, false, true>(ctx, a, T);
    return launch_cols_inst<NY, CW, false, false, false>(ctx, a, T);
}

// grid_blocks: row blocks per frame this launch covers (0 = all of them; otherwise a.blk_map / a.blk_map_pf name them)
template <int NX, int MODE, bool ABS>
int launch_rows_inv_inst(b4d_ctx* ctx, RowsInvArgs& a, int64_t T, int grid_blocks) {
    constexpr int TPF = NX / 16, WPG = TPF / 8, GPC = 16 / WPG, FPC = 4 * GPC;
    constexpr size_t smem = (size_t)FPC * (padded_len(NX) + 8) * sizeof(float2);
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) { B4D_CUDA(ctx, cudaFuncSetAttribute(rows_inv_kernel<NX, MODE, ABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[ctx->device] = true; }
    const int rows = 2 * FPC;
    if (a.ny % rows) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "ny=%d is not a multiple of %d rows per CTA", a.ny, rows);
    a.nblk = a.ny / rows;
    ProfScope ps(ctx, KC_ROWS_INV);
    a.pf_dist = rows_prefetch_dist(ctx);
    a.keep = ctx->keep_mode;
    rows_inv_kernel<NX, MODE, ABS><<<dim3(grid_blocks > 0 ? grid_blocks : a.ny / rows, (unsigned)T), 512, smem, ctx->stream>>>(a);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

template <int NX>
int launch_rows_inv(b4d_ctx* ctx, RowsInvArgs& a, int64_t T, int grid_blocks) {
    if (a.mag_mode == 2) return launch_rows_inv_inst<NX, 2, true>(ctx, a, T, grid_blocks);
    if (a.mag_mode == 1) return launch_rows_inv_inst<NX, 1, true>(ctx, a, T, grid_blocks);
    if (a.kindA) return launch_rows_inv_inst<NX, 0, true>(ctx, a, T, grid_blocks);
    return launch_rows_inv_inst<NX, 0, false>(ctx, a, T, grid_blocks);
}

template <int NX>
int launch_rows_inv_ac(b4d_ctx* ctx, RowsInvAcArgs& a, int64_t T, int* nblk_out) {
    constexpr int TPF = NX / 16, WPG = TPF / 8, GPC = 16 / WPG, FPC = 4 * GPC;
    constexpr size_t smem = (size_t)FPC * (padded_len(NX) + 4) * sizeof(float2);
    static bool attr[B4D_MAX_DEVICES] = {};
    if (!attr[ctx->device]) { B4D_CUDA(ctx, cudaFuncSetAttribute(rows_inv_ac_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[ctx->device] = true; }
    const int rows = 2 * FPC, nblk = (a.ny / 2 + 1 + rows - 1) / rows;
    if (a.ny % rows) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "ny=%d is not a multiple of %d rows per CTA", a.ny, rows);
    ProfScope ps(ctx, KC_ROWS_INV_AC);
    a.keep = ctx->keep_mode;
    rows_inv_ac_kernel<NX><<<dim3(nblk, (unsigned)T), 512, smem, ctx->stream>>>(a);
    B4D_LAUNCH_CHECK(ctx);
    *nblk_out = nblk;
    return B4D_OK;
}

// CTAs per frame of the column pass (= number of per-frame partial sums it leaves behind)

int rows_inv_blocks(int nx, int ny) {
    const int tpf = nx / 16, wpg = tpf / 8, gpc = 16 / wpg, fpc = 4 * gpc;
    return ny / (2 * fpc);
}

#define DISPATCH_N(n, CALL)                                   \
    switch (n) {                                              \
        case 128: { constexpr int N_ = 128; CALL; } break;    \
        case 256: { constexpr int N_ = 256; CALL; } break;    \
        case 512: { constexpr int N_ = 512; CALL; } break;    \
        case 1024: { constexpr int N_ = 1024; CALL; } break;  \
        default: { constexpr int N_ = 2048; CALL; } break;    \
    }

int check_fft_args(b4d_ctx* ctx, const char* who, const void* stack, int64_t T, int ny, int nx) {
    if (!stack || T < 1) return b4d_fail(ctx, B4D_ERR_INVALID, "%s: bad arguments", who);
    if (!fft_size_ok(ny) || !fft_size_ok(nx) || T > 65535)
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED,
                        "%s: frames must have power-of-two sides in [128, 2048] and at most 65535 frames per call; got T=%lld (ny, nx)=(%d, %d)",
                        who, (long long)T, ny, nx);
    return B4D_OK;
}

// Workspace carve-up for one call (all sizes per frame, times T)
struct Work {
    float* pilot = nullptr;
    float2* H = nullptr;
    float2* I2a = nullptr;
    float2* I2b = nullptr;
    float2* I2nyq = nullptr;    // (T, ny) Nyquist column of the packed autocorrelation intermediate
    double* mom = nullptr;      // (T, ny, 2) generous: per-CTA moment sums of the forward row pass
    double* acp = nullptr;
    double* spp = nullptr;
    ArgBest* bestA = nullptr;
    ArgBest* bestB = nullptr;
    unsigned* pk_idx = nullptr;
    float* pk_val = nullptr;
    unsigned* pk_idx2 = nullptr;    // the autocorrelation branch's own peak (it may run next to the tracker's branch)
    float* pk_val2 = nullptr;
    float* med = nullptr;
    long long* nvalid = nullptr;
};

// tH / tA / tB: frames the three big intermediates are sized for (0 = T, the whole batch)
int carve(b4d_ctx* ctx, int64_t T, int ny, int nx, bool needA, bool needB, Work* w, int64_t tH = 0, int64_t tA = 0, int64_t tB = 0) {
    const size_t per = (size_t)ny * (nx / 2) * sizeof(float2);
    void* p = nullptr;
    int rc = b4d_scratch(ctx, SCR_SPEC_A, per * (tH ? tH : T), &p);
    if (rc) return rc;
    w->H = static_cast<float2*>(p);
    if (needA) { rc = b4d_scratch(ctx, SCR_SPEC_B, per * (tA ? tA : T), &p); if (rc) return rc; w->I2a = static_cast<float2*>(p); }
    if (needB) { rc = b4d_scratch(ctx, SCR_SPEC_C, per * (tB ? tB : T), &p); if (rc) return rc; w->I2b = static_cast<float2*>(p); }
    const int ntiles = cols_tiles(ny, nx, false), nblk = ny;   // (upper bounds: narrow tiles; rows_inv CTAs per frame)
    size_t small = 0;
    auto take = [&](size_t bytes) { size_t o = small; small += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_pilot = take(sizeof(float) * T), o_acp = take(sizeof(double) * T * ntiles),
                 o_spp = take(sizeof(double) * T * ntiles * NSP), o_ba = take(sizeof(ArgBest) * T * nblk),
                 o_bb = take(sizeof(ArgBest) * T * nblk), o_pi = take(sizeof(unsigned) * T), o_pv = take(sizeof(float) * T),
                 o_pi2 = take(sizeof(unsigned) * T), o_pv2 = take(sizeof(float) * T),
                 o_med = take(sizeof(float) * 2 * T), o_nv = take(sizeof(long long) * T),
                 o_nyq = take(sizeof(float2) * T * ny), o_mom = take(sizeof(double) * 2 * T * ny);
    rc = b4d_scratch(ctx, SCR_MISC, small + 1024, &p);
    if (rc) return rc;
    char* base = static_cast<char*>(p) + 512;   // first 512 B reserved (quantiles upload)
    w->pilot = reinterpret_cast<float*>(base + o_pilot);
    w->acp = reinterpret_cast<double*>(base + o_acp);
    w->spp = reinterpret_cast<double*>(base + o_spp);
    w->bestA = reinterpret_cast<ArgBest*>(base + o_ba);
    w->bestB = reinterpret_cast<ArgBest*>(base + o_bb);
    w->pk_idx = reinterpret_cast<unsigned*>(base + o_pi);
    w->pk_val = reinterpret_cast<float*>(base + o_pv);
    w->pk_idx2 = reinterpret_cast<unsigned*>(base + o_pi2);
    w->pk_val2 = reinterpret_cast<float*>(base + o_pv2);
    w->med = reinterpret_cast<float*>(base + o_med);
    w->nvalid = reinterpret_cast<long long*>(base + o_nv);
    w->I2nyq = reinterpret_cast<float2*>(base + o_nyq);
    w->mom = reinterpret_cast<double*>(base + o_mom);
    return B4D_OK;
}

int ensure_theta(b4d_ctx* ctx) {
    if (!ctx->fft) ctx->fft = new FftPlanCache();
    if (ctx->fft->theta) return B4D_OK;
    std::vector<double> h(2 * NTHETA);
    for (int m = 0; m < NTHETA; ++m) {
        // np.linspace(0, 2*pi, ntheta, endpoint=False)[m] = m * (2*pi/ntheta)
        const double th = (double)m * ((2.0 * 3.141592653589793) / (double)NTHETA);
        h[2 * m] = sin(th);
        h[2 * m + 1] = cos(th);
    }
    B4D_CUDA(ctx, cudaMalloc(&ctx->fft->theta, h.size() * sizeof(double)));
    B4D_CUDA(ctx, cudaMemcpy(ctx->fft->theta, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return B4D_OK;
}

// rows forward for a batch, pilot included
// fr_moments: the row pass also sums x - K and (x - K)^2 and the table's mean / variance columns are filled from them
int run_rows_fwd(b4d_ctx* ctx, const float* stack, int64_t T, int ny, int nx, const float* gain, const float* dark,
                 Work& w, bool use_pilot, bool pilot_ready = false, double* fr_moments = nullptr) {
    int rc;
    if (use_pilot && !pilot_ready) { rc = b4d_frame_pilot_launch(ctx, stack, T, (int64_t)ny * nx, gain, dark, w.pilot); if (rc) return rc; }
    RowsFwdArgs a;
    a.stack = stack; a.gain = gain; a.dark = dark; a.pilot = use_pilot ? w.pilot : nullptr; a.H = w.H; a.ny = ny;
    a.mom = fr_moments ? w.mom : nullptr;
    a.pf_dist = 0; a.keep = 0;
    rc = get_twiddle_bases(ctx, nx, &a.tw);
    if (rc) return rc;
    DISPATCH_N(nx, rc = launch_rows_fwd<N_>(ctx, a, T));
    if (rc || !fr_moments) return rc;
    if (!use_pilot) return b4d_fail(ctx, B4D_ERR_INVALID, "rows_fwd: moments need the pilot shift");
    const int tpf = nx / 16, rows = 2 * (512 / tpf);
    rows_moments_finalize_kernel<<<(unsigned)T, 32, 0, ctx->stream>>>(w.mom, ny / rows, w.pilot, (double)ny * nx, fr_moments);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

ColsArgs cols_defaults(const Work& w, int nx, bool use_pilot) {
    ColsArgs c;
    memset(&c, 0, sizeof(c));
    c.H = w.H; c.pilot = use_pilot ? w.pilot : nullptr; c.nx = nx; c.psd_scale = 1.f;
    return c;
}

int run_cols(b4d_ctx* ctx, ColsArgs& c, int64_t T, int ny) {
    int rc = get_twiddle_bases(ctx, ny, &c.tw);
    if (rc) return rc;
    DISPATCH_N(ny, rc = launch_cols<N_>(ctx, c, T));
    return rc;
}

int run_rows_inv(b4d_ctx* ctx, RowsInvArgs& r, int64_t T, int nx, int grid_blocks = 0) {
    int rc = get_twiddle_bases(ctx, nx, &r.tw);
    if (rc) return rc;
    DISPATCH_N(nx, rc = launch_rows_inv<N_>(ctx, r, T, grid_blocks));
    return rc;
}

int run_rows_inv_ac(b4d_ctx* ctx, RowsInvAcArgs& r, int64_t T, int nx, int* nblk) {
    int rc = get_twiddle_bases(ctx, nx, &r.tw);
    if (rc) return rc;
    DISPATCH_N(nx, rc = launch_rows_inv_ac<N_>(ctx, r, T, nblk));
    return rc;
}

// Frames per internal batch. The kernels of this path are bound by the SM's load/store pipe, not by L2 hits, and the
// per-batch kernels that run one CTA per frame (sampling, final selects, grain) want many frames per launch: batches
// are as large as a scratch budget allows (a quarter of the free HBM, at most 128 frames).
int64_t batch_frames(b4d_ctx* ctx, int ny, int nx, int n_intermediates) {
    if (ctx->batch_override > 0) return ctx->batch_override;
    const size_t per = (size_t)ny * (nx / 2) * sizeof(float2) * (size_t)n_intermediates;
    // cudaMemGetInfo takes driver-wide locks (milliseconds now and then): ask once per scratch footprint
    for (int i = 0; i < 4; ++i) if (ctx->batch_key[i] == per && ctx->batch_val[i] > 0) return ctx->batch_val[i];
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)8 << 30;
    size_t held = 0;                                  // scratch already owned by this context counts as available
    for (int i = 0; i < B4D_NSCRATCH; ++i) held += ctx->scratch_bytes[i];
    int64_t b = (int64_t)(((free_b + held) / 4) / (per ? per : 1));
    if (b < 1) b = 1;
    if (b > 128) b = 128;
    for (int i = 0; i < 4; ++i) if (ctx->batch_val[i] == 0 || i == 3) { ctx->batch_key[i] = per; ctx->batch_val[i] = b; break; }
    return b;
}

// -------------------------------------------------------------------------------------------------
// template matching (normalised cross-correlation, signal/tracking.py:82-188 -> cv2.TM_CCOEFF_NORMED)
// -------------------------------------------------------------------------------------------------
// result(y, x) = sum_ij T'(i,j) I(y+i, x+j) / sqrt(sum T'^2 * sum_window (I - mean_window)^2),  (ny-h+1, nx-w+1) values.
// Numerator: circular correlation with the zero-mean template embedded at the origin (no wrap-around inside the valid
// range), through the FFT kernels above. Window sums: double-precision integral images of (I - K) and (I - K)^2, K the
// frame's pilot mean (the quotient does not depend on K), as cv2 takes them from integral(CV_64F).

// inclusive scans along x: one warp per row, P1 = sum v, P2 = sum v^2 (v = frame - K)
__global__ void __launch_bounds__(256) tm_scan_rows_kernel(const float* __restrict__ stack, const float* __restrict__ pilot, int ny, int nx,
                                                           double* __restrict__ P1, double* __restrict__ P2, int64_t rows) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float K = pilot[row / ny];
    const float* src = stack + row * nx;
    double* d1 = P1 + row * nx;
    double* d2 = P2 + row * nx;
    const int per = (nx + 31) / 32, x0 = lane * per, x1 = min(x0 + per, nx);
    double s1 = 0.0, s2 = 0.0;
    for (int x = x0; x < x1; ++x) { const double v = (double)(src[x] - K); s1 += v; s2 += v * v; }
    double o1 = s1, o2 = s2;                          // exclusive prefix over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double a = __shfl_up_sync(0xffffffffu, o1, o), b = __shfl_up_sync(0xffffffffu, o2, o);
        if (lane >= o) { o1 += a; o2 += b; }
    }
    o1 -= s1; o2 -= s2;
    for (int x = x0; x < x1; ++x) {
        const double v = (double)(src[x] - K);
        o1 += v; o2 += v * v;
        d1[x] = o1; d2[x] = o2;
    }
}

// in-place inclusive scans along y: one thread per column
__global__ void __launch_bounds__(256) tm_scan_cols_kernel(double* __restrict__ P1, double* __restrict__ P2, int ny, int nx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = blockIdx.y;
    if (x >= nx) return;
    double* p1 = P1 + (size_t)t * ny * nx + x;
    double* p2 = P2 + (size_t)t * ny * nx + x;
    double a = 0.0, b = 0.0;
    for (int y = 0; y < ny; ++y) {
        a += p1[(size_t)y * nx]; b += p2[(size_t)y * nx];
        p1[(size_t)y * nx] = a; p2[(size_t)y * nx] = b;
    }
}

// compact result map (T, oy, ox) from the shifted correlation map and the integral images (cv2's common_matchTemplate)
__global__ void __launch_bounds__(256) tm_normalise_kernel(const float* __restrict__ corr, const double* __restrict__ P1,
                                                           const double* __restrict__ P2, int ny, int nx, int h, int w,
                                                           const double* __restrict__ tfr, int tfr_stride, double eps,
                                                           float* __restrict__ out) {
    const int64_t t = blockIdx.y;
    tfr += (size_t)t * tfr_stride;
    const int oy = ny - h + 1, ox = nx - w + 1;
    const double area = (double)h * (double)w;
    // tpl_z = (tpl - mean) / (std + eps): its own standard deviation, and cv2's templNorm = sdv * sqrt(area)
    const double sd = sqrt(tfr[B4D_FR_M2]);
    const double tnorm = (sd / (sd + eps)) * sqrt(area);
    const float* c = corr + (size_t)t * ny * nx;
    const double* p1 = P1 + (size_t)t * ny * nx;
    const double* p2 = P2 + (size_t)t * ny * nx;
    auto box = [&](const double* p, int y, int xq) {
        const int ya = y - 1, yb = y + h - 1, xa = xq - 1, xb = xq + w - 1;
        double s = p[(size_t)yb * nx + xb];
        if (ya >= 0) s -= p[(size_t)ya * nx + xb];
        if (xa >= 0) s -= p[(size_t)yb * nx + xa];
        if (ya >= 0 && xa >= 0) s += p[(size_t)ya * nx + xa];
        return s;
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)oy * ox; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / ox), xq = (int)(i % ox);
        double num = (double)c[(size_t)((y + ny / 2) % ny) * nx + ((xq + nx / 2) % nx)];
        const double s1 = box(p1, y, xq), s2 = box(p2, y, xq);
        const double diff2 = fmax(s2 - s1 * s1 / area, 0.0);
        const double tt = diff2 <= 1.1920929e-6 * s2 ? 0.0 : sqrt(diff2) * tnorm;      // flat window: avoid rounding noise
        if (fabs(num) < tt) num /= tt;
        else if (fabs(num) < tt * 1.125) num = num > 0 ? 1.0 : -1.0;
        else num = 0.0;
        out[(size_t)t * oy * ox + i] = (float)num;
    }
}

// (dy, dx, peak, snr) from the compact result map (signal/tracking.py:170-188): argmax, float32 3x3 Taylor step with the
// reference's swapped terms, match position = peak + (h-1)/2 against the template's reference centre
__global__ void tm_finalize_kernel(const float* __restrict__ res, const unsigned* __restrict__ peak_idx, int oy, int ox, int h, int w,
                                   double ref_y, double ref_x, const float* __restrict__ med2, const long long* __restrict__ nvalid,
                                   int subpixel, double eps, double* __restrict__ out, int64_t T) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* c = res + (size_t)t * oy * ox;
    const int i = (int)(peak_idx[t] / (unsigned)ox), j = (int)(peak_idx[t] % (unsigned)ox);
    const float pk = c[(size_t)i * ox + j];
    double py = (double)i, px = (double)j;
    if (subpixel && i > 0 && i < oy - 1 && j > 0 && j < ox - 1) {
        auto C = [&](int a, int b) { return c[(size_t)a * ox + b]; };
        const float gy = __fdiv_rn(__fsub_rn(C(i + 1, j), C(i - 1, j)), 2.f);
        const float gyy = __fsub_rn(__fadd_rn(C(i + 1, j), C(i - 1, j)), __fmul_rn(2.f, C(i, j)));
        const float gx = __fdiv_rn(__fsub_rn(C(i, j + 1), C(i, j - 1)), 2.f);
        const float gxx = __fsub_rn(__fadd_rn(C(i, j + 1), C(i, j - 1)), __fmul_rn(2.f, C(i, j)));
        const float gxy = __fdiv_rn(__fadd_rn(__fsub_rn(__fsub_rn(C(i + 1, j + 1), C(i + 1, j - 1)), C(i - 1, j + 1)), C(i - 1, j - 1)), 4.f);
        const float det = __fsub_rn(__fmul_rn(gxx, gyy), __fmul_rn(gxy, gxy));
        if (det != 0.f) {
            const float inv = __fdiv_rn(1.f, det);
            py += (double)__fmul_rn(-__fsub_rn(__fmul_rn(gyy, gx), __fmul_rn(gxy, gy)), inv);
            px += (double)__fmul_rn(-__fsub_rn(__fmul_rn(gxx, gy), __fmul_rn(gxy, gx)), inv);
        }
    }
    const float med = (nvalid[t] & 1) ? med2[2 * t] : __fmul_rn(__fadd_rn(med2[2 * t], med2[2 * t + 1]), 0.5f);
    double* o = out + t * 4;
    o[0] = py + (double)(h - 1) / 2.0 - ref_y;
    o[1] = px + (double)(w - 1) / 2.0 - ref_x;
    o[2] = (double)pk;
    o[3] = fabs((double)pk) / ((double)med + eps);
}

// -------------------------------------------------------------------------------------------------
// arbitrary frame sides (Bluestein), see generic_dft.cuh
// -------------------------------------------------------------------------------------------------
#include "generic_dft.cuh"

GenCache*& gen_cache(b4d_ctx* ctx) {
    if (!ctx->fft) ctx->fft = new FftPlanCache();
    return reinterpret_cast<GenCache*&>(ctx->fft->gen);
}

// f95 of B4D_SP_*: two-level histogram over the integer radius keys of a shifted square PSD map (tab: (tc, B4D_SP_NCOLS))
int run_f95(b4d_ctx* ctx, const float* map_b, int n, int64_t tc, double* tab, int slot = SCR_SELECT) {
    const int nb0 = ((n / 2) * (n / 2) >> 10) + 1, nb1 = 1024;
    void* p = nullptr;
    const size_t hb = sizeof(double) * (size_t)tc * (nb0 > nb1 ? nb0 : nb1);
    int rc = b4d_scratch(ctx, slot, hb + (sizeof(int) + sizeof(double)) * tc + 256, &p);
    if (rc) return rc;
    double* hist = static_cast<double*>(p);
    double* below = reinterpret_cast<double*>(static_cast<char*>(p) + hb);
    int* cb = reinterpret_cast<int*>(below + tc);
    B4D_CUDA(ctx, cudaMemsetAsync(p, 0, hb, ctx->stream));
    int bx = (int)(((int64_t)n * n + 256 * 32 - 1) / (256 * 32));
    if (bx > 296) bx = 296;
    for (int level = 0; level < 2; ++level) {
        const int nb = level ? nb1 : nb0;
        f95_hist_kernel<<<dim3(bx, (unsigned)tc), 256, nb * sizeof(double), ctx->stream>>>(map_b, n, level, cb, hist, nb);
        B4D_LAUNCH_CHECK(ctx);
        f95_scan_kernel<<<(unsigned)((tc + 63) / 64), 64, 0, ctx->stream>>>(hist, nb, level, tab, cb, below, n, tab, tc);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

bool pow2_sides(int ny, int nx) { return fft_size_ok(ny) && fft_size_ok(nx); }

int check_gen_args(b4d_ctx* ctx, const char* who, const void* stack, int64_t T, int ny, int nx) {
    if (!stack || T < 1) return b4d_fail(ctx, B4D_ERR_INVALID, "%s: bad arguments", who);
    if (!gen_size_ok(ny) || !gen_size_ok(nx) || T > ((int64_t)1 << 24))
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED,
                        "%s: frames need sides in [1, 4096]; got T=%lld (ny, nx)=(%d, %d)",
                        who, (long long)T, ny, nx);
    return B4D_OK;
}

struct GenWork {
    float2* A = nullptr;
    float2* Bf = nullptr;
    double* fr = nullptr;
    double* spp = nullptr;
    unsigned* pk = nullptr;
    int nblk = 0;
};

int64_t gen_batch(int ny, int nx) {
    int64_t b = ((int64_t)512 << 20) / ((int64_t)16 * ny * nx);
    if (b < 1) b = 1;
    if (b > 32768) b = 32768;       // gridDim.y
    return b;
}

int gen_carve(b4d_ctx* ctx, int64_t tc, int ny, int nx, GenWork* w) {
    const size_t per = (size_t)ny * nx * sizeof(float2);
    void* p = nullptr;
    int rc;
    if ((rc = b4d_scratch(ctx, SCR_SPEC_A, per * tc, &p))) return rc;
    w->A = static_cast<float2*>(p);
    if ((rc = b4d_scratch(ctx, SCR_SPEC_B, per * tc, &p))) return rc;
    w->Bf = static_cast<float2*>(p);
    int nblk = (int)(((int64_t)ny * nx + 2047) / 2048);
    if (nblk > 64) nblk = 64;
    w->nblk = nblk;
    const size_t fr_b = (sizeof(double) * B4D_FR_NCOLS * tc + 255) & ~size_t(255);
    const size_t sp_b = (sizeof(double) * NSP * nblk * tc + 255) & ~size_t(255);
    if ((rc = b4d_scratch(ctx, SCR_NYQ, fr_b + sp_b + sizeof(unsigned) * tc + 256, &p))) return rc;
    w->fr = static_cast<double*>(p);
    w->spp = reinterpret_cast<double*>(static_cast<char*>(p) + fr_b);
    w->pk = reinterpret_cast<unsigned*>(static_cast<char*>(p) + fr_b + sp_b);
    return B4D_OK;
}

int gen_epilogue(b4d_ctx* ctx, const GenWork& w, int64_t tc, int ny, int nx, int dc_mode, float scale, float* psd, float2* cplx,
                 int sqmag, bool spectral) {
    GenEpiArgs e;
    memset(&e, 0, sizeof(e));
    e.F = w.A; e.fr = w.fr; e.ny = ny; e.nx = nx; e.dc_mode = dc_mode; e.scale = scale; e.psd_out = psd; e.cplx_out = cplx;
    e.sqmag_inplace = sqmag; e.spec_partials = spectral ? w.spp : nullptr;
    ProfScope ps(ctx, KC_GENERIC);
    gen_epilogue_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(e);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int gen_fft2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, float* out_c64) {
    const int64_t B = gen_batch(ny, nx);
    const size_t npix = (size_t)ny * nx;
    int rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        GenWork w;
        if ((rc = gen_carve(ctx, tc, ny, nx, &w))) return rc;
        const float* s0 = stack + t0 * npix;
        if ((rc = b4d_frame_reductions_nolock(ctx, s0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, w.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), s0, w.fr, tc, ny, nx, w.A, w.Bf))) return rc;
        if ((rc = gen_epilogue(ctx, w, tc, ny, nx, 0, 1.f, nullptr, reinterpret_cast<float2*>(out_c64) + t0 * npix, 0, false))) return rc;
    }
    return B4D_OK;
}

// phase_correlation for frame sides that are not powers of two (map-based: |corr| map, exact select for the median)
int gen_phase_set_reference(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx, int y0, int x0, double eps) {
    FftPlanCache* f = ctx->fft ? ctx->fft : (ctx->fft = new FftPlanCache());
    const size_t npix = (size_t)ny * nx;
    if (f->gref_elems != npix) {
        if (f->gref) { B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(f->gref); f->gref = nullptr; }
        B4D_CUDA(ctx, cudaMalloc(&f->gref, sizeof(float2) * npix));
        f->gref_elems = npix;
    }
    f->ref_ny = ny; f->ref_nx = nx;
    GenWork gw;
    int rc = gen_carve(ctx, 1, ny, nx, &gw);
    if (rc) return rc;
    void* p = nullptr;
    if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * npix + sizeof(double) * B4D_FR_NCOLS + 256, &p))) return rc;
    float* padded = static_cast<float*>(p);
    double* tfr = reinterpret_cast<double*>(padded + ((npix + 1) & ~size_t(1)));
    if ((rc = b4d_frame_reductions_nolock(ctx, tpl, 1, h, w, nullptr, nullptr, nan(""), 0.0, tfr))) return rc;
    embed_template_kernel<<<dim3(ctx->sm_count * 4, 1), 256, 0, ctx->stream>>>(tpl, h, w, ny, nx, y0, x0, (float)eps, tfr, padded);
    B4D_LAUNCH_CHECK(ctx);
    if ((rc = gen_forward(ctx, gen_cache(ctx), padded, nullptr, 1, ny, nx, gw.A, gw.Bf))) return rc;
    gen_conj_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(gw.A, (int64_t)npix);
    B4D_LAUNCH_CHECK(ctx);
    B4D_CUDA(ctx, cudaMemcpyAsync(f->gref, gw.A, sizeof(float2) * npix, cudaMemcpyDeviceToDevice, ctx->stream));
    return B4D_OK;
}

int gen_phase_track(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, int subpixel, double eps, double* out) {
    const size_t npix = (size_t)ny * nx;
    int64_t B = gen_batch(ny, nx);
    int rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        const float* s0 = stack + t0 * npix;
        GenWork w;
        if ((rc = gen_carve(ctx, tc, ny, nx, &w))) return rc;
        void* p = nullptr;
        if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * npix * tc + (sizeof(float) * 2 + sizeof(long long)) * tc + 512, &p))) return rc;
        float* mag = static_cast<float*>(p);
        float* med = mag + ((npix * tc + 63) & ~size_t(63));
        long long* nvalid = reinterpret_cast<long long*>(med + ((2 * tc + 1) & ~int64_t(1)));
        if ((rc = b4d_frame_reductions_nolock(ctx, s0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, w.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), s0, w.fr, tc, ny, nx, w.A, w.Bf))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_phase_product_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(w.A, ctx->fft->gref, w.fr, (int64_t)npix, (float)eps);
            B4D_LAUNCH_CHECK(ctx);
        }
        if ((rc = gen_inverse(ctx, gen_cache(ctx), tc, ny, nx, w.A, w.Bf))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_shift_out_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(w.A, ny, nx, (float)(1.0 / ((double)nx * (double)ny)), 1, mag);
            B4D_LAUNCH_CHECK(ctx);
            gen_argmax_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(mag, (int64_t)npix, w.pk);
            B4D_LAUNCH_CHECK(ctx);
        }
        void* q = nullptr;
        if ((rc = b4d_scratch(ctx, SCR_MISC, 1024, &q))) return rc;
        static const double half_q = 0.5;
        if ((rc = b4d_put_doubles(ctx, static_cast<double*>(q), &half_q, 1))) return rc;
        if ((rc = b4d_select_impl(ctx, mag, tc, (int64_t)npix, static_cast<const double*>(q), 1, 1, med, reinterpret_cast<int64_t*>(nvalid)))) return rc;
        phase_finalize_kernel<<<(unsigned)((tc + 63) / 64), 64, 0, ctx->stream>>>(mag, w.pk, ny, nx, med, nvalid, subpixel, eps, out + t0 * 4, tc);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

// template_matching for frame sides that are not powers of two: same steps as b4d_template_match with the chirp-z
// transforms in place of the FFT kernels (numerator on the mean-removed frame, window sums on the pilot-shifted one;
// the quotient depends on neither offset)
int gen_template_match(b4d_ctx* ctx, const float* tpl, int per_frame, int h, int w, const float* stack, int64_t n_frames, int ny, int nx,
                       double ref_y, double ref_x, int subpixel, double eps, double* out) {
    const size_t npix = (size_t)ny * nx;
    const int oy = ny - h + 1, ox = nx - w + 1;
    const size_t nres = (size_t)oy * ox;
    int64_t B = ((int64_t)1024 << 20) / (int64_t)(npix * (per_frame ? 48 : 40));
    if (B < 1) B = 1;
    if (B > n_frames) B = n_frames;
    const int64_t nt = per_frame ? B : 1;
    int rc;
    void* p = nullptr;
    if ((rc = b4d_scratch(ctx, SCR_SPEC_C, sizeof(float2) * npix * nt + sizeof(double) * B4D_FR_NCOLS * nt + 256, &p))) return rc;
    float2* tref = static_cast<float2*>(p);
    double* tfr = reinterpret_cast<double*>(tref + npix * nt);
    if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * (npix + nres) * B + sizeof(double) * 2 * npix * B +
                          (sizeof(float) * 3 + sizeof(long long)) * B + 1024, &p))) return rc;
    float* corr = static_cast<float*>(p);
    float* res = corr + npix * B;
    double* P1 = reinterpret_cast<double*>(res + ((nres * B + 1) & ~size_t(1)));
    double* P2 = P1 + npix * B;
    float* pilot = reinterpret_cast<float*>(P2 + npix * B);
    float* med = pilot + ((B + 1) & ~int64_t(1));
    long long* nvalid = reinterpret_cast<long long*>(med + 2 * B);
    auto template_spectra = [&](const float* tp, int64_t count) -> int {
        int r = b4d_frame_reductions_nolock(ctx, tp, count, h, w, nullptr, nullptr, nan(""), 0.0, tfr);
        if (r) return r;
        embed_template_kernel<<<dim3(ctx->sm_count * 2, (unsigned)count), 256, 0, ctx->stream>>>(tp, h, w, ny, nx, 0, 0, (float)eps, tfr, corr);
        B4D_LAUNCH_CHECK(ctx);
        GenWork gw;
        if ((r = gen_carve(ctx, count, ny, nx, &gw))) return r;
        if ((r = gen_forward(ctx, gen_cache(ctx), corr, nullptr, count, ny, nx, gw.A, gw.Bf))) return r;
        gen_conj_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(gw.A, (int64_t)(npix * count));
        B4D_LAUNCH_CHECK(ctx);
        B4D_CUDA(ctx, cudaMemcpyAsync(tref, gw.A, sizeof(float2) * npix * count, cudaMemcpyDeviceToDevice, ctx->stream));
        return B4D_OK;
    };
    if (!per_frame && (rc = template_spectra(tpl, 1))) return rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        const float* s0 = stack + t0 * npix;
        if (per_frame && (rc = template_spectra(tpl + (size_t)t0 * h * w, tc))) return rc;
        GenWork gw;
        if ((rc = gen_carve(ctx, tc, ny, nx, &gw))) return rc;
        if ((rc = b4d_frame_reductions_nolock(ctx, s0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, gw.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), s0, gw.fr, tc, ny, nx, gw.A, gw.Bf))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_mul_kernel<<<dim3(gw.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(gw.A, tref, per_frame ? npix : 0, (int64_t)npix);
            B4D_LAUNCH_CHECK(ctx);
        }
        if ((rc = gen_inverse(ctx, gen_cache(ctx), tc, ny, nx, gw.A, gw.Bf))) return rc;
        if ((rc = b4d_frame_pilot_launch(ctx, s0, tc, (int64_t)npix, nullptr, nullptr, pilot))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_shift_out_kernel<<<dim3(gw.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(gw.A, ny, nx, (float)(1.0 / ((double)nx * (double)ny)), 0, corr);
            B4D_LAUNCH_CHECK(ctx);
            const int64_t rows = tc * ny;
            tm_scan_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(s0, pilot, ny, nx, P1, P2, rows);
            B4D_LAUNCH_CHECK(ctx);
            tm_scan_cols_kernel<<<dim3((nx + 255) / 256, (unsigned)tc), 256, 0, ctx->stream>>>(P1, P2, ny, nx);
            B4D_LAUNCH_CHECK(ctx);
            int bx = (int)((nres + 2047) / 2048);
            if (bx > 592) bx = 592;
            tm_normalise_kernel<<<dim3(bx, (unsigned)tc), 256, 0, ctx->stream>>>(corr, P1, P2, ny, nx, h, w, tfr, per_frame ? B4D_FR_NCOLS : 0, eps, res);
            B4D_LAUNCH_CHECK(ctx);
            gen_argmax_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(res, (int64_t)nres, gw.pk);
            B4D_LAUNCH_CHECK(ctx);
        }
        void* q = nullptr;
        if ((rc = b4d_scratch(ctx, SCR_MISC, 1024, &q))) return rc;
        static const double half_q = 0.5;
        if ((rc = b4d_put_doubles(ctx, static_cast<double*>(q), &half_q, 1))) return rc;
        if ((rc = b4d_select_impl(ctx, res, tc, (int64_t)nres, static_cast<const double*>(q), 1, 1, med, reinterpret_cast<int64_t*>(nvalid)))) return rc;
        tm_finalize_kernel<<<(unsigned)((tc + 63) / 64), 64, 0, ctx->stream>>>(res, gw.pk, oy, ox, h, w, ref_y, ref_x, med, nvalid, subpixel, eps,
                                                                               out + t0 * 4, tc);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

// xcorr2d for frame sides that are not powers of two: ifft2(fft2(a) conj(fft2(b))), shifted, real part
int gen_xcorr2d(b4d_ctx* ctx, const float* a, const float* b, int64_t n_frames, int ny, int nx, int remove_mean, int normalize_peak,
                float* out) {
    int64_t B = gen_batch(ny, nx) / 2;
    if (B < 1) B = 1;
    const size_t npix = (size_t)ny * nx;
    int rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        GenWork w;
        if ((rc = gen_carve(ctx, tc, ny, nx, &w))) return rc;
        void* p = nullptr;
        const size_t frb = (sizeof(double) * B4D_FR_NCOLS * tc + 255) & ~size_t(255);
        if ((rc = b4d_scratch(ctx, SCR_SPEC_C, sizeof(float2) * npix * tc + frb + sizeof(float) * tc + 256, &p))) return rc;
        float2* Sb = static_cast<float2*>(p);
        double* fr_b = reinterpret_cast<double*>(Sb + npix * tc);
        float* pk = reinterpret_cast<float*>(reinterpret_cast<char*>(fr_b) + frb);
        const float* a0 = a + t0 * npix;
        const float* b0 = b + t0 * npix;
        float* o = out + t0 * npix;
        // spectrum of b (mean removed) -> Sb, then spectrum of a -> A
        if ((rc = b4d_frame_reductions_nolock(ctx, b0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, w.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), b0, w.fr, tc, ny, nx, w.A, w.Bf))) return rc;
        B4D_CUDA(ctx, cudaMemcpyAsync(Sb, w.A, sizeof(float2) * npix * tc, cudaMemcpyDeviceToDevice, ctx->stream));
        B4D_CUDA(ctx, cudaMemcpyAsync(fr_b, w.fr, sizeof(double) * B4D_FR_NCOLS * tc, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = b4d_frame_reductions_nolock(ctx, a0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, w.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), a0, w.fr, tc, ny, nx, w.A, w.Bf))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_cross_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(w.A, Sb, w.fr, fr_b, ny, nx, remove_mean ? 1 : 0);
            B4D_LAUNCH_CHECK(ctx);
        }
        if ((rc = gen_inverse(ctx, gen_cache(ctx), tc, ny, nx, w.A, w.Bf))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_real_out_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(w.A, ny, nx, (float)(1.0 / ((double)nx * (double)ny)), o);
            B4D_LAUNCH_CHECK(ctx);
        }
        if (normalize_peak) {
            absmax_kernel<<<(unsigned)tc, 256, 0, ctx->stream>>>(o, (int64_t)npix, pk);
            B4D_LAUNCH_CHECK(ctx);
            scale_by_kernel<<<dim3(64, (unsigned)tc), 256, 0, ctx->stream>>>(o, (int64_t)npix, pk, 1);
            B4D_LAUNCH_CHECK(ctx);
        }
    }
    return B4D_OK;
}

// ifft2d (signal/fft.py:240-258): ifft2(ifftshift(F)), complex in, complex out, any sides in [2, 4096]
int gen_ifft2d(b4d_ctx* ctx, const float2* spec, int64_t n_frames, int ny, int nx, float2* out) {
    const int64_t B = gen_batch(ny, nx);
    const size_t npix = (size_t)ny * nx;
    int rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        GenWork w;
        if ((rc = gen_carve(ctx, tc, ny, nx, &w))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_unshift_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(spec + t0 * npix, w.A, ny, nx, 1, 1.f);
            B4D_LAUNCH_CHECK(ctx);
        }
        if ((rc = gen_inverse(ctx, gen_cache(ctx), tc, ny, nx, w.A, w.Bf))) return rc;
        ProfScope ps(ctx, KC_GENERIC);
        gen_unshift_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(w.A, out + t0 * npix, ny, nx, 0,
                                                                               (float)(1.0 / ((double)nx * (double)ny)));
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

int gen_psd2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, float scale_factor, int sub_mean, int zero_dc,
              float* out_psd, double* spectral) {
    const int64_t B = gen_batch(ny, nx);
    const size_t npix = (size_t)ny * nx;
    const bool want_f95 = spectral && ny == nx;
    int rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        GenWork w;
        if ((rc = gen_carve(ctx, tc, ny, nx, &w))) return rc;
        const float* s0 = stack + t0 * npix;
        float* map_b = out_psd ? out_psd + t0 * npix : nullptr;
        if (want_f95 && !map_b) {
            void* p = nullptr;
            if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * npix * tc, &p))) return rc;
            map_b = static_cast<float*>(p);
        }
        if ((rc = b4d_frame_reductions_nolock(ctx, s0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, w.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), s0, w.fr, tc, ny, nx, w.A, w.Bf))) return rc;
        if ((rc = gen_epilogue(ctx, w, tc, ny, nx, (sub_mean || zero_dc) ? 1 : 0, scale_factor, map_b, nullptr, 0, spectral != nullptr))) return rc;
        if (spectral) {
            double* tab = spectral + t0 * B4D_SP_NCOLS;
            spec_finalize_kernel<<<(unsigned)((tc + 127) / 128), 128, 0, ctx->stream>>>(w.spp, w.nblk, tab, tc);
            B4D_LAUNCH_CHECK(ctx);
            if (want_f95 && (rc = run_f95(ctx, map_b, ny, tc, tab))) return rc;
        }
    }
    return B4D_OK;
}

int gen_autocorr2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, int use_norm, double norm_mult,
                   int remove_mean, float* out_ac, double fraction, double* grain_out) {
    const int64_t B = gen_batch(ny, nx);
    const size_t npix = (size_t)ny * nx;
    int rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        GenWork w;
        if ((rc = gen_carve(ctx, tc, ny, nx, &w))) return rc;
        const float* s0 = stack + t0 * npix;
        float* o = out_ac ? out_ac + t0 * npix : nullptr;
        if (!o) {
            void* p = nullptr;
            if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * npix * tc, &p))) return rc;
            o = static_cast<float*>(p);
        }
        if ((rc = b4d_frame_reductions_nolock(ctx, s0, tc, ny, nx, nullptr, nullptr, nan(""), 0.0, w.fr))) return rc;
        if ((rc = gen_forward(ctx, gen_cache(ctx), s0, w.fr, tc, ny, nx, w.A, w.Bf))) return rc;
        if ((rc = gen_epilogue(ctx, w, tc, ny, nx, remove_mean ? 1 : 0, 1.f, nullptr, nullptr, 1, false))) return rc;
        if ((rc = gen_inverse(ctx, gen_cache(ctx), tc, ny, nx, w.A, w.Bf))) return rc;
        {
            ProfScope ps(ctx, KC_GENERIC);
            gen_autocorr_out_kernel<<<dim3(w.nblk, (unsigned)tc), 256, 0, ctx->stream>>>(w.A, ny, nx, use_norm, norm_mult,
                                                                                        1.0 / ((double)nx * (double)ny), o);
            B4D_LAUNCH_CHECK(ctx);
        }
        if (grain_out) {
            if ((rc = ensure_theta(ctx))) return rc;
            gen_argmax_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(o, (int64_t)npix, w.pk);
            B4D_LAUNCH_CHECK(ctx);
            ProfScope ps(ctx, KC_GRAIN);
            grain_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(o, ny, w.pk, ctx->fft->theta, fraction, grain_out + t0 * 4);
            B4D_LAUNCH_CHECK(ctx);
        }
    }
    return B4D_OK;
}

// The stack pipeline for frame sides that are not powers of two: the same outputs, composed from the stand-alone
// chirp-z paths (three forward transforms per frame instead of one shared; this is the compatibility path, the fused
// kernels above are the hot one).
int gen_stack_pipeline(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain, const float* dark,
                       double sat_value, double zero_eps, float psd_scale, int subpixel, double eps, double q_lo, double q_hi,
                       double* fr_out, float* quant_out, int64_t* nvalid_out, float* psd_out, float* ac_out, double* grain_out,
                       double* track_out) {
    const size_t npix = (size_t)ny * nx;
    int rc;
    int64_t B = ((int64_t)1024 << 20) / (int64_t)(npix * 4);
    if (B < 1) B = 1;
    if (B > 32768) B = 32768;
    if (!gain) B = n_frames;                            // nothing to materialise: the sub-paths batch on their own
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        const float* s0 = stack + t0 * npix;
        if (gain) {
            void* p = nullptr;
            if ((rc = b4d_scratch(ctx, SCR_GEN, sizeof(float) * npix * tc, &p))) return rc;
            int bx = (int)((npix + 2047) / 2048);
            if (bx > 592) bx = 592;
            gen_apply_gain_kernel<<<dim3(bx, (unsigned)tc), 256, 0, ctx->stream>>>(s0, gain, dark, (int64_t)npix, static_cast<float*>(p));
            B4D_LAUNCH_CHECK(ctx);
            s0 = static_cast<const float*>(p);
        }
        if (fr_out || quant_out) {
            double* frp = fr_out ? fr_out + t0 * B4D_FR_NCOLS : nullptr;
            if (!frp) {
                void* p = nullptr;
                if ((rc = b4d_scratch(ctx, SCR_SPEC_C, sizeof(double) * B4D_FR_NCOLS * tc, &p))) return rc;
                frp = static_cast<double*>(p);
            }
            FrTails tl = {q_lo, q_hi, quant_out ? quant_out + 4 * t0 : nullptr, quant_out ? nvalid_out + t0 : nullptr};
            if ((rc = b4d_frame_reductions_ex(ctx, s0, tc, ny, nx, nullptr, nullptr, sat_value, zero_eps, frp, quant_out ? &tl : nullptr, nullptr)))
                return rc;
        }
        if (psd_out && (rc = gen_psd2d(ctx, s0, tc, ny, nx, psd_scale, 0, 0, psd_out + t0 * npix, nullptr))) return rc;
        if ((ac_out || grain_out) &&
            (rc = gen_autocorr2d(ctx, s0, tc, ny, nx, 1, 1.0, 1, ac_out ? ac_out + t0 * npix : nullptr, 0.36787944117144233,
                                 grain_out ? grain_out + t0 * 4 : nullptr)))
            return rc;
        if (track_out && (rc = gen_phase_track(ctx, s0, tc, ny, nx, subpixel, eps, track_out + t0 * 4))) return rc;
    }
    return B4D_OK;
}

}  // namespace

void pipe_graphs_release(b4d_ctx* ctx);

void b4d_fft_release(b4d_ctx* ctx) {
    pipe_graphs_release(ctx);
    if (!ctx->fft) return;
    for (auto& b : ctx->fft->ref_pool) {
        if (b.ref) cudaFree(b.ref);
        if (b.ref_nyq) cudaFree(b.ref_nyq);
        if (b.gref) cudaFree(b.gref);
    }
    ctx->fft->ref_pool.clear();
    gen_release(static_cast<GenCache*>(ctx->fft->gen));
    if (ctx->fft->gref) cudaFree(ctx->fft->gref);
    for (int i = 0; i < 6; ++i) if (ctx->fft->twb[i]) cudaFree(ctx->fft->twb[i]);
    if (ctx->fft->ref) cudaFree(ctx->fft->ref);
    if (ctx->fft->ref_nyq) cudaFree(ctx->fft->ref_nyq);
    if (ctx->fft->theta) cudaFree(ctx->fft->theta);
    if (ctx->fft->blk_list) cudaFree(ctx->fft->blk_list);
    delete ctx->fft;
    ctx->fft = nullptr;
}

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int b4d_fft2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, float* out_c64) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!out_c64) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_fft2d: null output");
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_fft2d", stack, n_frames, ny, nx);
        return rcg ? rcg : gen_fft2d(ctx, stack, n_frames, ny, nx, out_c64);
    }
    int rc = check_fft_args(ctx, "b4d_fft2d", stack, n_frames, ny, nx);
    if (rc) return rc;
    const int64_t B = batch_frames(ctx, ny, nx, 1);
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        Work w;
        if ((rc = carve(ctx, tc, ny, nx, false, false, &w))) return rc;
        if ((rc = run_rows_fwd(ctx, stack + (size_t)t0 * ny * nx, tc, ny, nx, nullptr, nullptr, w, true))) return rc;
        ColsArgs c = cols_defaults(w, nx, true);
        c.cplx_out = reinterpret_cast<float2*>(out_c64) + (size_t)t0 * ny * nx;
        if ((rc = run_cols(ctx, c, tc, ny))) return rc;
    }
    return B4D_OK;
}

extern "C" int b4d_psd2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, float scale_factor,
                         int sub_mean, int zero_dc, float* out_psd, double* spectral) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!out_psd && !spectral) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_psd2d: nothing to compute");
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_psd2d", stack, n_frames, ny, nx);
        return rcg ? rcg : gen_psd2d(ctx, stack, n_frames, ny, nx, scale_factor, sub_mean, zero_dc, out_psd, spectral);
    }
    int rc = check_fft_args(ctx, "b4d_psd2d", stack, n_frames, ny, nx);
    if (rc) return rc;
    const bool want_f95 = spectral && ny == nx;
    float* map = out_psd;
    const int64_t B = batch_frames(ctx, ny, nx, 1);
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        Work w;
        if ((rc = carve(ctx, tc, ny, nx, false, false, &w))) return rc;
        float* map_b = map ? map + (size_t)t0 * ny * nx : nullptr;
        if (want_f95 && !map) {   // f95 needs the map: keep it in scratch
            void* p = nullptr;
            if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * (size_t)tc * ny * nx, &p))) return rc;
            map_b = static_cast<float*>(p);
        }
        if ((rc = run_rows_fwd(ctx, stack + (size_t)t0 * ny * nx, tc, ny, nx, nullptr, nullptr, w, true))) return rc;
        ColsArgs c = cols_defaults(w, nx, true);
        c.zero_dc = (sub_mean || zero_dc) ? 1 : 0;
        c.psd_out = map_b;
        c.psd_scale = scale_factor;
        c.spec_partials = spectral ? w.spp : nullptr;
        if ((rc = run_cols(ctx, c, tc, ny))) return rc;
        if (spectral) {
            double* tab = spectral + t0 * B4D_SP_NCOLS;
            spec_finalize_kernel<<<(unsigned)((tc + 127) / 128), 128, 0, ctx->stream>>>(w.spp, cols_tiles(ny, nx, cols_wide_out(c)), tab, tc);
            B4D_LAUNCH_CHECK(ctx);
            if (want_f95 && (rc = run_f95(ctx, map_b, ny, tc, tab))) return rc;
        }
    }
    return B4D_OK;
}

namespace {

// autocorrelation of a batch already in HBM; out_ac may live in scratch
int autocorr_batch(b4d_ctx* ctx, const float* stack, int64_t tc, int ny, int nx, int use_norm, double norm_mult,
                   float* out_ac, double fraction, double* grain_out, int zero_dc) {
    Work w;
    int rc;
    if ((rc = carve(ctx, tc, ny, nx, true, false, &w))) return rc;
    if ((rc = run_rows_fwd(ctx, stack, tc, ny, nx, nullptr, nullptr, w, true))) return rc;
    ColsArgs c = cols_defaults(w, nx, true);
    c.zero_dc = zero_dc;
    c.i2_ac = w.I2a;
    c.i2_ac_nyq = w.I2nyq;
    c.ac_partials = w.acp;
    if ((rc = run_cols(ctx, c, tc, ny))) return rc;
    RowsInvAcArgs r;
    memset(&r, 0, sizeof(r));
    r.Iz = w.I2a; r.Inyq = w.I2nyq; r.ny = ny; r.ch_log2 = log2i(cols_cw(ny, cols_wide_out(c)) / 2); r.out = out_ac;
    r.norm = use_norm ? w.acp : nullptr; r.n_norm = cols_tiles(ny, nx, cols_wide_out(c)); r.norm_mult = norm_mult;
    r.scale = 1.0 / ((double)nx * (double)ny);
    r.best = grain_out ? w.bestA : nullptr;
    int nblk = 0;
    if ((rc = run_rows_inv_ac(ctx, r, tc, nx, &nblk))) return rc;
    if (grain_out) {
        if ((rc = ensure_theta(ctx))) return rc;
        argmax_reduce_kernel<<<(unsigned)tc, 128, 0, ctx->stream>>>(w.bestA, nblk, w.pk_idx, w.pk_val);
        B4D_LAUNCH_CHECK(ctx);
        ProfScope ps(ctx, KC_GRAIN);
        grain_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(out_ac, ny, w.pk_idx, ctx->fft->theta, fraction, grain_out);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

}  // namespace

extern "C" int b4d_autocorr2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, int remove_mean,
                              int standardize, int normalize_peak, float* out_ac, double fraction, double* grain_out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    const bool generic = !pow2_sides(ny, nx);
    int rc = generic ? check_gen_args(ctx, "b4d_autocorr2d", stack, n_frames, ny, nx)
                     : check_fft_args(ctx, "b4d_autocorr2d", stack, n_frames, ny, nx);
    if (rc) return rc;
    if (!out_ac && !grain_out) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_autocorr2d: nothing to compute");
    if (grain_out && ny != nx) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_autocorr2d: grain widths need a square frame (pad_to_square first)");
    if (grain_out && !(fraction > 0.0 && fraction < 1.0)) return b4d_fail(ctx, B4D_ERR_INVALID, "fraction must be in (0, 1).");
    if (standardize && !normalize_peak && !remove_mean)
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_autocorr2d: standardize=1 with remove_mean=0 and normalize='none' is not built");
    // Peak normalisation cancels the /std of `standardize` exactly. Without it, the zero lag of the
    // standardised map is n = ny*nx (n var / var), i.e. the peak-normalised map times n.
    const int use_norm = normalize_peak || standardize;
    const double norm_mult = (standardize && !normalize_peak) ? (double)ny * (double)nx : 1.0;
    if (generic) return gen_autocorr2d(ctx, stack, n_frames, ny, nx, use_norm, norm_mult, remove_mean ? 1 : 0, out_ac, fraction, grain_out);
    const int64_t B = batch_frames(ctx, ny, nx, 2);
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        float* o = out_ac ? out_ac + (size_t)t0 * ny * nx : nullptr;
        if (!o) {
            void* p = nullptr;
            if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * (size_t)tc * ny * nx, &p))) return rc;
            o = static_cast<float*>(p);
        }
        rc = autocorr_batch(ctx, stack + (size_t)t0 * ny * nx, tc, ny, nx, use_norm, norm_mult, o, fraction,
                            grain_out ? grain_out + t0 * 4 : nullptr, remove_mean ? 1 : 0);
        if (rc) return rc;
    }
    return B4D_OK;
}

extern "C" int b4d_xcorr2d(b4d_ctx* ctx, const float* a, const float* b, int64_t n_frames, int ny, int nx, int remove_mean,
                           int standardize, int normalize_peak, float* out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!b || !out) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_xcorr2d: null pointer");
    if (standardize && !normalize_peak)
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_xcorr2d: standardize=1 with normalize='none' is applied by the host wrapper");
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_xcorr2d", a, n_frames, ny, nx);
        return rcg ? rcg : gen_xcorr2d(ctx, a, b, n_frames, ny, nx, remove_mean, normalize_peak, out);
    }
    int rc = check_fft_args(ctx, "b4d_xcorr2d", a, n_frames, ny, nx);
    if (rc) return rc;
    const int64_t B = batch_frames(ctx, ny, nx, 3);
    const size_t per = (size_t)ny * (nx / 2);
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        Work w;
        if ((rc = carve(ctx, tc, ny, nx, true, true, &w))) return rc;
        // conj spectrum of b -> I2b (blocked) + Nyquist columns in scratch
        void* p = nullptr;
        if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float2) * (size_t)tc * ny + sizeof(double) * 2 * B4D_FR_NCOLS * tc, &p))) return rc;
        float2* bnyq = static_cast<float2*>(p);
        if ((rc = run_rows_fwd(ctx, b + (size_t)t0 * ny * nx, tc, ny, nx, nullptr, nullptr, w, true))) return rc;
        ColsArgs cb = cols_defaults(w, nx, true);
        cb.zero_dc = remove_mean ? 1 : 0;
        cb.conj_out = w.I2b;
        cb.conj_nyq_out = bnyq;
        if ((rc = run_cols(ctx, cb, tc, ny))) return rc;
        // spectrum of a times conj spectrum of b, inverse along y -> I2a
        if ((rc = run_rows_fwd(ctx, a + (size_t)t0 * ny * nx, tc, ny, nx, nullptr, nullptr, w, true))) return rc;
        ColsArgs ca = cols_defaults(w, nx, true);
        ca.zero_dc = remove_mean ? 1 : 0;
        ca.i2_pc = w.I2a;
        ca.R = w.I2b; ca.Rnyq = bnyq; ca.r_stride = per; ca.rnyq_stride = (size_t)ny;
        ca.whiten = 0; ca.eps = 0.f;
        if ((rc = run_cols(ctx, ca, tc, ny))) return rc;
        RowsInvArgs r;
        memset(&r, 0, sizeof(r));
        float* o = out + (size_t)t0 * ny * nx;
        r.Ia = w.I2a; r.ny = ny; r.outA = o; r.kindA = 0;
        r.scaleA = 1.0 / ((double)nx * (double)ny);
        if ((rc = run_rows_inv(ctx, r, tc, nx))) return rc;
        if (standardize && !normalize_peak)
            return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_xcorr2d: standardize=1 with normalize='none' is applied by the host wrapper");
        if (normalize_peak) {
            absmax_kernel<<<(unsigned)tc, 256, 0, ctx->stream>>>(o, (int64_t)ny * nx, w.pk_val);
            B4D_LAUNCH_CHECK(ctx);
            scale_by_kernel<<<dim3(64, (unsigned)tc), 256, 0, ctx->stream>>>(o, (int64_t)ny * nx, w.pk_val, 1);
            B4D_LAUNCH_CHECK(ctx);
        }
    }
    return B4D_OK;
}

// A tracker reference owned by its creator (b4d_phase_reference_create): the conjugate spectrum of the z-scored,
// embedded template, in the layout the column pass reads (powers of two) or in natural order (any other size).
void pipe_graphs_release(b4d_ctx* ctx);

struct b4d_ref {
    float2* ref = nullptr;
    float2* ref_nyq = nullptr;
    float2* gref = nullptr;
    size_t gref_elems = 0;
    int ny = 0, nx = 0;
};

namespace {
// Installs a reference's buffers as the context's current reference for the duration of one (locked) call.
struct RefSwap {
    FftPlanCache* f;
    float2 *ref, *ref_nyq, *gref;
    size_t gref_elems;
    int ny, nx;
    RefSwap(b4d_ctx* ctx, const b4d_ref* r) {
        if (!ctx->fft) ctx->fft = new FftPlanCache();
        f = ctx->fft;
        ref = f->ref; ref_nyq = f->ref_nyq; gref = f->gref; gref_elems = f->gref_elems; ny = f->ref_ny; nx = f->ref_nx;
        if (r) { f->ref = r->ref; f->ref_nyq = r->ref_nyq; f->gref = r->gref; f->gref_elems = r->gref_elems; f->ref_ny = r->ny; f->ref_nx = r->nx; }
        else { f->ref = nullptr; f->ref_nyq = nullptr; f->gref = nullptr; f->gref_elems = 0; f->ref_ny = 0; f->ref_nx = 0; }
    }
    ~RefSwap() { f->ref = ref; f->ref_nyq = ref_nyq; f->gref = gref; f->gref_elems = gref_elems; f->ref_ny = ny; f->ref_nx = nx; }
};
}  // namespace

int phase_set_reference_body(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx, int y0, int x0, double eps);

extern "C" int b4d_phase_set_reference(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx, int y0, int x0,
                                       double eps) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    return phase_set_reference_body(ctx, tpl, h, w, ny, nx, y0, x0, eps);
}

extern "C" int b4d_phase_reference_create(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx, int y0, int x0,
                                          double eps, b4d_ref** out) {
    if (!ctx || !out) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    *out = nullptr;
    int rc;
    b4d_ref* r = new b4d_ref();
    {
        // the context's own reference is put aside; buffers of a released reference of the same shape are reused,
        // otherwise the body allocates fresh ones
        b4d_ref seed;
        if (ctx->fft) {
            auto& pool = ctx->fft->ref_pool;
            for (size_t i = 0; i < pool.size(); ++i)
                if (pool[i].ny == ny && pool[i].nx == nx) {
                    seed.ref = pool[i].ref; seed.ref_nyq = pool[i].ref_nyq; seed.gref = pool[i].gref;
                    seed.gref_elems = pool[i].gref_elems; seed.ny = ny; seed.nx = nx;
                    pool.erase(pool.begin() + i);
                    break;
                }
        }
        RefSwap sw(ctx, seed.ny ? &seed : nullptr);
        rc = phase_set_reference_body(ctx, tpl, h, w, ny, nx, y0, x0, eps);
        FftPlanCache* f = ctx->fft;
        r->ref = f->ref; r->ref_nyq = f->ref_nyq; r->gref = f->gref; r->gref_elems = f->gref_elems; r->ny = f->ref_ny; r->nx = f->ref_nx;
    }
    if (rc) {
        if (r->ref) cudaFree(r->ref);
        if (r->ref_nyq) cudaFree(r->ref_nyq);
        if (r->gref) cudaFree(r->gref);
        delete r;
        return rc;
    }
    *out = r;
    return B4D_OK;
}

extern "C" int b4d_phase_reference_destroy(b4d_ctx* ctx, b4d_ref* r) {
    if (!ctx) return B4D_ERR_INVALID;
    if (!r) return B4D_OK;
    B4dCall g(ctx);
    cudaDeviceSynchronize();                          // (launches that read it may be in flight on any of the context's streams)
    RefBuffers b;
    b.ref = r->ref; b.ref_nyq = r->ref_nyq; b.gref = r->gref; b.gref_elems = r->gref_elems; b.ny = r->ny; b.nx = r->nx;
    delete r;
    if (!ctx->fft) ctx->fft = new FftPlanCache();
    auto& pool = ctx->fft->ref_pool;
    pool.push_back(b);
    if (pool.size() > 4) {                            // really freed: captured batches may carry its pointers
        pipe_graphs_release(ctx);
        RefBuffers old = pool.front();
        pool.erase(pool.begin());
        if (old.ref) cudaFree(old.ref);
        if (old.ref_nyq) cudaFree(old.ref_nyq);
        if (old.gref) cudaFree(old.gref);
    }
    return B4D_OK;
}

int phase_set_reference_body(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx, int y0, int x0, double eps) {
    if (h < 1 || w < 1 || y0 < 0 || x0 < 0 || y0 + h > ny || x0 + w > nx)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_phase_set_reference: template does not fit the frame");
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_phase_set_reference", tpl, 1, ny, nx);
        return rcg ? rcg : gen_phase_set_reference(ctx, tpl, h, w, ny, nx, y0, x0, eps);
    }
    int rc = check_fft_args(ctx, "b4d_phase_set_reference", tpl, 1, ny, nx);
    if (rc) return rc;
    if (!ctx->fft) ctx->fft = new FftPlanCache();
    FftPlanCache* f = ctx->fft;
    if (f->ref_ny != ny || f->ref_nx != nx) {
        if (f->ref) { B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(f->ref); cudaFree(f->ref_nyq); f->ref = nullptr; f->ref_nyq = nullptr; }
        B4D_CUDA(ctx, cudaMalloc(&f->ref, sizeof(float2) * (size_t)ny * (nx / 2)));
        B4D_CUDA(ctx, cudaMalloc(&f->ref_nyq, sizeof(float2) * (size_t)ny));
        f->ref_ny = ny; f->ref_nx = nx;
    }
    void* p = nullptr;
    if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * (size_t)ny * nx + sizeof(double) * B4D_FR_NCOLS, &p))) return rc;
    float* padded = static_cast<float*>(p);
    double* tfr = reinterpret_cast<double*>(padded + (size_t)ny * nx);
    if ((rc = b4d_frame_reductions_nolock(ctx, tpl, 1, h, w, nullptr, nullptr, nan(""), 0.0, tfr))) return rc;
    embed_template_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(tpl, h, w, ny, nx, y0, x0, (float)eps, tfr, padded);
    B4D_LAUNCH_CHECK(ctx);
    Work wk;
    if ((rc = carve(ctx, 1, ny, nx, false, false, &wk))) return rc;
    if ((rc = run_rows_fwd(ctx, padded, 1, ny, nx, nullptr, nullptr, wk, false))) return rc;
    ColsArgs c = cols_defaults(wk, nx, false);
    c.conj_out = f->ref;
    c.conj_nyq_out = f->ref_nyq;
    return run_cols(ctx, c, 1, ny);
}

namespace {

int track_finish(b4d_ctx* ctx, Work& w, const float* mag, int64_t tc, int ny, int nx, int nblk, int subpixel, double eps,
                 double* out) {
    argmax_reduce_kernel<<<(unsigned)tc, 128, 0, ctx->stream>>>(w.bestB, nblk, w.pk_idx, w.pk_val);
    B4D_LAUNCH_CHECK(ctx);
    void* p = nullptr;
    int rc = b4d_scratch(ctx, SCR_MISC, 1024, &p);   // already sized by carve(); first 512 B hold the quantile
    if (rc) return rc;
    static const double half = 0.5;
    if ((rc = b4d_put_doubles(ctx, static_cast<double*>(p), &half, 1))) return rc;
    if ((rc = b4d_select_impl(ctx, mag, tc, (int64_t)ny * nx, static_cast<const double*>(p), 1, 1, w.med,
                              reinterpret_cast<int64_t*>(w.nvalid))))
        return rc;
    phase_finalize_kernel<<<(unsigned)((tc + 63) / 64), 64, 0, ctx->stream>>>(mag, w.pk_idx, ny, nx, w.med, w.nvalid, subpixel,
                                                                              eps, out, tc);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

}  // namespace

int b4d_fused_median_begin(b4d_ctx* ctx, int64_t T, int regions, FusedMedian* fm);
int b4d_fused_median_bracket(b4d_ctx* ctx, const FusedMedian& fm, const float* samples, int m, int64_t T);
int b4d_fused_median_final(b4d_ctx* ctx, const FusedMedian& fm, int64_t T, float* out, int64_t* nvalid);

namespace {

// number of sample row blocks for the fused median (0: the frame is too small, use the map-based path)
int fused_sample_blocks(int ny, int nx) {
    const int nblk = rows_inv_blocks(nx, ny), rpc = ny / nblk;
    if ((int64_t)ny * nx < 65536 || nblk < 4) return 0;
    int ns = 32768 / (rpc * nx);
    if (ns > nblk / 2) ns = nblk / 2;
    return ns < 1 ? 0 : ns;
}

// Inverse rows + tracker results with the |corr| map never written (fused median), in three steps so that the row
// passes can follow the column pass a few frames at a time:
//   begin: scratch of the whole batch (tc frames), sample-block list;
//   range: frames [t0, t0 + tcs) of the batch -- sample rows, bracket, census of every other row block, peak, and the
//          three row blocks around the peak once more for the 3x3 neighbourhood. `r` is set up for a full launch over the
//          range's frames (r.Ia = the range's first frame; map A = the |.| map, argmax partials in w.bestB, batch-wide);
//   end:   exact medians from the candidates, results of the whole batch.
struct TrackFused {
    FusedMedian fm;
    int ns, nblk, rpc, m;
    float* samples;     // (tc, m)
    float* window;      // (tc, 3 * rpc, nx)
    int* blk3;          // (tc, 3)
};

int track_fused_begin(b4d_ctx* ctx, int64_t tc, int ny, int nx, int ns, float* scratch, TrackFused* tf) {
    const int nblk = rows_inv_blocks(nx, ny), rpc = ny / nblk;
    tf->ns = ns; tf->nblk = nblk; tf->rpc = rpc; tf->m = ns * rpc * nx;
    FftPlanCache* f = ctx->fft;
    if (f->blk_n != nblk || f->blk_ns != ns) {
        std::vector<int> h;
        std::vector<char> used(nblk, 0);
        for (int i = 0; i < ns; ++i) { const int b = (int)(((int64_t)(2 * i + 1) * nblk) / (2 * ns)); h.push_back(b); used[b] = 1; }
        for (int b = 0; b < nblk; ++b) if (!used[b]) h.push_back(b);
        if (f->blk_list) { B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(f->blk_list); f->blk_list = nullptr; }
        B4D_CUDA(ctx, cudaMalloc(&f->blk_list, sizeof(int) * nblk));
        B4D_CUDA(ctx, cudaMemcpy(f->blk_list, h.data(), sizeof(int) * nblk, cudaMemcpyHostToDevice));
        f->blk_n = nblk; f->blk_ns = ns;
    }
    tf->samples = scratch;
    tf->window = tf->samples + (size_t)tc * tf->m;
    tf->blk3 = reinterpret_cast<int*>(tf->window + (size_t)tc * 3 * rpc * nx);
    return b4d_fused_median_begin(ctx, tc, nblk, &tf->fm);
}

int track_fused_range(b4d_ctx* ctx, const Work& w, const TrackFused& tf, RowsInvArgs r, int64_t t0, int64_t tcs, int nx) {
    const int nblk = tf.nblk, rpc = tf.rpc, ns = tf.ns, m = tf.m;
    FftPlanCache* f = ctx->fft;
    FusedMedian fm = tf.fm;
    fm.st += t0; fm.need += t0; fm.bhist += (size_t)t0 * SEL_BINS; fm.cnt3 += (size_t)t0 * fm.regions * 3;
    fm.cand += (size_t)t0 * ((size_t)fm.regions * FM_REGION + FM_SAMPLE_CAP);
    float* samples = tf.samples + (size_t)t0 * m;
    float* window = tf.window + (size_t)t0 * 3 * rpc * nx;
    int* blk3 = tf.blk3 + t0 * 3;
    r.bestA = w.bestB + (size_t)t0 * nblk;
    int rc;
    // 1. sample rows: |.| rows into the compact sample buffer
    RowsInvArgs r1 = r;
    r1.blk_map = f->blk_list; r1.mag_mode = 1;
    r1.outA = samples;
    if ((rc = run_rows_inv(ctx, r1, tcs, nx, ns))) return rc;
    // 2. bracket around the median + census of the sample rows
    if ((rc = b4d_fused_median_bracket(ctx, fm, samples, m, tcs))) return rc;
    // 3. every other row block: census in the epilogue, no |.| map
    RowsInvArgs r2 = r;
    r2.blk_map = f->blk_list + ns; r2.mag_mode = 2;
    r2.sel = fm.st; r2.cand = fm.cand; r2.cnt3 = fm.cnt3; r2.bhist = fm.bhist; r2.regions = fm.regions;
    if ((rc = run_rows_inv(ctx, r2, tcs, nx, nblk - ns))) return rc;
    // 4. peak, then the three row blocks around it once more for the 3x3 neighbourhood
    argmax_reduce_kernel<<<(unsigned)tcs, 128, 0, ctx->stream>>>(r.bestA, nblk, w.pk_idx + t0, w.pk_val + t0);
    B4D_LAUNCH_CHECK(ctx);
    window_blocks_kernel<<<(unsigned)((tcs + 63) / 64), 64, 0, ctx->stream>>>(w.pk_idx + t0, r.ny, nx, rpc, blk3, tcs);
    B4D_LAUNCH_CHECK(ctx);
    RowsInvArgs r3 = r;
    r3.blk_map_pf = blk3; r3.mag_mode = 1; r3.bestA = nullptr;
    r3.outA = window;
    return run_rows_inv(ctx, r3, tcs, nx, 3);
}

int track_fused_end(b4d_ctx* ctx, const Work& w, const TrackFused& tf, int64_t tc, int ny, int nx, int subpixel, double eps,
                    double* out) {
    // 5. exact median from the candidates, results
    int rc;
    if ((rc = b4d_fused_median_final(ctx, tf.fm, tc, w.med, reinterpret_cast<int64_t*>(w.nvalid)))) return rc;
    phase_finalize_window_kernel<<<(unsigned)((tc + 63) / 64), 64, 0, ctx->stream>>>(tf.window, tf.blk3, tf.rpc, w.pk_idx, ny, nx, w.med,
                                                                                     w.nvalid, tf.fm.need, subpixel, eps, out, tc);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

int rows_inv_track_fused(b4d_ctx* ctx, Work& w, RowsInvArgs r, int64_t tc, int ny, int nx, int ns, float* scratch,
                         int subpixel, double eps, double* out) {
    TrackFused tf;
    int rc;
    if ((rc = track_fused_begin(ctx, tc, ny, nx, ns, scratch, &tf))) return rc;
    if ((rc = track_fused_range(ctx, w, tf, r, 0, tc, nx))) return rc;
    return track_fused_end(ctx, w, tf, tc, ny, nx, subpixel, eps, out);
}

size_t fused_scratch_floats(int ny, int nx, int ns, int64_t tc) {
    const int nblk = rows_inv_blocks(nx, ny), rpc = ny / nblk;
    return (size_t)tc * ((size_t)ns * rpc * nx + (size_t)3 * rpc * nx + 4);
}

}  // namespace

int phase_track_body(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, int subpixel, double eps, double* out);

extern "C" int b4d_phase_track(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, int subpixel, double eps,
                               double* out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    return phase_track_body(ctx, stack, n_frames, ny, nx, subpixel, eps, out);
}

extern "C" int b4d_phase_track_ref(b4d_ctx* ctx, const b4d_ref* ref, const float* stack, int64_t n_frames, int ny, int nx,
                                   int subpixel, double eps, int map_median, double* out) {
    if (!ctx || !ref) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    RefSwap sw(ctx, ref);
    const bool fm = ctx->fused_median;
    if (map_median) ctx->fused_median = false;
    const int rc = phase_track_body(ctx, stack, n_frames, ny, nx, subpixel, eps, out);
    ctx->fused_median = fm;
    return rc;
}

int phase_track_body(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, int subpixel, double eps, double* out) {
    if (!out) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_phase_track: null output");
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_phase_track", stack, n_frames, ny, nx);
        if (rcg) return rcg;
        if (!ctx->fft || !ctx->fft->gref || ctx->fft->ref_ny != ny || ctx->fft->ref_nx != nx || ctx->fft->gref_elems != (size_t)ny * nx)
            return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_phase_track: call b4d_phase_set_reference for (%d, %d) frames first", ny, nx);
        return gen_phase_track(ctx, stack, n_frames, ny, nx, subpixel, eps, out);
    }
    int rc = check_fft_args(ctx, "b4d_phase_track", stack, n_frames, ny, nx);
    if (rc) return rc;
    if (!ctx->fft || !ctx->fft->ref || ctx->fft->ref_ny != ny || ctx->fft->ref_nx != nx)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_phase_track: call b4d_phase_set_reference for (%d, %d) frames first", ny, nx);
    const int64_t B = batch_frames(ctx, ny, nx, 4);
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        const float* s0 = stack + (size_t)t0 * ny * nx;
        Work w;
        if ((rc = carve(ctx, tc, ny, nx, false, true, &w))) return rc;
        const int ns = ctx->fused_median ? fused_sample_blocks(ny, nx) : 0;
        void* p = nullptr;
        const size_t map_floats = ns ? fused_scratch_floats(ny, nx, ns, tc) : (size_t)tc * ny * nx;
        if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(double) * B4D_FR_NCOLS * tc + sizeof(float) * map_floats + 256, &p))) return rc;
        double* fr = static_cast<double*>(p);
        float* mag = reinterpret_cast<float*>(fr + (size_t)B4D_FR_NCOLS * tc);
        // mean and standard deviation for the z-score come out of the forward row pass (no separate reduction pass)
        if ((rc = run_rows_fwd(ctx, s0, tc, ny, nx, nullptr, nullptr, w, true, false, fr))) return rc;
        ColsArgs c = cols_defaults(w, nx, true);
        c.i2_pc = w.I2b;
        c.R = ctx->fft->ref; c.Rnyq = ctx->fft->ref_nyq; c.r_stride = 0; c.rnyq_stride = 0;
        c.whiten = 1; c.eps = (float)eps;
        c.fr = fr; c.fr_stride = B4D_FR_NCOLS;
        if ((rc = run_cols(ctx, c, tc, ny))) return rc;
        RowsInvArgs r;
        memset(&r, 0, sizeof(r));
        r.Ia = w.I2b; r.ny = ny; r.outA = mag; r.kindA = 1;
        r.scaleA = 1.0 / ((double)nx * (double)ny);
        r.bestA = w.bestB;
        if (ns) {
            if ((rc = rows_inv_track_fused(ctx, w, r, tc, ny, nx, ns, mag, subpixel, eps, out + t0 * 4))) return rc;
        } else {
            if ((rc = run_rows_inv(ctx, r, tc, nx))) return rc;
            if ((rc = track_finish(ctx, w, mag, tc, ny, nx, rows_inv_blocks(nx, ny), subpixel, eps, out + t0 * 4))) return rc;
        }
    }
    return B4D_OK;
}

// side stream + fork / join events of the fused pipeline (experiment knob: B4D_SIDE_STREAM=0 keeps one stream)
bool side_stream_ready(b4d_ctx* ctx) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("B4D_SIDE_STREAM"); on = (e && atoi(e) == 0) ? 0 : 1; }
    if (!on) return false;
    if (!ctx->side) {
        int lo = 0, hi = 0;                             // highest priority: its CTAs are placed first whenever an SM has room
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, hi) != cudaSuccess) { ctx->side = nullptr; return false; }
        if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) return false;
    }
    return ctx->ev_fork && ctx->ev_join;
}

// -------------------------------------------------------------------------------------------------
// Frame-pipelined schedule of the fused stack pipeline ("lanes")
// -------------------------------------------------------------------------------------------------
// A batch of 128 frames pushes 3.5 frames' worth of row <-> column intermediates per frame through HBM (2 - 4 GB per
// batch against 126 MB of L2). Here the five big kernels run `sub` frames at a time instead, in the order
//     reduce(s) -> rows_fwd(s) -> cols(s)                         on the caller's stream
//                                  cols(s) -> rows_inv_ac(s)      on side stream A
//                                  cols(s) -> tracker rows(s)     on side stream S (sample rows, bracket, census, window)
// so that (a) the forward row pass finds the frames the reduction pass has just streamed in L2, (b) the column pass finds
// the half spectra of the forward row pass there, (c) the two inverse row passes find the column pass's output there.
// The intermediates live in ring slots of `sub` frames (one slot for the forward half spectra, `slots` for each inverse
// branch) that are overwritten while still in L2; the per-frame kernels that never touch them (tail percentiles, final
// medians, 3x3 fit, autocorrelation argmax, grain widths) run once per batch after the join, as before.
struct Sched { int sub, lanes, slots, keep; };

Sched pipeline_sched(b4d_ctx* ctx) {
    static int e_sub = -2, e_lanes = -2, e_slots = -2, e_keep = -2;
    if (e_sub == -2) {
        const char* e = getenv("B4D_SUB"); e_sub = e ? atoi(e) : -1;
        e = getenv("B4D_LANES"); e_lanes = e ? atoi(e) : -1;
        e = getenv("B4D_SLOTS"); e_slots = e ? atoi(e) : -1;
        e = getenv("B4D_KEEP"); e_keep = e ? atoi(e) : -1;
    }
    Sched sc;
    sc.sub = ctx->sched_sub >= 0 ? ctx->sched_sub : (e_sub >= 0 ? e_sub : B4D_SCHED_SUB_DEFAULT);
    sc.lanes = ctx->sched_lanes >= 1 ? ctx->sched_lanes : (e_lanes >= 1 ? e_lanes : 2);
    sc.slots = ctx->sched_slots >= 1 ? ctx->sched_slots : (e_slots >= 1 ? e_slots : 1);
    sc.keep = ctx->sched_keep >= 0 ? ctx->sched_keep : (e_keep >= 0 ? e_keep : 1);
    if (sc.slots > 8) sc.slots = 8;
    if (sc.lanes > 8) sc.lanes = 8;
    return sc;
}

// streams: lane l owns lane_streams[3 l .. 3 l + 2] = (columns, autocorrelation rows, tracker rows); lane 0's first
// stream is the caller's. events: per lane 3 slots-wide rings (columns done, autocorrelation rows done, tracker rows
// done) + 3 joins, then one fork event.
bool lanes_ready(b4d_ctx* ctx, int lanes, int slots) {
    const size_t ns = 3 * (size_t)lanes;
    while (ctx->lane_streams.size() < ns) {
        cudaStream_t st;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        // the inverse row passes drain intermediates: they go first whenever an SM has room
        const int prio = (ctx->lane_streams.size() % 3) ? hi : lo;
        if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio) != cudaSuccess) return false;
        ctx->lane_streams.push_back(st);
    }
    const size_t need = (size_t)lanes * (3 * (size_t)slots + 3) + 1;
    while (ctx->lane_ev.size() < need) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
        ctx->lane_ev.push_back(e);
    }
    return true;
}

int pipeline_batch_lanes(b4d_ctx* ctx, const Sched& sc, const float* s0, int64_t tc, int ny, int nx, const float* gain,
                         const float* dark, double sat_value, double zero_eps, float psd_scale, int subpixel, double eps,
                         double q_lo, double q_hi, double* fr_out, float* quant_out, int64_t* nvalid_out, float* psd_out,
                         float* ac_out, double* grain_out, double* track_out, int ns, double* spectral_out) {
    const bool want_ac = ac_out || grain_out, want_pc = track_out != nullptr;
    const size_t npix = (size_t)ny * nx, per = (size_t)ny * (nx / 2);
    const int64_t F = sc.sub;
    const int NB = sc.slots, L = sc.lanes;
    int rc;
    Work w;
    if ((rc = carve(ctx, tc, ny, nx, want_ac, want_pc, &w, F * L, F * NB * L, F * NB * L))) return rc;
    void* p = nullptr;
    const size_t mag_floats = !want_pc ? 0 : ((fused_scratch_floats(ny, nx, ns, tc) + 63) & ~size_t(63));
    const size_t ac_floats = (want_ac && !ac_out) ? npix * tc : 0;
    const bool psd_scratch = spectral_out && ny == nx && !psd_out;     // f95 reads the PSD map
    const size_t need = sizeof(float) * (mag_floats + ac_floats + (psd_scratch ? npix * tc : 0)) + sizeof(double) * B4D_FR_NCOLS * tc + 256;
    if ((rc = b4d_scratch(ctx, SCR_MAP, need, &p))) return rc;
    double* fr = static_cast<double*>(p);
    float* mag = reinterpret_cast<float*>(fr + (size_t)B4D_FR_NCOLS * tc);
    float* acm = ac_out ? ac_out : mag + mag_floats;
    if (psd_scratch) psd_out = mag + mag_floats + ac_floats;
    double* frp = fr_out ? fr_out : fr;
    const bool reduced = fr_out || want_pc || quant_out;
    if (grain_out && (rc = ensure_theta(ctx))) return rc;
    // twiddle tables are created on first use with a blocking copy: do that before anything is in flight
    { const float2* tw; if ((rc = get_twiddle_bases(ctx, nx, &tw)) || (rc = get_twiddle_bases(ctx, ny, &tw))) return rc; }

    cudaStream_t M0 = ctx->stream;
    const int EPL = 3 * NB + 3;                           // events per lane
    cudaEvent_t* ev = ctx->lane_ev.data();
    cudaEvent_t e_pre = ev[(size_t)L * EPL];
    auto Ms = [&](int l) { return l == 0 ? M0 : ctx->lane_streams[3 * l]; };
    auto As = [&](int l) { return ctx->lane_streams[3 * l + 1]; };
    auto Ss = [&](int l) { return ctx->lane_streams[3 * l + 2]; };

    FrPlan pl;
    FrTails tl = {q_lo, q_hi, quant_out, nvalid_out};
    if (reduced && (rc = b4d_fr_begin(ctx, s0, tc, ny, nx, gain, dark, sat_value, zero_eps, frp, quant_out ? &tl : nullptr, w.pilot, &pl)))
        return rc;
    TrackFused tf;
    if (want_pc && (rc = track_fused_begin(ctx, tc, ny, nx, ns, mag, &tf))) return rc;
    const int64_t n_sub = (tc + F - 1) / F;
    const int lanes_used = (int)(n_sub < L ? n_sub : L);
    B4D_CUDA(ctx, cudaEventRecord(e_pre, M0));
    for (int l = 0; l < lanes_used; ++l) {
        if (l > 0) B4D_CUDA(ctx, cudaStreamWaitEvent(Ms(l), e_pre, 0));
        if (want_ac) B4D_CUDA(ctx, cudaStreamWaitEvent(As(l), e_pre, 0));
        if (want_pc) B4D_CUDA(ctx, cudaStreamWaitEvent(Ss(l), e_pre, 0));
    }

    const int rows_cta = ny / rows_inv_blocks(nx, ny), nblk_ac = (ny / 2 + 1 + rows_cta - 1) / rows_cta;
    ctx->keep_mode = sc.keep;
    auto body = [&]() -> int {
        int r_;
        int64_t a = 0;
        for (int64_t s = 0; a < tc; ++s, a += F) {
            const int64_t f = tc - a < F ? tc - a : F;
            const int l = (int)(s % L), slot = (int)((s / L) % NB);
            cudaStream_t M = Ms(l), A = As(l), S = Ss(l);
            cudaEvent_t* e_cols = ev + (size_t)l * EPL;
            cudaEvent_t* e_ac = e_cols + NB;
            cudaEvent_t* e_tr = e_cols + 2 * NB;
            ctx->stream = M;
            if (reduced && (r_ = b4d_fr_range(ctx, pl, a, f))) return r_;
            if (!(psd_out || want_ac || want_pc)) continue;
            Work ws = w;
            ws.pilot = w.pilot + a;
            ws.H = w.H + (size_t)l * F * per;
            if ((r_ = run_rows_fwd(ctx, s0 + (size_t)a * npix, f, ny, nx, gain, dark, ws, true, reduced))) return r_;
            if (s / L >= NB) {                            // the slot's previous tenants must have been consumed
                if (want_ac) B4D_CUDA(ctx, cudaStreamWaitEvent(M, e_ac[slot], 0));
                if (want_pc) B4D_CUDA(ctx, cudaStreamWaitEvent(M, e_tr[slot], 0));
            }
            ColsArgs c = cols_defaults(ws, nx, true);
            c.psd_out = psd_out ? psd_out + (size_t)a * npix : nullptr;
            c.psd_scale = psd_scale;
            const int ntl = cols_tiles(ny, nx, cols_wide_out(c));
            float2* i2a = want_ac ? w.I2a + (size_t)(l * NB + slot) * F * per : nullptr;
            float2* i2b = want_pc ? w.I2b + (size_t)(l * NB + slot) * F * per : nullptr;
            if (want_ac) { c.i2_ac = i2a; c.i2_ac_nyq = w.I2nyq + (size_t)a * ny; c.ac_partials = w.acp + (size_t)a * ntl; }
            if (spectral_out) c.spec_partials = w.spp + (size_t)a * ntl * NSP;
            if (want_pc) {
                c.i2_pc = i2b;
                c.R = ctx->fft->ref; c.Rnyq = ctx->fft->ref_nyq; c.r_stride = 0; c.rnyq_stride = 0;
                c.whiten = 1; c.eps = (float)eps; c.fr = frp + a * B4D_FR_NCOLS; c.fr_stride = B4D_FR_NCOLS;
            }
            c.zero_dc = 0;
            c.ac_zero_dc = 1;
            if ((r_ = run_cols(ctx, c, f, ny))) return r_;
            B4D_CUDA(ctx, cudaEventRecord(e_cols[slot], M));
            if (want_ac) {
                B4D_CUDA(ctx, cudaStreamWaitEvent(A, e_cols[slot], 0));
                ctx->stream = A;
                RowsInvAcArgs ra;
                memset(&ra, 0, sizeof(ra));
                ra.Iz = i2a; ra.Inyq = w.I2nyq + (size_t)a * ny; ra.ny = ny; ra.ch_log2 = log2i(cols_cw(ny, cols_wide_out(c)) / 2);
                ra.out = acm + (size_t)a * npix; ra.norm = w.acp + (size_t)a * ntl; ra.n_norm = ntl; ra.norm_mult = 1.0;
                ra.scale = 1.0 / ((double)nx * ny);
                ra.best = grain_out ? w.bestA + (size_t)a * nblk_ac : nullptr;
                int nb_ = 0;
                r_ = run_rows_inv_ac(ctx, ra, f, nx, &nb_);
                if (!r_ && cudaEventRecord(e_ac[slot], A) != cudaSuccess) r_ = b4d_fail(ctx, B4D_ERR_CUDA, "cudaEventRecord failed");
                if (r_) return r_;
            }
            if (want_pc) {
                B4D_CUDA(ctx, cudaStreamWaitEvent(S, e_cols[slot], 0));
                ctx->stream = S;
                RowsInvArgs r;
                memset(&r, 0, sizeof(r));
                r.ny = ny; r.Ia = i2b; r.outA = nullptr; r.kindA = 1; r.scaleA = 1.0 / ((double)nx * ny); r.bestA = w.bestB;
                r_ = track_fused_range(ctx, w, tf, r, a, f, nx);
                if (!r_ && cudaEventRecord(e_tr[slot], S) != cudaSuccess) r_ = b4d_fail(ctx, B4D_ERR_CUDA, "cudaEventRecord failed");
                if (r_) return r_;
            }
        }
        return B4D_OK;
    };
    rc = body();
    ctx->stream = M0;
    ctx->keep_mode = 0;
    // joined on every path: the side streams' work is in flight and the scratch is reused
    for (int l = 0; l < lanes_used; ++l) {
        cudaEvent_t* ej = ev + (size_t)l * EPL + 3 * NB;
        bool ok = true;
        if (l > 0) ok = ok && cudaEventRecord(ej[0], Ms(l)) == cudaSuccess && cudaStreamWaitEvent(M0, ej[0], 0) == cudaSuccess;
        if (want_ac) ok = ok && cudaEventRecord(ej[1], As(l)) == cudaSuccess && cudaStreamWaitEvent(M0, ej[1], 0) == cudaSuccess;
        if (want_pc) ok = ok && cudaEventRecord(ej[2], Ss(l)) == cudaSuccess && cudaStreamWaitEvent(M0, ej[2], 0) == cudaSuccess;
        if (!ok && !rc) rc = b4d_fail(ctx, B4D_ERR_CUDA, "joining the lane streams failed");
    }
    if (rc) return rc;
    if (reduced && (rc = b4d_fr_end(ctx, pl))) return rc;
    if (want_pc && (rc = track_fused_end(ctx, w, tf, tc, ny, nx, subpixel, eps, track_out))) return rc;
    if (spectral_out) {
        spec_finalize_kernel<<<(unsigned)((tc + 127) / 128), 128, 0, ctx->stream>>>(w.spp, cols_tiles(ny, nx, psd_out != nullptr), spectral_out, tc);
        B4D_LAUNCH_CHECK(ctx);
        if (ny == nx && (rc = run_f95(ctx, psd_out, ny, tc, spectral_out, SCR_F95))) return rc;
    }
    if (grain_out) {
        argmax_reduce_kernel<<<(unsigned)tc, 128, 0, ctx->stream>>>(w.bestA, nblk_ac, w.pk_idx2, w.pk_val2);
        B4D_LAUNCH_CHECK(ctx);
        ProfScope ps(ctx, KC_GRAIN);
        grain_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(acm, ny, w.pk_idx2, ctx->fft->theta, 0.36787944117144233, grain_out);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

// ---- CUDA graphs of the pipelined schedule -------------------------------------------------------------
struct PipeKey {                      // everything the captured launches depend on (compared bytewise: zero it first)
    const void* p[16];
    int64_t tc;
    double d[6];
    float psd_scale;
    int i[8];
    void* scratch[B4D_NSCRATCH];
};
struct PipeGraph {
    PipeKey key;
    int seen = 0;
    bool bad = false;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
    uint64_t last_use = 0;
};
struct PipeGraphCache {
    std::vector<PipeGraph> entries;
    uint64_t clock = 0;
};

bool graphs_enabled(b4d_ctx* ctx) {
    static int env = -2;
    if (env == -2) { const char* e = getenv("B4D_GRAPHS"); env = e ? atoi(e) : -1; }
    const int v = ctx->use_graphs >= 0 ? ctx->use_graphs : (env >= 0 ? env : 1);
    if (!v || ctx->prof_on) return false;
    if (!ctx->pipe && cudaStreamCreateWithFlags(&ctx->pipe, cudaStreamNonBlocking) != cudaSuccess) { ctx->pipe = nullptr; return false; }
    if (!ctx->ev_in && cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming) != cudaSuccess) return false;
    if (!ctx->ev_out && cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming) != cudaSuccess) return false;
    return true;
}

void pipe_graphs_release(b4d_ctx* ctx) {
    PipeGraphCache* c = static_cast<PipeGraphCache*>(ctx->pipe_graphs);
    if (!c) return;
    for (auto& e : c->entries) if (e.exec) cudaGraphExecDestroy(e.exec);
    delete c;
    ctx->pipe_graphs = nullptr;
}

// One batch of the pipelined schedule: directly the first time a set of arguments is seen (scratch arenas, tables and
// function attributes are created then), captured on the second, replayed from then on.
int pipeline_batch_lanes_graph(b4d_ctx* ctx, const Sched& sc, const float* s0, int64_t tc, int ny, int nx, const float* gain,
                               const float* dark, double sat_value, double zero_eps, float psd_scale, int subpixel, double eps,
                               double q_lo, double q_hi, double* fr_out, float* quant_out, int64_t* nvalid_out, float* psd_out,
                               float* ac_out, double* grain_out, double* track_out, int ns, double* spectral_out) {
    auto direct = [&]() {
        return pipeline_batch_lanes(ctx, sc, s0, tc, ny, nx, gain, dark, sat_value, zero_eps, psd_scale, subpixel, eps, q_lo, q_hi,
                                    fr_out, quant_out, nvalid_out, psd_out, ac_out, grain_out, track_out, ns, spectral_out);
    };
    if (!graphs_enabled(ctx)) return direct();
    if (!ctx->pipe_graphs) ctx->pipe_graphs = new PipeGraphCache();
    PipeGraphCache* cache = static_cast<PipeGraphCache*>(ctx->pipe_graphs);
    PipeKey key;
    memset(&key, 0, sizeof(key));
    const void* ptrs[] = {s0, gain, dark, fr_out, quant_out, nvalid_out, psd_out, ac_out, grain_out, track_out,
                          ctx->fft ? ctx->fft->ref : nullptr, ctx->fft ? ctx->fft->ref_nyq : nullptr,
                          ctx->fft ? ctx->fft->blk_list : nullptr, ctx->fft ? ctx->fft->theta : nullptr, spectral_out};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) key.p[i] = ptrs[i];
    key.tc = tc;
    key.d[0] = sat_value == sat_value ? sat_value : -1.2345e300; key.d[1] = zero_eps; key.d[2] = eps; key.d[3] = q_lo; key.d[4] = q_hi;
    key.psd_scale = psd_scale;
    key.i[0] = ny; key.i[1] = nx; key.i[2] = subpixel; key.i[3] = ns; key.i[4] = sc.sub; key.i[5] = sc.slots; key.i[6] = sc.keep; key.i[7] = sc.lanes;
    PipeGraph* ent = nullptr;
    auto refresh_scratch = [&](PipeKey& k) { for (int i = 0; i < B4D_NSCRATCH; ++i) k.scratch[i] = ctx->scratch[i]; };
    auto refresh_tables = [&](PipeKey& k) { k.p[12] = ctx->fft ? ctx->fft->blk_list : nullptr; k.p[13] = ctx->fft ? ctx->fft->theta : nullptr; };
    refresh_scratch(key);
    for (auto& e : cache->entries) if (memcmp(&e.key, &key, sizeof(key)) == 0) { ent = &e; break; }
    if (!ent) {
        // first sight (or the scratch arenas moved since): run directly, remember the arguments with the arenas as they
        // are AFTER the run -- a capture must find everything allocated
        int rc = direct();
        if (rc) return rc;
        refresh_scratch(key);
        refresh_tables(key);
        for (auto& e : cache->entries) if (memcmp(&e.key, &key, sizeof(key)) == 0) return B4D_OK;
        if (cache->entries.size() >= 16) {
            size_t lru = 0;
            for (size_t i = 1; i < cache->entries.size(); ++i) if (cache->entries[i].last_use < cache->entries[lru].last_use) lru = i;
            if (cache->entries[lru].exec) { cudaStreamSynchronize(ctx->pipe); cudaGraphExecDestroy(cache->entries[lru].exec); }
            cache->entries.erase(cache->entries.begin() + lru);
        }
        PipeGraph g;
        g.key = key; g.seen = 1; g.last_use = ++cache->clock;
        cache->entries.push_back(g);
        return B4D_OK;
    }
    ent->last_use = ++cache->clock;
    if (ent->bad) return direct();
    cudaStream_t M = ctx->stream;
    if (!ent->exec) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(ctx->pipe, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); ent->bad = true; return direct(); }
        const int64_t l0 = ctx->launches;
        ctx->stream = ctx->pipe;
        int rc = direct();
        ctx->stream = M;
        const cudaError_t ee = cudaStreamEndCapture(ctx->pipe, &graph);
        bool moved = false;
        for (int i = 0; i < B4D_NSCRATCH; ++i) moved = moved || key.scratch[i] != ctx->scratch[i];
        if (rc || ee != cudaSuccess || !graph || moved || cudaGraphInstantiate(&ent->exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            ent->exec = nullptr;
            ent->bad = true;
            ctx->launches = l0;
            return direct();
        }
        cudaGraphDestroy(graph);
        ent->launches = ctx->launches - l0;
        ctx->launches = l0;
    }
    B4D_CUDA(ctx, cudaEventRecord(ctx->ev_in, M));
    B4D_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe, ctx->ev_in, 0));
    B4D_CUDA(ctx, cudaGraphLaunch(ent->exec, ctx->pipe));
    B4D_CUDA(ctx, cudaEventRecord(ctx->ev_out, ctx->pipe));
    B4D_CUDA(ctx, cudaStreamWaitEvent(M, ctx->ev_out, 0));
    ctx->launches += ent->launches;
    return B4D_OK;
}

int stack_pipeline_body(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                        const float* dark, double sat_value, double zero_eps, float psd_scale, int subpixel,
                        double eps, double q_lo, double q_hi, double* fr_out, float* quant_out,
                        int64_t* nvalid_out, float* psd_out, float* ac_out, double* grain_out, double* track_out,
                        double* spectral_out);

extern "C" int b4d_stack_pipeline(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                                  const float* dark, double sat_value, double zero_eps, float psd_scale, int subpixel,
                                  double eps, double q_lo, double q_hi, double* fr_out, float* quant_out,
                                  int64_t* nvalid_out, float* psd_out, float* ac_out, double* grain_out, double* track_out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    return stack_pipeline_body(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, psd_scale, subpixel, eps, q_lo, q_hi,
                               fr_out, quant_out, nvalid_out, psd_out, ac_out, grain_out, track_out, nullptr);
}

extern "C" int b4d_stack_pipeline_ref(b4d_ctx* ctx, const b4d_ref* ref, const float* stack, int64_t n_frames, int ny, int nx,
                                      const float* gain, const float* dark, double sat_value, double zero_eps, float psd_scale,
                                      int subpixel, double eps, double q_lo, double q_hi, double* fr_out, float* quant_out,
                                      int64_t* nvalid_out, float* psd_out, float* ac_out, double* grain_out, double* track_out,
                                      double* spectral_out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (track_out && !ref) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline_ref: tracking needs a reference");
    RefSwap sw(ctx, ref);
    return stack_pipeline_body(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, psd_scale, subpixel, eps, q_lo, q_hi,
                               fr_out, quant_out, nvalid_out, psd_out, ac_out, grain_out, track_out, spectral_out);
}

int stack_pipeline_body(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                        const float* dark, double sat_value, double zero_eps, float psd_scale, int subpixel,
                        double eps, double q_lo, double q_hi, double* fr_out, float* quant_out,
                        int64_t* nvalid_out, float* psd_out, float* ac_out, double* grain_out, double* track_out,
                        double* spectral_out) {
    if (spectral_out && !(ac_out || grain_out))
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline: the spectral sums ride on the autocorrelation branch (ask for ac_out or grain_out)");
    if (spectral_out && !pow2_sides(ny, nx))
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_stack_pipeline: spectral sums in the fused pass need power-of-two sides (use b4d_psd2d)");
    if (dark && !gain) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline: dark given without gain");
    if (grain_out && ny != nx) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline: grain widths need square frames");
    if (quant_out && (!nvalid_out || !(q_lo >= 0.0 && q_lo < q_hi && q_hi <= 1.0)))
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline: tail percentiles need 0 <= q_lo < q_hi <= 1 and nvalid_out");
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_stack_pipeline", stack, n_frames, ny, nx);
        if (rcg) return rcg;
        if (track_out && (!ctx->fft || !ctx->fft->gref || ctx->fft->ref_ny != ny || ctx->fft->ref_nx != nx))
            return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline: tracking needs b4d_phase_set_reference for (%d, %d) frames", ny, nx);
        return gen_stack_pipeline(ctx, stack, n_frames, ny, nx, gain, dark, sat_value, zero_eps, psd_scale, subpixel, eps, q_lo, q_hi,
                                  fr_out, quant_out, nvalid_out, psd_out, ac_out, grain_out, track_out);
    }
    int rc = check_fft_args(ctx, "b4d_stack_pipeline", stack, n_frames, ny, nx);
    if (rc) return rc;
    const bool want_ac = ac_out || grain_out, want_pc = track_out != nullptr;
    if (want_pc && (!ctx->fft || !ctx->fft->ref || ctx->fft->ref_ny != ny || ctx->fft->ref_nx != nx))
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_stack_pipeline: tracking needs b4d_phase_set_reference for (%d, %d) frames", ny, nx);
    const int64_t B = batch_frames(ctx, ny, nx, 5);
    const size_t npix = (size_t)ny * nx;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        const float* s0 = stack + (size_t)t0 * npix;
        // tracker: fused median (sample rows + 3x3 window instead of the |corr| map) unless the frame is too small
        const int ns = (want_pc && ctx->fused_median) ? fused_sample_blocks(ny, nx) : 0;
        const Sched sc = pipeline_sched(ctx);
        if (sc.sub > 0 && sc.sub < tc && (!want_pc || ns > 0) && lanes_ready(ctx, sc.lanes, sc.slots)) {
            rc = pipeline_batch_lanes_graph(ctx, sc, s0, tc, ny, nx, gain, dark, sat_value, zero_eps, psd_scale, subpixel, eps, q_lo, q_hi,
                                      fr_out ? fr_out + t0 * B4D_FR_NCOLS : nullptr, quant_out ? quant_out + 4 * t0 : nullptr,
                                      quant_out ? nvalid_out + t0 : nullptr, psd_out ? psd_out + (size_t)t0 * npix : nullptr,
                                      ac_out ? ac_out + (size_t)t0 * npix : nullptr, grain_out ? grain_out + t0 * 4 : nullptr,
                                      track_out ? track_out + t0 * 4 : nullptr, ns,
                                      spectral_out ? spectral_out + t0 * B4D_SP_NCOLS : nullptr);
            if (rc) return rc;
            continue;
        }
        Work w;
        if ((rc = carve(ctx, tc, ny, nx, want_ac, want_pc, &w))) return rc;
        // scratch maps: |corr| always, autocorr when the caller does not keep it, PSD when only f95 reads it
        void* p = nullptr;
        const size_t mag_floats = !want_pc ? 0 : (ns ? ((fused_scratch_floats(ny, nx, ns, tc) + 63) & ~size_t(63)) : npix * tc);
        const size_t ac_floats = (want_ac && !ac_out) ? npix * tc : 0;
        const bool psd_scratch = spectral_out && ny == nx && !psd_out;
        const size_t need = sizeof(float) * (mag_floats + ac_floats + (psd_scratch ? npix * tc : 0)) +
                            sizeof(double) * B4D_FR_NCOLS * tc + 256;
        if ((rc = b4d_scratch(ctx, SCR_MAP, need, &p))) return rc;
        double* fr = static_cast<double*>(p);
        float* mag = reinterpret_cast<float*>(fr + (size_t)B4D_FR_NCOLS * tc);
        float* acm = ac_out ? ac_out + (size_t)t0 * npix : mag + mag_floats;
        float* psd_b = psd_out ? psd_out + (size_t)t0 * npix : (psd_scratch ? mag + mag_floats + ac_floats : nullptr);
        double* frp = fr_out ? fr_out + t0 * B4D_FR_NCOLS : fr;
        const bool reduced = fr_out || want_pc || quant_out;
        // The reduction pass and the forward row pass both stream the frames. They alternate `pair` frames at a time, so
        // that the row pass finds in L2 what the reduction pass has just read from HBM (one HBM read of every frame
        // instead of two); the per-frame tail kernels of the reductions run once per batch after the last pair.
        static int pair_env = -2;
        if (pair_env == -2) { const char* e = getenv("B4D_PAIR"); pair_env = e ? atoi(e) : B4D_PAIR_DEFAULT; }
        const int64_t pair = (ctx->sched_pair >= 0 ? ctx->sched_pair : pair_env);
        const bool need_fft = psd_out || want_ac || want_pc;
        if (reduced) {
            FrTails tl = {q_lo, q_hi, quant_out ? quant_out + 4 * t0 : nullptr, quant_out ? nvalid_out + t0 : nullptr};
            if (pair > 0 && pair < tc && need_fft) {
                FrPlan pl;
                if ((rc = b4d_fr_begin(ctx, s0, tc, ny, nx, gain, dark, sat_value, zero_eps, frp, quant_out ? &tl : nullptr, w.pilot, &pl)))
                    return rc;
                for (int64_t a = 0; a < tc; a += pair) {
                    const int64_t f = tc - a < pair ? tc - a : pair;
                    if ((rc = b4d_fr_range(ctx, pl, a, f))) return rc;
                    Work ws = w;
                    ws.pilot = w.pilot + a;
                    ws.H = w.H + (size_t)a * ny * (nx / 2);
                    if ((rc = run_rows_fwd(ctx, s0 + (size_t)a * npix, f, ny, nx, gain, dark, ws, true, true))) return rc;
                }
                if ((rc = b4d_fr_end(ctx, pl))) return rc;
            } else {
                if ((rc = b4d_frame_reductions_ex(ctx, s0, tc, ny, nx, gain, dark, sat_value, zero_eps, frp, quant_out ? &tl : nullptr,
                                                  w.pilot)))
                    return rc;
                if (need_fft && (rc = run_rows_fwd(ctx, s0, tc, ny, nx, gain, dark, w, true, true))) return rc;
            }
        } else if (need_fft && (rc = run_rows_fwd(ctx, s0, tc, ny, nx, gain, dark, w, true, false))) return rc;
        if (!need_fft) continue;
        ColsArgs c = cols_defaults(w, nx, true);
        c.psd_out = psd_b;
        c.psd_scale = psd_scale;
        if (want_ac) { c.i2_ac = w.I2a; c.i2_ac_nyq = w.I2nyq; c.ac_partials = w.acp; }
        if (spectral_out) c.spec_partials = w.spp;   // bandwidth / spectral-entropy sums from the same forward transform
        if (want_pc) {
            c.i2_pc = w.I2b;
            c.R = ctx->fft->ref; c.Rnyq = ctx->fft->ref_nyq; c.r_stride = 0; c.rnyq_stride = 0;
            c.whiten = 1; c.eps = (float)eps; c.fr = frp; c.fr_stride = B4D_FR_NCOLS;
        }
        // psd2d keeps the DC bin, the autocorrelation (remove_mean) drops it: ac_zero_dc handles that in-kernel.
        c.zero_dc = 0;
        c.ac_zero_dc = 1;
        if ((rc = run_cols(ctx, c, tc, ny))) return rc;
        RowsInvArgs r;
        memset(&r, 0, sizeof(r));
        r.ny = ny;
        int nblk = 0;
        // After the column pass the two branches are independent. With both on, the tracker's branch (issue-bound row pass
        // between small latency-bound kernels: sample rows, median bracket, final median, 3x3 window) goes to the
        // high-priority side stream and the autocorrelation branch (DRAM-bound row pass, argmax, grain widths) stays on the
        // caller's stream: the small kernels are placed as soon as an SM has room, the big row pass fills every gap they
        // leave, and CTAs of the two row passes share the SMs. Joined before the batch ends (the scratch is reused).
        const bool forked = want_ac && want_pc && side_stream_ready(ctx);
        cudaStream_t main_stream = ctx->stream;
        auto ac_branch = [&]() -> int {
            // the column pass left the autocorrelation branch packed (two real columns per transform): its own row pass
            // over the rows 0 .. ny/2; the tracker's rows go on their own (two rows per transform)
            int nblk_ac = 0;
            RowsInvAcArgs ra;
            memset(&ra, 0, sizeof(ra));
            ra.Iz = w.I2a; ra.Inyq = w.I2nyq; ra.ny = ny; ra.ch_log2 = log2i(cols_cw(ny, cols_wide_out(c)) / 2);
            ra.out = acm; ra.norm = w.acp; ra.n_norm = cols_tiles(ny, nx, cols_wide_out(c)); ra.norm_mult = 1.0; ra.scale = 1.0 / ((double)nx * ny);
            ra.best = grain_out ? w.bestA : nullptr;
            int r2 = run_rows_inv_ac(ctx, ra, tc, nx, &nblk_ac);
            if (r2) return r2;
            if (grain_out) {
                argmax_reduce_kernel<<<(unsigned)tc, 128, 0, ctx->stream>>>(w.bestA, nblk_ac, w.pk_idx2, w.pk_val2);
                B4D_LAUNCH_CHECK(ctx);
                ProfScope ps(ctx, KC_GRAIN);
                grain_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(acm, ny, w.pk_idx2, ctx->fft->theta, 0.36787944117144233,
                                                                     grain_out + t0 * 4);
                B4D_LAUNCH_CHECK(ctx);
            }
            return B4D_OK;
        };
        if (grain_out && (rc = ensure_theta(ctx))) return rc;
        if (forked) {
            B4D_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
            B4D_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            ctx->stream = ctx->side;
        }
        if (want_pc) {
            r.Ia = w.I2b; r.outA = mag; r.kindA = 1; r.scaleA = 1.0 / ((double)nx * ny); r.bestA = w.bestB;
            nblk = rows_inv_blocks(nx, ny);
            if (ns) {
                rc = rows_inv_track_fused(ctx, w, r, tc, ny, nx, ns, mag, subpixel, eps, track_out + t0 * 4);
            } else {
                rc = run_rows_inv(ctx, r, tc, nx);
                if (!rc) rc = track_finish(ctx, w, mag, tc, ny, nx, nblk, subpixel, eps, track_out + t0 * 4);
            }
        }
        if (forked) {
            ctx->stream = main_stream;
            if (!rc && cudaEventRecord(ctx->ev_join, ctx->side) != cudaSuccess) rc = b4d_fail(ctx, B4D_ERR_CUDA, "cudaEventRecord failed");
        }
        if (!rc && want_ac) rc = ac_branch();
        if (forked) {                                   // joined on every path: the side stream's work is in flight
            cudaError_t je = cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
            if (!rc && je != cudaSuccess) rc = b4d_fail(ctx, B4D_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(je));
        }
        if (rc) return rc;
        if (spectral_out) {
            double* tab = spectral_out + t0 * B4D_SP_NCOLS;
            spec_finalize_kernel<<<(unsigned)((tc + 127) / 128), 128, 0, ctx->stream>>>(w.spp, cols_tiles(ny, nx, cols_wide_out(c)), tab, tc);
            B4D_LAUNCH_CHECK(ctx);
            if (ny == nx && (rc = run_f95(ctx, psd_b, ny, tc, tab, SCR_F95))) return rc;
        }
    }
    return B4D_OK;
}

extern "C" int b4d_template_match(b4d_ctx* ctx, const float* tpl, int per_frame, int h, int w, const float* stack, int64_t n_frames,
                                  int ny, int nx, double ref_y, double ref_x, int subpixel, double eps, double* out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!tpl || !out || h < 1 || w < 1 || h > ny || w > nx)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_template_match: template shape (%d, %d) must fit inside image shape (%d, %d)", h, w, ny, nx);
    if (!pow2_sides(ny, nx)) {
        int rcg = check_gen_args(ctx, "b4d_template_match", stack, n_frames, ny, nx);
        return rcg ? rcg : gen_template_match(ctx, tpl, per_frame, h, w, stack, n_frames, ny, nx, ref_y, ref_x, subpixel, eps, out);
    }
    int rc = check_fft_args(ctx, "b4d_template_match", stack, n_frames, ny, nx);
    if (rc) return rc;
    const size_t npix = (size_t)ny * nx, half = (size_t)ny * (nx / 2);
    const int oy = ny - h + 1, ox = nx - w + 1;
    const size_t nres = (size_t)oy * ox;
    // batch: correlation map + result map (float), two integral images (double) and, with one template per frame, its
    // conjugate spectrum, per frame; at most ~1.5 GB
    int64_t B = ((int64_t)1536 << 20) / (int64_t)(npix * (per_frame ? 28 : 24));
    if (B < 1) B = 1;
    const int64_t Bfft = batch_frames(ctx, ny, nx, 2);
    if (B > Bfft) B = Bfft;
    if (B > n_frames) B = n_frames;
    const int64_t nt = per_frame ? B : 1;               // template spectra alive at a time
    void* p = nullptr;
    if ((rc = b4d_scratch(ctx, SCR_NYQ, sizeof(float2) * (half + ny) * nt + sizeof(double) * B4D_FR_NCOLS * nt + 256, &p))) return rc;
    float2* tref = static_cast<float2*>(p);
    float2* tref_nyq = tref + half * nt;
    double* tfr = reinterpret_cast<double*>(tref_nyq + (size_t)ny * nt);
    if ((rc = b4d_scratch(ctx, SCR_MAP, sizeof(float) * (npix + nres) * B + sizeof(double) * 2 * npix * B + 256, &p))) return rc;
    float* corr = static_cast<float*>(p);
    float* res = corr + npix * B;
    double* P1 = reinterpret_cast<double*>(res + ((nres * B + 1) & ~size_t(1)));
    double* P2 = P1 + npix * B;
    // templates: z-scored over the ROI (tracking.py:151, :308-311), embedded at the origin, conjugate spectra
    auto template_spectra = [&](const float* tp, int64_t count) -> int {
        int r = b4d_frame_reductions_nolock(ctx, tp, count, h, w, nullptr, nullptr, nan(""), 0.0, tfr);
        if (r) return r;
        embed_template_kernel<<<dim3(ctx->sm_count * 2, (unsigned)count), 256, 0, ctx->stream>>>(tp, h, w, ny, nx, 0, 0, (float)eps, tfr, corr);
        B4D_LAUNCH_CHECK(ctx);
        Work wk;
        if ((r = carve(ctx, count, ny, nx, false, false, &wk))) return r;
        if ((r = run_rows_fwd(ctx, corr, count, ny, nx, nullptr, nullptr, wk, false))) return r;
        ColsArgs c = cols_defaults(wk, nx, false);
        c.conj_out = tref;
        c.conj_nyq_out = tref_nyq;
        return run_cols(ctx, c, count, ny);
    };
    if (!per_frame && (rc = template_spectra(tpl, 1))) return rc;
    for (int64_t t0 = 0; t0 < n_frames; t0 += B) {
        const int64_t tc = n_frames - t0 < B ? n_frames - t0 : B;
        const float* s0 = stack + t0 * npix;
        if (per_frame && (rc = template_spectra(tpl + (size_t)t0 * h * w, tc))) return rc;
        Work w_;
        if ((rc = carve(ctx, tc, ny, nx, true, false, &w_))) return rc;
        if ((rc = run_rows_fwd(ctx, s0, tc, ny, nx, nullptr, nullptr, w_, true))) return rc;      // frames - pilot
        ColsArgs c = cols_defaults(w_, nx, false);                                                  // (no DC add-back)
        c.i2_pc = w_.I2a;
        c.R = tref; c.Rnyq = tref_nyq; c.r_stride = per_frame ? half : 0; c.rnyq_stride = per_frame ? (size_t)ny : 0;
        c.whiten = 0; c.eps = 0.f;
        if ((rc = run_cols(ctx, c, tc, ny))) return rc;
        RowsInvArgs r;
        memset(&r, 0, sizeof(r));
        r.Ia = w_.I2a; r.ny = ny; r.outA = corr; r.kindA = 0; r.scaleA = 1.0 / ((double)nx * (double)ny);
        if ((rc = run_rows_inv(ctx, r, tc, nx))) return rc;
        {
            ProfScope ps(ctx, KC_SMALL);
            const int64_t rows = tc * ny;
            tm_scan_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(s0, w_.pilot, ny, nx, P1, P2, rows);
            B4D_LAUNCH_CHECK(ctx);
            tm_scan_cols_kernel<<<dim3((nx + 255) / 256, (unsigned)tc), 256, 0, ctx->stream>>>(P1, P2, ny, nx);
            B4D_LAUNCH_CHECK(ctx);
            int bx = (int)((nres + 2047) / 2048);
            if (bx > 592) bx = 592;
            tm_normalise_kernel<<<dim3(bx, (unsigned)tc), 256, 0, ctx->stream>>>(corr, P1, P2, ny, nx, h, w, tfr,
                                                                               per_frame ? B4D_FR_NCOLS : 0, eps, res);
            B4D_LAUNCH_CHECK(ctx);
            gen_argmax_kernel<<<(unsigned)tc, 1024, 0, ctx->stream>>>(res, (int64_t)nres, w_.pk_idx);
            B4D_LAUNCH_CHECK(ctx);
        }
        void* q = nullptr;
        if ((rc = b4d_scratch(ctx, SCR_MISC, 1024, &q))) return rc;    // sized by carve(); first 512 B hold the quantile
        static const double half_q = 0.5;
        if ((rc = b4d_put_doubles(ctx, static_cast<double*>(q), &half_q, 1))) return rc;
        if ((rc = b4d_select_impl(ctx, res, tc, (int64_t)nres, static_cast<const double*>(q), 1, 1, w_.med,
                                  reinterpret_cast<int64_t*>(w_.nvalid))))
            return rc;
        tm_finalize_kernel<<<(unsigned)((tc + 63) / 64), 64, 0, ctx->stream>>>(res, w_.pk_idx, oy, ox, h, w, ref_y, ref_x, w_.med,
                                                                               w_.nvalid, subpixel, eps, out + t0 * 4, tc);
        B4D_LAUNCH_CHECK(ctx);
    }
    return B4D_OK;
}

extern "C" int b4d_ifft2d(b4d_ctx* ctx, const float* spec_c64, int64_t n_frames, int ny, int nx, float* out_c64) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!out_c64) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_ifft2d: null output");
    int rc = check_gen_args(ctx, "b4d_ifft2d", spec_c64, n_frames, ny, nx);
    if (rc) return rc;
    return gen_ifft2d(ctx, reinterpret_cast<const float2*>(spec_c64), n_frames, ny, nx, reinterpret_cast<float2*>(out_c64));
}
