// Shared infrastructure of libb4d.so: context, error plumbing, warp/block reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/b4d.h"

struct FftPlanCache;   // spectral.cu

// kernel classes for the optional CUDA-event profile (b4d_profile_*)
enum {
    KC_PILOT = 0, KC_FRAME_REDUCE, KC_ROWS_FWD, KC_COLS, KC_ROWS_INV, KC_SELECT_HIST, KC_SELECT_SCAN, KC_GRAIN,
    KC_TEMPORAL, KC_FLATFIELD, KC_SMALL, KC_SELECT_SAMPLE, KC_SELECT_COLLECT, KC_SELECT_FINAL, KC_ROWS_INV_AC, KC_GENERIC, KC_COUNT
};

struct ProfSpan {
    int klass;
    cudaEvent_t a, b;
};

// cudaFuncSetAttribute is per device: the "already raised the shared-memory limit" flags of the launchers are too
constexpr int B4D_MAX_DEVICES = 64;

constexpr int B4D_NSCRATCH = 12;

struct b4d_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    // grow-only scratch arenas (device)
    void* scratch[B4D_NSCRATCH] = {};
    size_t scratch_bytes[B4D_NSCRATCH] = {};
    FftPlanCache* fft = nullptr;
    int64_t batch_override = 0;       // frames per internal batch of the FFT pipeline (0 = automatic)
    size_t batch_key[4] = {0, 0, 0, 0};   // automatic batch sizes already worked out: per-frame scratch bytes -> frames
    int64_t batch_val[4] = {0, 0, 0, 0};
    bool fused_median = true;         // tracker SNR: median of |corr| taken inside the inverse row pass (no map)
    bool prof_on = false;
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_pool;
    int cur_class = KC_SMALL;
    // side stream of the fused stack pipeline: the autocorrelation branch (row pass, argmax, grain widths) runs on it
    // next to the tracker's branch on `stream`; created on first use, non-blocking
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // frame-pipelined schedule of the fused stack pipeline (spectral.cu, "lanes"): the big kernels run `sched_sub` frames
    // at a time so that the row <-> column intermediates are consumed from L2; the autocorrelation row pass and the
    // tracker's row passes follow the column pass on two side streams through `sched_slots` ring slots of intermediates.
    // -1 = take the default / the environment (B4D_SUB, B4D_SLOTS, B4D_KEEP).
    int sched_sub = -1, sched_lanes = -1, sched_slots = -1, sched_keep = -1, sched_pair = -1;
    int keep_mode = 0;                // cache policy the FFT launchers pass to their kernels (see st_inter in spectral.cu)
    std::vector<cudaStream_t> lane_streams;
    std::vector<cudaEvent_t> lane_ev;
    // CUDA graphs of the frame-pipelined schedule (spectral.cu): a batch is ~10 launches and a few event operations per
    // step of 1 - 2 frames, more than the host can issue in the time the GPU needs for them, so the second call with the
    // same arguments captures the whole batch and every later one replays it. use_graphs: -1 = default / B4D_GRAPHS.
    int use_graphs = -1;
    cudaStream_t pipe = nullptr;      // capture / replay stream (the caller's stream may be the legacy default stream)
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    void* pipe_graphs = nullptr;
    std::mutex lock;
};

// Every extern "C" entry point holds one of these: the context's lock, and the context's device made current for the
// duration of the call (a process may drive several devices, each through its own context).
struct B4dCall {
    std::lock_guard<std::mutex> guard;
    int restore = -1;
    explicit B4dCall(b4d_ctx* c) : guard(c->lock) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != c->device && cudaSetDevice(c->device) == cudaSuccess) restore = cur;
    }
    ~B4dCall() { if (restore >= 0) cudaSetDevice(restore); }
    B4dCall(const B4dCall&) = delete;
    B4dCall& operator=(const B4dCall&) = delete;
};

// Brackets the next launches with CUDA events when profiling is on (events live on ctx->stream).
struct ProfScope {
    b4d_ctx* c;
    ProfSpan s;
    bool on;
    ProfScope(b4d_ctx* ctx, int klass) : c(ctx), on(ctx->prof_on) {
        if (!on) return;
        s.klass = klass;
        for (cudaEvent_t* e : {&s.a, &s.b}) {
            if (!c->prof_pool.empty()) { *e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else cudaEventCreate(e);
        }
        cudaEventRecord(s.a, c->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(s.b, c->stream);
        c->prof_spans.push_back(s);
    }
};

// scratch slot ids
enum { SCR_REDUCE = 0, SCR_PILOT = 1, SCR_SELECT = 2, SCR_SPEC_A = 3, SCR_SPEC_B = 4, SCR_SPEC_C = 5, SCR_MISC = 6, SCR_MAP = 7, SCR_NYQ = 8, SCR_GEN = 9, SCR_F95 = 10 };

inline int b4d_fail(b4d_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        ctx->last_error = buf;
    }
    return code;
}

#define B4D_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return b4d_fail((ctx), B4D_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__,     \
                            __LINE__, cudaGetErrorString(_e));                                   \
    } while (0)

#define B4D_LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                         \
        (ctx)->launches++;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess)                                                                   \
            return b4d_fail((ctx), B4D_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, \
                            __LINE__, cudaGetErrorString(_e));                                   \
    } while (0)

// Returns a device scratch buffer of at least `bytes` bytes (grow-only, stream-ordered reuse).
inline int b4d_scratch(b4d_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            B4D_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t want = (bytes + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess)
            return b4d_fail(ctx, B4D_ERR_NOMEM, "cudaMalloc(%zu) for scratch slot %d: %s", want, slot,
                            cudaGetErrorString(e));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return B4D_OK;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

#endif  // __CUDACC__
