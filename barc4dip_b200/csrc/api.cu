// Context management and plain memory helpers of the C ABI (include/b4d.h).
#include "common.cuh"

void b4d_fft_release(b4d_ctx* ctx);   // fft.cu

extern "C" const char* b4d_version(void) { return "barc4dip_b200 0.1 (sm_100a)"; }

extern "C" int b4d_create(int device, b4d_ctx** out) {
    if (!out) return B4D_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev || device >= B4D_MAX_DEVICES) return B4D_ERR_CUDA;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return B4D_ERR_CUDA;
    b4d_ctx* c = new (std::nothrow) b4d_ctx();
    if (!c) { if (prev >= 0) cudaSetDevice(prev); return B4D_ERR_NOMEM; }
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    *out = c;
    if (prev >= 0 && prev != device) cudaSetDevice(prev);     // the caller's current device is left as it was
    return B4D_OK;
}

extern "C" int b4d_destroy(b4d_ctx* ctx) {
    if (!ctx) return B4D_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    b4d_fft_release(ctx);
    for (int i = 0; i < B4D_NSCRATCH; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    for (auto& sp : ctx->prof_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    for (auto& st : ctx->lane_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (ctx->pipe) { cudaStreamSynchronize(ctx->pipe); cudaStreamDestroy(ctx->pipe); }
    for (auto& e : ctx->lane_ev) cudaEventDestroy(e);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (ctx->ev_out) cudaEventDestroy(ctx->ev_out);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
    return B4D_OK;
}

extern "C" int b4d_set_stream(b4d_ctx* ctx, void* cuda_stream) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return B4D_OK;
}

extern "C" int b4d_synchronize(b4d_ctx* ctx) {
    if (!ctx) return B4D_ERR_INVALID;
    B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B4D_OK;
}

extern "C" const char* b4d_last_error(b4d_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int64_t b4d_launch_count(b4d_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int b4d_device_sm_count(b4d_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int b4d_profile_begin(b4d_ctx* ctx) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    for (auto& sp : ctx->prof_spans) { ctx->prof_pool.push_back(sp.a); ctx->prof_pool.push_back(sp.b); }
    ctx->prof_spans.clear();
    ctx->prof_on = true;
    return B4D_OK;
}

extern "C" int b4d_profile_end(b4d_ctx* ctx, double* ms_per_class, int64_t* launches_per_class) {
    if (!ctx || !ms_per_class || !launches_per_class) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    ctx->prof_on = false;
    B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < KC_COUNT; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
    for (auto& sp : ctx->prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { ms_per_class[sp.klass] += ms; launches_per_class[sp.klass]++; }
        ctx->prof_pool.push_back(sp.a);
        ctx->prof_pool.push_back(sp.b);
    }
    ctx->prof_spans.clear();
    return B4D_OK;
}

extern "C" const char* b4d_profile_class_name(int k) {
    static const char* names[KC_COUNT] = {"pilot", "frame_reduce", "rows_fwd", "cols", "rows_inv", "select_hist",
                                          "select_scan", "grain", "temporal", "flatfield", "small", "select_sample", "select_collect",
                                          "select_final", "rows_inv_ac", "generic_dft"};
    return (k >= 0 && k < KC_COUNT) ? names[k] : "?";
}

extern "C" int b4d_set_batch_frames(b4d_ctx* ctx, int64_t frames) {
    if (!ctx || frames < 0) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    ctx->batch_override = frames;
    return B4D_OK;
}

extern "C" int b4d_set_schedule(b4d_ctx* ctx, int sub_frames, int lanes, int ring_slots, int keep, int use_graphs) {
    if (!ctx || ring_slots > 8 || lanes > 8) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    ctx->sched_sub = sub_frames;
    ctx->sched_lanes = lanes;
    ctx->sched_slots = ring_slots;
    ctx->sched_keep = keep;
    ctx->use_graphs = use_graphs;
    return B4D_OK;
}

extern "C" int b4d_set_pairing(b4d_ctx* ctx, int pair_frames) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    ctx->sched_pair = pair_frames;
    return B4D_OK;
}

extern "C" int b4d_set_fused_median(b4d_ctx* ctx, int on) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    ctx->fused_median = on != 0;
    return B4D_OK;
}

extern "C" int b4d_malloc(b4d_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return B4D_ERR_INVALID;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) return b4d_fail(ctx, B4D_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return B4D_OK;
}

extern "C" int b4d_free(b4d_ctx* ctx, void* p) {
    if (!ctx) return B4D_ERR_INVALID;
    B4D_CUDA(ctx, cudaFree(p));
    return B4D_OK;
}

extern "C" int b4d_memcpy_h2d(b4d_ctx* ctx, void* dst, const void* src_host, size_t bytes) {
    if (!ctx || !dst || !src_host) return B4D_ERR_INVALID;
    B4D_CUDA(ctx, cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B4D_OK;
}

extern "C" int b4d_memcpy_d2h(b4d_ctx* ctx, void* dst_host, const void* src, size_t bytes) {
    if (!ctx || !dst_host || !src) return B4D_ERR_INVALID;
    B4D_CUDA(ctx, cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B4D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B4D_OK;
}

// ---- stack ingestion: detector-native integer frames are uploaded as they are and widened on the device ------------
// (the reference casts integer images to float32 / float64 on the host: signal/tracking.py:299-305, metrics/*.py)
namespace {
template <typename T>
__global__ void __launch_bounds__(256) cast_to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
    // 8 elements per thread: one 16-byte (or narrower) load of the source type, two float4 stores
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 + 8 <= n) {
        T v[8];
        if (sizeof(T) == 2) *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(src + i0));
        else if (sizeof(T) == 1) *reinterpret_cast<uint2*>(v) = __ldcs(reinterpret_cast<const uint2*>(src + i0));
        else { *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(src + i0)); *reinterpret_cast<uint4*>(v + 4) = __ldcs(reinterpret_cast<const uint4*>(src + i0 + 4)); }
        float4 a = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
        float4 b = make_float4((float)v[4], (float)v[5], (float)v[6], (float)v[7]);
        reinterpret_cast<float4*>(dst + i0)[0] = a;
        reinterpret_cast<float4*>(dst + i0)[1] = b;
    } else {
        for (int64_t i = i0; i < n; ++i) dst[i] = (float)src[i];
    }
}
template <typename T>
int launch_cast(b4d_ctx* ctx, const void* src, float* dst, int64_t n) {
    const int64_t threads = (n + 7) / 8;
    ProfScope ps(ctx, KC_SMALL);
    cast_to_f32_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(static_cast<const T*>(src), dst, n);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}
}  // namespace

extern "C" int b4d_cast_to_f32(b4d_ctx* ctx, const void* src, int dtype, float* dst, int64_t n) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!src || !dst || n < 1) return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_cast_to_f32: bad arguments");
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_cast_to_f32: pointers must be 16-byte aligned");
    if (n > ((int64_t)1 << 40)) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_cast_to_f32: too many elements");
    switch (dtype) {
        case B4D_U8: return launch_cast<uint8_t>(ctx, src, dst, n);
        case B4D_U16: return launch_cast<uint16_t>(ctx, src, dst, n);
        case B4D_I16: return launch_cast<int16_t>(ctx, src, dst, n);
        case B4D_I32: return launch_cast<int32_t>(ctx, src, dst, n);
        case B4D_U32: return launch_cast<uint32_t>(ctx, src, dst, n);
        default: return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_cast_to_f32: dtype code %d", dtype);
    }
}
