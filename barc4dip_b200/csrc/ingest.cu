// Stack ingestion on the device: compressed HDF5 chunks -> float32 frames in HBM without the host inflating anything.
//
// The reference inflates a stack on the host inside h5py (io/h5.py:80 `dset[()]`, gzip-4 chunks written at :204-210) and
// casts frame by frame (signal/tracking.py:299-305). Here the deflate streams cross PCIe as stored, the GPU's hardware
// decompression engine inflates them (cuMemBatchDecompressAsync, CUDA 12.8+, one batch per block of frames), and one
// kernel undoes the rest of the HDF5 filter pipeline and chunk layout on the way to float32:
//     byte-shuffle filter (planes of k-th bytes)  +  chunk tiles (c0, cy, cx) -> row-major frames  +  integer -> float32.
//
// libcuda is reached through cudaGetDriverEntryPoint only: libb4d.so keeps loading where no driver is installed.

#include <cuda.h>

#include "common.cuh"

namespace {

typedef CUresult (*PfnBatchDecompress)(CUmemDecompressParams*, size_t, unsigned int, size_t*, CUstream);
typedef CUresult (*PfnDeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
typedef CUresult (*PfnDeviceGet)(CUdevice*, int);
typedef CUresult (*PfnGetErrorString)(CUresult, const char**);

template <typename F>
bool driver_entry(const char* name, F* out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(p);
    return true;
}

struct InflateCaps {
    bool known = false;
    int algo_mask = 0;
    int64_t max_bytes = 0;
    PfnBatchDecompress submit = nullptr;
};
InflateCaps g_caps[B4D_MAX_DEVICES];
std::mutex g_caps_lock;

// parameter arrays of the batches in flight: the driver's contract does not say when it is done reading them, so the
// last few are kept alive (a block of frames is one batch; ingestion keeps two blocks in flight)
constexpr int PARAM_RING = 8;
std::vector<CUmemDecompressParams> g_params[PARAM_RING];
int g_param_next = 0;

const InflateCaps& inflate_caps(int device) {
    std::lock_guard<std::mutex> g(g_caps_lock);
    InflateCaps& c = g_caps[device % B4D_MAX_DEVICES];
    if (c.known) return c;
    c.known = true;
    PfnDeviceGetAttribute get_attr = nullptr;
    PfnDeviceGet dev_get = nullptr;
    if (!driver_entry("cuDeviceGetAttribute", &get_attr) || !driver_entry("cuDeviceGet", &dev_get)) return c;
    if (!driver_entry("cuMemBatchDecompressAsync", &c.submit)) return c;          // driver older than 12.8
    CUdevice d;
    int mask = 0, maxlen = 0;
    if (dev_get(&d, device) != CUDA_SUCCESS) return c;
    if (get_attr(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, d) != CUDA_SUCCESS) return c;
    if (get_attr(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, d) != CUDA_SUCCESS) return c;
    c.algo_mask = mask;
    c.max_bytes = maxlen;
    return c;
}

template <typename T> struct Quad;   // four consecutive elements as one aligned load
template <> struct Quad<uint8_t> { typedef uint32_t type; };
template <> struct Quad<uint16_t> { typedef uint2 type; };
template <> struct Quad<int16_t> { typedef uint2 type; };
template <> struct Quad<int32_t> { typedef uint4 type; };
template <> struct Quad<uint32_t> { typedef uint4 type; };
template <> struct Quad<float> { typedef uint4 type; };

// One thread = V consecutive pixels (V = 4 when nx and cx are multiples of 4, else 1) of UNCHUNK_ROWS consecutive rows: the
// grid is (pieces of a row, groups of rows, frames), the three divisions (frame / c0, row / cy, x / cx) are paid once
// per thread and the walk down the rows is additions -- the element index moves by cx, or to the next tile row.
// (History, ncu on 16 frames of 2048^2 uint16: flattened 64-bit index 229 instructions per warp and row, 140 us; 3-D grid
// with one row per thread 152, 98 us; this form: see profiles/r02_unchunk_ncu.txt.)
// chunk order: [frame block][tile row][tile column], each chunk whole (edge chunks padded), elements (c0, cy, cx) row-major;
// with SHUF the chunk holds sizeof(T) planes of n_elems bytes: plane k = k-th byte of every element.
constexpr int UNCHUNK_ROWS = 8;

template <typename T, int V, bool SHUF>
__global__ void __launch_bounds__(256) unchunk_to_f32_kernel(const uint8_t* __restrict__ chunks, float* __restrict__ out,
                                                             int ny, int nx, int c0, int cy, int cx, int gy, int gx, int first) {
    const unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (x >= (unsigned)nx) return;
    const unsigned y0 = blockIdx.y * UNCHUNK_ROWS;
    const unsigned f = blockIdx.z;
    const unsigned fg = f + (unsigned)first;
    const unsigned fb = fg / (unsigned)c0, fz = fg - fb * (unsigned)c0;
    unsigned iy = y0 / (unsigned)cy, yy = y0 - iy * (unsigned)cy;
    const unsigned ix = x / (unsigned)cx, xx = x - ix * (unsigned)cx;
    const int64_t n_elems = (int64_t)c0 * cy * cx;
    const int64_t chunk_bytes = n_elems * (int64_t)sizeof(T);
    const uint8_t* base = chunks + (((int64_t)fb * gy + iy) * gx + ix) * chunk_bytes;      // tile (fb, iy, ix)
    int64_t e = ((int64_t)fz * cy + yy) * cx + xx;                                         // element inside the tile
    float* o = out + ((int64_t)f * ny + y0) * nx + x;
    const int rows = min(UNCHUNK_ROWS, ny - (int)y0);
    // all loads of the thread first (UNCHUNK_ROWS independent requests in flight), then the conversions and stores
    T v[UNCHUNK_ROWS][V];
#pragma unroll
    for (int r = 0; r < UNCHUNK_ROWS; ++r) {
        if (r < rows) {
            if (!SHUF) {
                if (V == 4) *reinterpret_cast<typename Quad<T>::type*>(v[r]) = __ldcs(reinterpret_cast<const typename Quad<T>::type*>(base + e * sizeof(T)));
                else v[r][0] = *reinterpret_cast<const T*>(base + e * sizeof(T));
            } else {
                uint8_t b[sizeof(T)][V];
#pragma unroll
                for (int k = 0; k < (int)sizeof(T); ++k) {
                    if (V == 4) *reinterpret_cast<uint32_t*>(b[k]) = __ldcs(reinterpret_cast<const uint32_t*>(base + k * n_elems + e));
                    else b[k][0] = base[k * n_elems + e];
                }
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    uint8_t w[sizeof(T)];
#pragma unroll
                    for (int k = 0; k < (int)sizeof(T); ++k) w[k] = b[k][i];
                    memcpy(&v[r][i], w, sizeof(T));
                }
            }
            e += cx;
            if (++yy == (unsigned)cy) {                   // next tile row: same column of tiles, first row of the tile
                yy = 0;
                base += (int64_t)gx * chunk_bytes;
                e = (int64_t)fz * cy * cx + xx;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < UNCHUNK_ROWS; ++r) {
        if (r < rows) {
            if (V == 4) *reinterpret_cast<float4*>(o) = make_float4((float)v[r][0], (float)v[r][1], (float)v[r][2], (float)v[r][3]);
            else o[0] = (float)v[r][0];
            o += nx;
        }
    }
}

template <typename T>
int launch_unchunk(b4d_ctx* ctx, const void* chunks, int shuffled, int64_t n_frames, int ny, int nx, int c0, int cy, int cx,
                   int first, float* out) {
    const int gy = (ny + cy - 1) / cy, gx = (nx + cx - 1) / cx;
    const bool quad = (nx % 4 == 0) && (cx % 4 == 0);
    const int upr = quad ? nx / 4 : nx;                       // threads per row
    const dim3 grid((unsigned)((upr + 255) / 256), (unsigned)((ny + UNCHUNK_ROWS - 1) / UNCHUNK_ROWS), (unsigned)n_frames);
    const uint8_t* src = static_cast<const uint8_t*>(chunks);
    ProfScope ps(ctx, KC_SMALL);
    if (quad && shuffled) unchunk_to_f32_kernel<T, 4, true><<<grid, 256, 0, ctx->stream>>>(src, out, ny, nx, c0, cy, cx, gy, gx, first);
    else if (quad) unchunk_to_f32_kernel<T, 4, false><<<grid, 256, 0, ctx->stream>>>(src, out, ny, nx, c0, cy, cx, gy, gx, first);
    else if (shuffled) unchunk_to_f32_kernel<T, 1, true><<<grid, 256, 0, ctx->stream>>>(src, out, ny, nx, c0, cy, cx, gy, gx, first);
    else unchunk_to_f32_kernel<T, 1, false><<<grid, 256, 0, ctx->stream>>>(src, out, ny, nx, c0, cy, cx, gy, gx, first);
    B4D_LAUNCH_CHECK(ctx);
    return B4D_OK;
}

}  // namespace

extern "C" int b4d_inflate_caps(b4d_ctx* ctx, int* algo_mask, int64_t* max_bytes) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    const InflateCaps& c = inflate_caps(ctx->device);
    if (algo_mask) *algo_mask = c.submit ? c.algo_mask : 0;
    if (max_bytes) *max_bytes = c.submit ? c.max_bytes : 0;
    return B4D_OK;
}

extern "C" int b4d_inflate_batch(b4d_ctx* ctx, const void* src, const int64_t* src_offset, const int64_t* src_bytes,
                                 void* dst, int64_t dst_stride, uint32_t* actual, int64_t n) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!src || !src_offset || !src_bytes || !dst || !actual || n < 1 || dst_stride < 1)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_inflate_batch: bad arguments");
    const InflateCaps& c = inflate_caps(ctx->device);
    if (!c.submit || !(c.algo_mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE))
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_inflate_batch: this device / driver has no deflate decompression engine");
    if (dst_stride > c.max_bytes)
        return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_inflate_batch: chunks of %lld bytes exceed the engine's limit of %lld",
                        (long long)dst_stride, (long long)c.max_bytes);
    std::lock_guard<std::mutex> ring(g_caps_lock);
    std::vector<CUmemDecompressParams>& p = g_params[g_param_next];
    g_param_next = (g_param_next + 1) % PARAM_RING;
    p.assign((size_t)n, CUmemDecompressParams{});
    for (int64_t i = 0; i < n; ++i) {
        if (src_bytes[i] < 1 || src_offset[i] < 0)
            return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_inflate_batch: stream %lld has no bytes", (long long)i);
        p[i].srcNumBytes = (size_t)src_bytes[i];
        p[i].dstNumBytes = (size_t)dst_stride;
        p[i].dstActBytes = actual + i;
        p[i].src = static_cast<const uint8_t*>(src) + src_offset[i];
        p[i].dst = static_cast<uint8_t*>(dst) + i * dst_stride;
        p[i].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
    }
    size_t bad = 0;
    ProfScope ps(ctx, KC_SMALL);
    const CUresult rc = c.submit(p.data(), (size_t)n, 0, &bad, reinterpret_cast<CUstream>(ctx->stream));
    if (rc != CUDA_SUCCESS) {
        PfnGetErrorString es = nullptr;
        const char* msg = "?";
        if (driver_entry("cuGetErrorString", &es)) es(rc, &msg);
        return b4d_fail(ctx, B4D_ERR_CUDA, "cuMemBatchDecompressAsync failed (%d: %s) at stream %lld", (int)rc, msg, (long long)bad);
    }
    return B4D_OK;
}

extern "C" int b4d_unchunk_to_f32(b4d_ctx* ctx, const void* chunks, int dtype, int shuffled, int64_t n_frames, int ny, int nx,
                                  int c0, int cy, int cx, int first, float* out) {
    if (!ctx) return B4D_ERR_INVALID;
    B4dCall g(ctx);
    if (!chunks || !out || n_frames < 1 || ny < 1 || nx < 1 || c0 < 1 || cy < 1 || cx < 1 || first < 0 || first >= c0)
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_unchunk_to_f32: bad arguments");
    if ((reinterpret_cast<uintptr_t>(chunks) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return b4d_fail(ctx, B4D_ERR_INVALID, "b4d_unchunk_to_f32: pointers must be 16-byte aligned");
    if (n_frames > 65535 || ny > 65535) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_unchunk_to_f32: at most 65535 frames per call and rows per frame");
    switch (dtype) {
        case B4D_U8: return launch_unchunk<uint8_t>(ctx, chunks, shuffled, n_frames, ny, nx, c0, cy, cx, first, out);
        case B4D_U16: return launch_unchunk<uint16_t>(ctx, chunks, shuffled, n_frames, ny, nx, c0, cy, cx, first, out);
        case B4D_I16: return launch_unchunk<int16_t>(ctx, chunks, shuffled, n_frames, ny, nx, c0, cy, cx, first, out);
        case B4D_I32: return launch_unchunk<int32_t>(ctx, chunks, shuffled, n_frames, ny, nx, c0, cy, cx, first, out);
        case B4D_U32: return launch_unchunk<uint32_t>(ctx, chunks, shuffled, n_frames, ny, nx, c0, cy, cx, first, out);
        case B4D_F32: return launch_unchunk<float>(ctx, chunks, shuffled, n_frames, ny, nx, c0, cy, cx, first, out);
        default: return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, "b4d_unchunk_to_f32: dtype code %d", dtype);
    }
}
