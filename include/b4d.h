/*
 * b4d.h -- C ABI of libb4d.so, the B200 (sm_100a) implementation of the barc4dip
 * stack-analysis hot path.
 *
 * The reference (barc4dip, pure Python) has no FFI: its boundary for this path is the set of
 * Python functions listed in SURVEY.md 8(b).  Each entry point below names the reference
 * function(s) whose arithmetic it replaces (file:line under src/barc4dip/); the thin Python
 * package `barc4dip_b200` binds them with ctypes and keeps the reference's signatures.
 * INTEGRATION.md shows the binding a barc4dip maintainer would add.
 *
 * Conventions
 *   - every array argument is a DEVICE pointer to contiguous, C-ordered data unless the name
 *     ends in _host; frames are float32 (ny, nx), stacks float32 (T, ny, nx);
 *   - scalar result tables are float64, one row per frame;
 *   - the caller allocates all outputs; the context owns only scratch (twiddles, work buffers);
 *   - every call is asynchronous on the context's stream (b4d_set_stream), except the ones
 *     documented as synchronising;
 *   - return value: 0 = ok, negative = error (b4d_status); text via b4d_last_error().
 *     No C++ exception crosses this boundary.
 */
#ifndef B4D_H
#define B4D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b4d_ctx b4d_ctx;

typedef enum b4d_status {
    B4D_OK = 0,
    B4D_ERR_INVALID = -1,     /* bad argument (shape, null pointer, unsupported size) */
    B4D_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    B4D_ERR_UNSUPPORTED = -3, /* size/feature outside what the sm_100a kernels cover */
    B4D_ERR_NOMEM = -4
} b4d_status;

/* ---- context -------------------------------------------------------------------------- */
int b4d_create(int device, b4d_ctx** out);
int b4d_destroy(b4d_ctx* ctx);
/* cudaStream_t to launch on (0 = legacy default stream). */
int b4d_set_stream(b4d_ctx* ctx, void* cuda_stream);
int b4d_synchronize(b4d_ctx* ctx);
const char* b4d_last_error(b4d_ctx* ctx);
const char* b4d_version(void);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t b4d_launch_count(b4d_ctx* ctx);
int b4d_device_sm_count(b4d_ctx* ctx);

/*
 * Optional per-kernel-class timing with CUDA events on the context's stream (used by bench.py to report
 * the dominant kernel's duration inside the timed region).  b4d_profile_end synchronises the stream and
 * returns, per class, the summed device time in milliseconds and the number of bracketed launches.
 */
#define B4D_PROF_NCLASS 16
int b4d_profile_begin(b4d_ctx* ctx);
int b4d_profile_end(b4d_ctx* ctx, double* ms_per_class, int64_t* launches_per_class);
const char* b4d_profile_class_name(int klass);
/* Frames per internal batch of the FFT pipeline (0 = automatic: as many as a quarter of the free HBM holds, at most 128). */
int b4d_set_batch_frames(b4d_ctx* ctx, int64_t frames);
/*
 * Schedule of b4d_stack_pipeline inside a batch (replaces the per-frame loop of metrics/speckles.py:300-325,347-386).
 * sub_frames > 0: the big kernels run sub_frames frames at a time, the steps dealt round-robin to `lanes` (1..8)
 * independent lanes; in a lane, reduce -> rows -> columns run on one stream and the two inverse row passes follow the
 * column pass on two more, through ring_slots (1..8) ring slots of intermediates per lane, so that every intermediate
 * is consumed from L2 instead of HBM while the lanes fill one another's launch gaps.  keep = cache-policy bits of the
 * intermediates (1 stores stay in L2, 2 loads keep normal priority, 4 the column pass discards its consumed tiles).
 * The kernels honour `keep` only in builds made with -DB4D_KEEP_POLICY=1 (the run-time switch costs issue slots in every
 * memory access of the default schedule); the default build streams the intermediates both ways.
 * sub_frames = 0: whole batches, one kernel after the other.  use_graphs != 0: the second call with identical
 * arguments captures the batch's launches in a CUDA graph and later calls replay it (the host cannot issue ~10
 * launches per 1 - 2 frames as fast as the GPU runs them); calls bracketed by b4d_profile_begin/end always launch
 * directly.  A negative value of any argument selects the built-in default (environment: B4D_SUB, B4D_LANES,
 * B4D_SLOTS, B4D_KEEP, B4D_GRAPHS).  Results do not depend on the schedule.
 */
int b4d_set_schedule(b4d_ctx* ctx, int sub_frames, int lanes, int ring_slots, int keep, int use_graphs);
/*
 * Whole-batch schedule: the reduction pass and the forward row pass alternate pair_frames frames at a time, so that the
 * second finds in L2 the frames the first has just read from HBM (0 = one pass over the whole batch each; negative =
 * default / B4D_PAIR).  Results do not depend on it.
 */
int b4d_set_pairing(b4d_ctx* ctx, int pair_frames);

/* Plain device-memory helpers so that a non-Python host can drive the library. */
int b4d_malloc(b4d_ctx* ctx, size_t bytes, void** out);
int b4d_free(b4d_ctx* ctx, void* p);
int b4d_memcpy_h2d(b4d_ctx* ctx, void* dst, const void* src_host, size_t bytes);
int b4d_memcpy_d2h(b4d_ctx* ctx, void* dst_host, const void* src, size_t bytes); /* synchronises */

/*
 * Stack ingestion: integer detector frames (uint8 / uint16 / int16 / int32 / uint32) are uploaded in their own type
 * (half or a quarter of the PCIe bytes of float32) and widened to float32 on the device, the way the reference's
 * entry points cast integer images on the host (signal/tracking.py:299-305 -> float32; the metrics cast further to
 * float64, which this path evaluates in float32 anyway). src and dst: device pointers, 16-byte
 * aligned; n elements.
 */
enum b4d_dtype { B4D_U8 = 0, B4D_U16 = 1, B4D_I16 = 2, B4D_I32 = 3, B4D_U32 = 4, B4D_F32 = 5 /* b4d_unchunk_to_f32 only */ };
int b4d_cast_to_f32(b4d_ctx* ctx, const void* src, int dtype, float* dst, int64_t n);

/*
 * Compressed stacks: the reference's files are HDF5 datasets of gzip-4 chunks (io/h5.py:204-210, written; :80 `dset[()]`,
 * read and inflated on the host by h5py). Here the deflate streams are uploaded as stored and inflated by the GPU's
 * hardware decompression engine (cuMemBatchDecompressAsync; B200: deflate / snappy / lz4, 4 MiB per stream).
 *
 * b4d_inflate_caps: *algo_mask = CUmemDecompressAlgorithm bits of the device (0: no engine / driver older than 12.8),
 *   *max_bytes = largest single stream's output.
 * b4d_inflate_batch: n RAW deflate streams (RFC 1951: a zlib stream minus its 2-byte header and 4-byte Adler-32)
 *   at src + src_offset[i], src_bytes[i] long, inflate to dst + i * dst_stride; actual[i] receives the bytes written.
 *   src, dst, actual: device memory from cudaMalloc (16-byte aligned offsets); src_offset, src_bytes: host arrays.
 *   One batch is one unit of stream-ordered work on the context's stream. The engine does not verify its input: a
 *   malformed stream surfaces as a sticky CUDA error at the next synchronisation, so callers check the zlib header
 *   before stripping it and compare actual[] with the expected size afterwards.
 * b4d_unchunk_to_f32: undoes the rest of the HDF5 storage on the way to float32 frames: `chunks` holds whole inflated
 *   chunks of (c0, cy, cx) elements of `dtype` in [frame block][tile row][tile column] order (edge chunks padded, as
 *   stored), byte-shuffled per chunk when `shuffled` (HDF5 filter 2); out[f][y][x], f < n_frames, is element
 *   (first + f) of the frame axis of those chunks (first < c0: a frame range may start inside a chunk).
 */
int b4d_inflate_caps(b4d_ctx* ctx, int* algo_mask, int64_t* max_bytes);
int b4d_inflate_batch(b4d_ctx* ctx, const void* src, const int64_t* src_offset, const int64_t* src_bytes, void* dst,
                      int64_t dst_stride, uint32_t* actual, int64_t n);
int b4d_unchunk_to_f32(b4d_ctx* ctx, const void* chunks, int dtype, int shuffled, int64_t n_frames, int ny, int nx, int c0,
                       int cy, int cx, int first, float* out);

/* ---- per-frame single-pass reductions ------------------------------------------------- */
/*
 * One streaming pass per frame producing everything the reference's scalar metrics need:
 *   distribution_moments      metrics/statistics.py:17-125
 *   tenengrad                 metrics/sharpness.py:405-476  (scipy.ndimage.sobel, mode="reflect")
 *   laplacian_variance        metrics/sharpness.py:482-530  (scipy.ndimage.laplace, mode="reflect")
 *   amplitude (visibility)    metrics/speckles.py:636-645   (nanmean / nanstd)
 * Optional fused flat field: if gain != NULL the pixel fed to the metrics is
 * (raw - dark) * gain  (gain = scale/(flat-dark), 0 on bad pixels; preprocessing/normalize.py:104-131).
 *
 * out: T rows of B4D_FR_NCOLS doubles, columns B4D_FR_*.
 */
enum {
    B4D_FR_COUNT = 0,   /* number of finite pixels                         */
    B4D_FR_MEAN = 1,    /* mean over finite pixels                          */
    B4D_FR_M2 = 2,      /* central moments (biased): variance               */
    B4D_FR_M3 = 3,
    B4D_FR_M4 = 4,
    B4D_FR_NZERO = 5,   /* #finite pixels with |x| <= zero_eps              */
    B4D_FR_NSAT = 6,    /* #finite pixels with x >= sat_value (0 if NaN)    */
    B4D_FR_SGX2 = 7,    /* sum over finite pixels of sobel_x^2              */
    B4D_FR_SGY2 = 8,
    B4D_FR_SLAP = 9,    /* sum / sum of squares of the 5-point Laplacian    */
    B4D_FR_SLAP2 = 10,
    B4D_FR_NPIX = 11,   /* ny*nx                                            */
    B4D_FR_NNAN = 12,   /* number of NaN pixels (infinities = NPIX - COUNT - NNAN) */
    B4D_FR_NCOLS = 13
};
int b4d_frame_reductions(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                         const float* gain, const float* dark,
                         double sat_value, double zero_eps, double* out);

/*
 * b4d_frame_reductions plus, from the same pass, the order statistics that bracket two tail percentiles
 * (amplitude(): np.nanpercentile(img, 0.05 / 99.95), metrics/speckles.py:647).  quant_out / nvalid_out as in
 * b4d_stack_pipeline.  Meant for small tails (each side of the frame up to ~0.3 % of the pixels).
 */
int b4d_frame_reductions_tails(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                               const float* gain, const float* dark, double sat_value, double zero_eps,
                               double q_lo, double q_hi, double* out, float* quant_out, int64_t* nvalid_out);

/* ---- exact order statistics ------------------------------------------------------------ */
/*
 * For each frame, the k-th smallest finite value (0-based ranks, ascending) for every k in
 * ranks[0..n_ranks) -- the primitive behind np.nanpercentile (utils/range.py:44-54, used by
 * amplitude, metrics/speckles.py:647) and np.median (signal/tracking.py:319,
 * preprocessing/normalize.py:110,126).  NaNs are excluded as in nanpercentile; n_valid[t] gets the
 * number of non-NaN values so that the host can place the fractional rank.
 *   ranks: HOST array of n_ranks int64 (ranks are clamped to [0, n_valid-1] per frame);
 *   if ranks_from_quantiles != NULL (HOST, n_ranks doubles in [0,1]) ranks are derived per frame
 *   as floor(q*(n_valid-1)) and floor(q*(n_valid-1))+1 -> out has 2*n_ranks columns.
 *   out: DEVICE float32 (n_frames, n_cols); n_valid: DEVICE int64 (n_frames) or NULL.
 */
int b4d_select_ranks(b4d_ctx* ctx, const float* stack, int64_t n_frames, int64_t frame_elems,
                     const double* quantiles_host, int n_q, int use_abs,
                     float* out, int64_t* n_valid);

/* ---- flat field ------------------------------------------------------------------------ */
/*
 * flat_field_correction, preprocessing/normalize.py:12-145 (without the median repair):
 *   out = ((img - dark) / den_safe) * scale ; out[bad] = 0,   bad = (flat - dark) <= eps.
 * Same float32 operation order as the reference, so the result is bit-identical.
 * dark may be NULL (zeros); scale_value is the already-resolved multiplier (1.0 for scale="none").
 */
int b4d_flat_field(b4d_ctx* ctx, const float* images, int64_t n_frames, int ny, int nx,
                   const float* flat, const float* dark, float eps, float scale_value,
                   int apply_scale, float* out);
/*
 * bad_pixel_removal=True (preprocessing/normalize.py:134-140): frames[:, bad] = median_filter(frames, 3x3)[:, bad] with
 * scipy's 'reflect' border, in place on the output of b4d_flat_field (bad pixels are 0 there); same flat / dark / eps.
 */
int b4d_bad_pixel_repair(b4d_ctx* ctx, float* frames, int64_t n_frames, int ny, int nx,
                         const float* flat, const float* dark, float eps);
/* gain[p] = bad ? 0 : scale/(flat-dark): the per-pixel multiplier the fused loaders use. */
int b4d_flat_gain(b4d_ctx* ctx, const float* flat, const float* dark, int ny, int nx,
                  float eps, float scale_value, float* gain);
/* den = flat - dark and its count of valid (den > eps) pixels (helper for the medians). */
int b4d_sub(b4d_ctx* ctx, const float* a, const float* b, int64_t n, float* out);

/* ---- per-pixel temporal moments (lifts statistics.py:75-81 along t) - */
/*
 * Accumulate shifted power sums  S_k(p) += sum_t (x_t(p) - shift(p))^k , k = 1..4, over the
 * n_frames given.  sums: DEVICE float64 (4, ny, nx), caller-zeroed before the first chunk;
 * shift: DEVICE float32 (ny, nx), identical on every rank so that sums add across GPUs.
 * Optional fused flat field as in b4d_frame_reductions.
 */
int b4d_temporal_accumulate(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                            const float* gain, const float* dark, const float* shift, double* sums);
/* shift(p) = mean over the first n_frames frames of the (corrected) pixel: a pilot estimate. */
int b4d_temporal_pilot(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                       const float* gain, const float* dark, float* shift);
/* maps: DEVICE float64 (5, ny, nx) = mean, std, variance, skewness, excess kurtosis. */
int b4d_temporal_finalize(b4d_ctx* ctx, const double* sums, const float* shift, int64_t n_total,
                          int ny, int nx, double* maps);

/* ---- 2-D FFT family ---------------------------------------------------------------------- */
/*
 * Frame sizes. Powers of two in [B4D_FFT_MIN, B4D_FFT_MAX] run the hot-path kernels. Every entry point below also accepts
 * ANY sides in [2, B4D_DFT_MAX] (the 227 / 228-pixel sub-tiles of the reference's tiling executor, metrics/common.py:278-378,
 * detectors that are not 2^k wide or wider than 2048 pixels, e.g. 2560 x 2160; the reference is size-agnostic,
 * signal/fft.py:236): those go through Bluestein's algorithm on the same FFT core (csrc/generic_dft.cuh), with the
 * stack pipeline composed from the stand-alone paths. Larger sides return B4D_ERR_UNSUPPORTED (the Python layer raises;
 * there is no CPU fallback).
 */
#define B4D_FFT_MIN 128
#define B4D_FFT_MAX 2048
#define B4D_DFT_MAX 4096

/* fft2d (signal/fft.py:198-237): out = fftshift(fft2(frame)), complex64 interleaved (ny, nx). */
int b4d_fft2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, float* out_c64);

/* ifft2d (signal/fft.py:240-258): out = ifft2(ifftshift(F)), complex64 in and out (ny, nx), any sides in [2, 4096]
 * (chirp-z path; a building block and test hook, not a hot-path kernel). */
int b4d_ifft2d(b4d_ctx* ctx, const float* spec_c64, int64_t n_frames, int ny, int nx, float* out_c64);

/*
 * psd2d (signal/fft.py:261-309): out = |fftshift(fft2(frame - sub))|^2 * scale_factor, float32.
 * sub_mean != 0 subtracts the frame mean first (bandwidth / spectral_entropy call sites,
 * metrics/speckles.py:744-748, metrics/sharpness.py:593-596); zero_dc != 0 clears the DC bin.
 * spectral (nullable): DEVICE float64 (n_frames, B4D_SP_NCOLS) sums for bandwidth()/spectral_entropy().
 */
enum {
    B4D_SP_TOTAL = 0,   /* sum P                 over FR <= f_max (DC excluded)          */
    B4D_SP_FX2 = 1,     /* sum fx^2 P            (fx, fy in cycles/pixel)                */
    B4D_SP_FY2 = 2,
    B4D_SP_P2 = 3,      /* sum P^2               over the same mask                      */
    B4D_SP_ALL = 4,     /* sum P                 over all bins except DC                 */
    B4D_SP_PLOGP = 5,   /* sum P ln P            over all bins except DC (P > 0)         */
    B4D_SP_F95 = 6,     /* radius (cycles/pixel) where the radius-sorted cumulative PSD first reaches 0.95 */
    B4D_SP_NCOLS = 8
};
int b4d_psd2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
              float scale_factor, int sub_mean, int zero_dc, float* out_psd, double* spectral);

/*
 * autocorr2d (signal/corr.py:256-320 via xcorr2d :169-253) of each frame, float32 output:
 *   remove_mean / standardize / normalize_peak as in the reference; zero lag at (ny//2, nx//2).
 * grain_out (nullable): DEVICE float64 (n_frames, 4) = lx, ly, leq, lx/ly computed from the map
 *   exactly as grain() does (metrics/speckles.py:546-575, maths/stats.py:9-155, maths/radial.py:101-169)
 *   with threshold `fraction`.
 * out_ac may be NULL when only grain_out is wanted (the map then lives in scratch).
 */
int b4d_autocorr2d(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                   int remove_mean, int standardize, int normalize_peak,
                   float* out_ac, double fraction, double* grain_out);

/*
 * template_matching (signal/tracking.py:82-188, cv2.matchTemplate TM_CCOEFF_NORMED on the z-scored template): for every
 * frame the normalised cross-correlation map of shape (ny-h+1, nx-w+1), its first-occurrence argmax, the reference's
 * 3x3 Taylor step and (dy, dx, peak, snr) with dy = peak_y + (h-1)/2 - ref_y, snr = |peak| / (median|map| + eps).
 * tpl: DEVICE float32, raw (the z-score is applied here): one (h, w) template for every frame (per_frame = 0) or
 * (n_frames, h, w), template t matched against frame t (per_frame = 1: the incremental tracking of speckle_stack_stats,
 * metrics/speckles.py:373-384). ref_y, ref_x: centre of the template's reference position, (start + stop - 1) / 2 of
 * its slices. Power-of-two frames. out: DEVICE float64 (n_frames, 4).
 */
int b4d_template_match(b4d_ctx* ctx, const float* tpl, int per_frame, int h, int w, const float* stack, int64_t n_frames,
                       int ny, int nx, double ref_y, double ref_x, int subpixel, double eps, double* out);

/* xcorr2d (signal/corr.py:169-253) of frame pairs (a[t], b[t]); real float32 output. */
int b4d_xcorr2d(b4d_ctx* ctx, const float* a, const float* b, int64_t n_frames, int ny, int nx,
                int remove_mean, int standardize, int normalize_peak, float* out);

/*
 * phase_correlation, backend="internal" (signal/tracking.py:192-297, helpers :299-375).
 * b4d_phase_set_reference z-scores the (h, w) template over its own pixels, embeds it at
 * (y0, x0) in a zero (ny, nx) frame and caches conj(FFT) in the context.
 * b4d_phase_track correlates every frame of the stack against it:
 *   out: DEVICE float64 (n_frames, 4) = dy, dx, peak, snr  (sub-pixel terms applied -- with the
 *   reference's swapped order -- when subpixel != 0).
 */
/*
 * Tracker SNR = |peak| / (median(|corr|) + eps) (signal/tracking.py:314-321).  By default the exact median is taken
 * inside the inverse row pass, without ever writing the |corr| map: a few sample rows fix a bracket around the
 * median, every other value is counted / collected against it from registers.  A frame whose bracket missed reports
 * snr = NaN (dy, dx, peak stay valid); redo such frames with b4d_set_fused_median(ctx, 0), which selects the
 * map-based path (|corr| map written, stand-alone exact select).
 */
int b4d_set_fused_median(b4d_ctx* ctx, int on);
int b4d_phase_set_reference(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx,
                            int y0, int x0, double eps);
int b4d_phase_track(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                    int subpixel, double eps, double* out);
/*
 * The same tracker with a reference its creator OWNS (signal/tracking.py:192-297 takes the template as an argument of
 * every call; the reference's callers drive it from joblib threads, metrics/speckles.py:323): any number of
 * references may be alive per context and frame shape, none is re-targeted by another b4d_phase_set_reference /
 * b4d_phase_reference_create.  b4d_phase_track_ref: map_median != 0 selects the map-based median for this call only
 * (the redo path of frames whose fused bracket missed).  Destroy a reference only after the work that reads it.
 */
typedef struct b4d_ref b4d_ref;
int b4d_phase_reference_create(b4d_ctx* ctx, const float* tpl, int h, int w, int ny, int nx,
                               int y0, int x0, double eps, b4d_ref** out);
int b4d_phase_reference_destroy(b4d_ctx* ctx, b4d_ref* ref);
int b4d_phase_track_ref(b4d_ctx* ctx, const b4d_ref* ref, const float* stack, int64_t n_frames, int ny, int nx,
                        int subpixel, double eps, int map_median, double* out);

/*
 * The fused stack pass of the north-star pipeline: one call per chunk of frames produces
 *   fr_out    (n_frames, B4D_FR_NCOLS)  frame reductions           (nullable)
 *   quant_out (n_frames, 4) float32     the two order statistics bracketing the q_lo and the q_hi percentile
 *                                       (numpy 'linear' neighbours v[floor(h)], v[floor(h)+1]; amplitude(),
 *                                       metrics/speckles.py:647, utils/range.py:51-54), collected inside the
 *                                       reduction pass; nvalid_out (n_frames) int64 = non-NaN pixels, or -1 for a
 *                                       frame whose tails were not resolved (run b4d_select_ranks on it). (nullable)
 *   psd_out   (n_frames, ny, nx)        psd2d                      (nullable)
 *   ac_out    (n_frames, ny, nx)        autocorr2d (defaults)      (nullable)
 *   grain_out (n_frames, 4)             grain widths               (nullable, needs autocorr)
 *   track_out (n_frames, 4)             phase correlation vs the cached reference (nullable)
 */
int b4d_stack_pipeline(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx,
                       const float* gain, const float* dark, double sat_value, double zero_eps,
                       float psd_scale, int subpixel, double eps, double q_lo, double q_hi,
                       double* fr_out, float* quant_out, int64_t* nvalid_out, float* psd_out, float* ac_out,
                       double* grain_out, double* track_out);

/*
 * b4d_stack_pipeline against an owned reference (nullable when track_out is null), plus
 *   spectral_out (n_frames, B4D_SP_NCOLS)  the spectral sums of bandwidth() / spectral_entropy() (metrics/speckles.py:740-796,
 *                                       metrics/sharpness.py:581-629) taken in the column pass of the SAME forward transform
 *                                       that feeds the PSD map, the autocorrelation and the tracker (nullable; needs
 *                                       ac_out or grain_out and power-of-two sides; f95 on square frames only).  The
 *                                       sums are of P * psd_scale with the DC bin left out; every metric derived from
 *                                       them is invariant to that scale.
 */
int b4d_stack_pipeline_ref(b4d_ctx* ctx, const b4d_ref* ref, const float* stack, int64_t n_frames, int ny, int nx,
                           const float* gain, const float* dark, double sat_value, double zero_eps,
                           float psd_scale, int subpixel, double eps, double q_lo, double q_hi,
                           double* fr_out, float* quant_out, int64_t* nvalid_out, float* psd_out, float* ac_out,
                           double* grain_out, double* track_out, double* spectral_out);

#ifdef __cplusplus
}
#endif
#endif /* B4D_H */
