// Internal interface of the frame-reduction pass (reduce.cu) for the other translation units of libb4d.so.
#pragma once

#include "common.cuh"

struct FrTails {              // optional fused tail percentiles (amplitude contrast)
    double q_lo, q_hi;
    float* quant_out;          // (T, 4): (v[lo], v[hi]) of q_lo, then of q_hi
    int64_t* nvalid_out;       // (T): non-NaN pixels, -1 = unresolved (use b4d_select_ranks for that frame)
};

// One reduction pass over n_frames frames, split so that its streaming kernel can be issued a few frames at a time.
struct FrPlan {
    const float* stack; const float* gain; const float* dark;
    int64_t n_frames; int ny, nx;
    double* out;
    FrTails tails; bool has_tails, fuse_tails, vec;
    float* pilot; float* thr;
    double* partials; unsigned* cnt; unsigned* eq; int* flag; float* cand;
    unsigned gcap; int nstrips, nitems, nblocks;
    float sat, zeps; int has_sat;
};

int b4d_fr_begin(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain, const float* dark,
                 double sat_value, double zero_eps, double* out, const FrTails* tails, float* pilot_out, FrPlan* pl);
int b4d_fr_range(b4d_ctx* ctx, const FrPlan& pl, int64_t t0, int64_t tc);
int b4d_fr_end(b4d_ctx* ctx, const FrPlan& pl);

int b4d_frame_reductions_ex(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                            const float* dark, double sat_value, double zero_eps, double* out, const FrTails* tails,
                            float* pilot_out);
int b4d_frame_reductions_nolock(b4d_ctx* ctx, const float* stack, int64_t n_frames, int ny, int nx, const float* gain,
                                const float* dark, double sat_value, double zero_eps, double* out);
int b4d_frame_pilot_launch(b4d_ctx* ctx, const float* stack, int64_t T, int64_t npix, const float* gain,
                           const float* dark, float* pilot);
