// Temporary stubs for the FFT family until fft.cu lands (keeps the ABI complete).
#include "common.cuh"
void b4d_fft_release(b4d_ctx*) {}
#define STUB(name) return b4d_fail(ctx, B4D_ERR_UNSUPPORTED, name ": not built yet")
extern "C" int b4d_fft2d(b4d_ctx* ctx, const float*, int64_t, int, int, float*) { STUB("b4d_fft2d"); }
extern "C" int b4d_psd2d(b4d_ctx* ctx, const float*, int64_t, int, int, float, int, int, float*, double*) { STUB("b4d_psd2d"); }
extern "C" int b4d_autocorr2d(b4d_ctx* ctx, const float*, int64_t, int, int, int, int, int, float*, double, double*) { STUB("b4d_autocorr2d"); }
extern "C" int b4d_xcorr2d(b4d_ctx* ctx, const float*, const float*, int64_t, int, int, int, int, int, float*) { STUB("b4d_xcorr2d"); }
extern "C" int b4d_phase_set_reference(b4d_ctx* ctx, const float*, int, int, int, int, int, int, double) { STUB("b4d_phase_set_reference"); }
extern "C" int b4d_phase_track(b4d_ctx* ctx, const float*, int64_t, int, int, int, double, double*) { STUB("b4d_phase_track"); }
extern "C" int b4d_stack_pipeline(b4d_ctx* ctx, const float*, int64_t, int, int, const float*, const float*, double, double, float, int, double, double*, float*, float*, double*, double*) { STUB("b4d_stack_pipeline"); }
