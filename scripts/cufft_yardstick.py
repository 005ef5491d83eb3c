#!/usr/bin/env python
"""Yardstick only (not on the product path): how long does cuFFT (through torch.fft) take for the bare TRANSFORMS of the
fused pipeline -- one real-to-complex 2-D FFT per frame, one complex-to-real inverse for the tracker, one for the
autocorrelation -- on the same 128 x 2048^2 batch, without any of the epilogues (|F|^2, whitening product, census,
argmax, shifted map stores, reductions)? Compared with the FFT kernels of libb4d.so on the same box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

n, F = 2048, 128
dev = torch.device("cuda:0")
x = torch.randn((F, n, n), device=dev)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


spec = torch.fft.rfft2(x)
out = {"frames": F, "size": n}
out["cufft_rfft2_ms"] = timed(lambda: torch.fft.rfft2(x))
out["cufft_irfft2_ms"] = timed(lambda: torch.fft.irfft2(spec, s=(n, n)))
out["cufft_transforms_of_the_pipeline_ms"] = out["cufft_rfft2_ms"] + 2 * out["cufft_irfft2_ms"]
del spec
torch.cuda.empty_cache()

from barc4dip_b200 import engine
from barc4dip_b200._lib import get_context
ctx = get_context(0)
tr = engine.PhaseTracker(x[0], (n, n), y0=0, x0=0)
psd = torch.empty((F, n, n), device=dev)
ac = torch.empty((F, n, n), device=dev)
run = lambda: engine.stack_pipeline(x, tracker=tr, want_reductions=True, psd_out=psd, ac_out=ac, tail_quantiles=None)
timed(run, reps=2)
ctx.profile_begin()
ms = timed(run, reps=10, warm=0)
prof = ctx.profile_end()
out["b4d_pipeline_ms"] = ms
out["b4d_fft_kernels_ms"] = {k: prof[k][0] / 10 for k in ("rows_fwd", "cols", "rows_inv", "rows_inv_ac")}
out["b4d_fft_kernels_sum_ms"] = sum(out["b4d_fft_kernels_ms"].values())
out["note"] = ("cuFFT: transforms only, each materialising its full complex / real result in HBM; b4d: the four FFT kernels "
               "with every epilogue fused (PSD map, whitened product against the reference spectrum, packed autocorrelation, "
               "argmax, exact-median census, shifted map stores); event spans of the two inverse row passes overlap (two streams)")
print(json.dumps(out))
