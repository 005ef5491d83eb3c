#!/usr/bin/env python
"""Marginal cost of the pipeline's branches: kernel-class times (ms per call) for several output selections."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from barc4dip_b200 import engine, synth
from barc4dip_b200._lib import get_context

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 2048
dev = torch.device("cuda:0")
base = torch.from_numpy(synth.speckle_frame(n, grain=6.0, seed=0)).to(dev)
stack = base[None].repeat(F, 1, 1) + 10.0 * torch.randn((F, n, n), device=dev)
engine.PhaseTracker(stack[0], (n, n), y0=0, x0=0, device=0)
ctx = get_context(0)
ctx.set_batch_frames(F)
psd = torch.empty((F, n, n), device=dev)
ac = torch.empty((F, n, n), device=dev)
cases = {
    "full": dict(want_psd=True, want_autocorr=True, want_grain=True, want_tracking=True, tail_quantiles=(0.0005, 0.9995)),
    "no_psd": dict(want_psd=False, want_autocorr=True, want_grain=True, want_tracking=True),
    "no_track": dict(want_psd=True, want_autocorr=True, want_grain=True, want_tracking=False),
    "no_ac": dict(want_psd=True, want_autocorr=False, want_grain=False, want_tracking=True),
    "psd_only": dict(want_psd=True, want_autocorr=False, want_grain=False, want_tracking=False),
    "ac_only": dict(want_psd=False, want_autocorr=True, want_grain=False, want_tracking=False),
    "track_only": dict(want_psd=False, want_autocorr=False, want_grain=False, want_tracking=True),
}
for name, kw in cases.items():
    for it in range(3):
        if it == 2:
            ctx.profile_begin()
        engine.stack_pipeline(stack, psd_out=psd if kw["want_psd"] else None, ac_out=ac if kw["want_autocorr"] else None, **kw)
    prof = ctx.profile_end()
    tot = sum(v[0] for v in prof.values())
    print(f"{name:10s} total {tot:6.2f} ms ({tot / F * 1e3:5.1f} us/frame) ", {k: round(v[0], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[0] > 0.005})
