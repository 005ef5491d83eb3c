#!/usr/bin/env python
"""Small end-to-end pass of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck).
usage: compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py [n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from barc4dip_b200 import engine, synth

sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 2048]
for n in sizes:
    T = 2
    stack, _ = synth.tracking_stack(T, n, grain=5.0, seed=3, integer_every=2)
    d = engine.as_stack(stack)
    engine.PhaseTracker(stack[0], (n, n), y0=0, x0=0)
    res = engine.stack_pipeline(d, tail_quantiles=(0.0005, 0.9995))
    ac, g = engine.autocorr2d(d, want_grain=True)
    tr = engine.PhaseTracker(stack[0], (n, n), y0=0, x0=0).track(d)
    x = engine.xcorr2d(d, d.flip(0).contiguous())
    torch.cuda.synchronize()
    print(n, "ok", float(res["autocorr"][0, n // 2, n // 2]), tr[1, :2], float(x.abs().max()))
m = 256
raw, flat, dark = synth.flatfield_case(4, m, seed=5, dead_frac=1e-3)
out = engine.temporal_moments(engine.as_stack(raw), return_device=True)
torch.cuda.synchronize()
print("temporal ok")
