#!/usr/bin/env python
"""Frames/s of the BASELINE.json configs C2..C5 on one B200 (HBM-resident synthetic stacks, CUDA-event timing),
each with its algorithmic-bytes roofline fraction. One JSON line per config.
usage: python scripts/configs_bench.py [frames_2048] [frames_1024]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from barc4dip_b200 import engine, synth

F2 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
F1 = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
HBM = float(peaks["hbm_gbs"])


def timed(fn, reps=5, warm=3):
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


def line(name, frames, ms, algo_bytes_per_frame, note):
    fps = frames / (ms / 1e3)
    print(json.dumps({"config": name, "frames": frames, "ms": round(ms, 3), "frames_per_s": round(fps, 1),
                      "algorithmic_MB_per_frame": algo_bytes_per_frame / 1e6,
                      "hbm_frac_of_measured": round(fps * algo_bytes_per_frame / 1e9 / HBM, 4), "note": note}), flush=True)


n = 2048
base = torch.from_numpy(synth.speckle_frame(n, grain=6.0, seed=0)).to(dev)
stack = base[None].repeat(F2, 1, 1) + 10.0 * torch.randn((F2, n, n), device=dev)
B = n * n * 4
# C2: sharpness scan (moments + Tenengrad + Laplacian variance per frame)
tab = torch.empty((F2, 13), dtype=torch.float64, device=dev)
ms = timed(lambda: engine.frame_reductions(stack, return_device=True))
line("C2 sharpness scan 2048^2 (frame reductions)", F2, ms, B, "one streaming read per frame")
# C3: PSD + autocorrelation maps
psd = torch.empty((F2, n, n), device=dev)
ac = torch.empty((F2, n, n), device=dev)
ms = timed(lambda: engine.stack_pipeline(stack, want_reductions=False, want_psd=True, want_autocorr=True, want_grain=True,
                                         want_tracking=False, psd_out=psd, ac_out=ac))
line("C3 PSD + autocorrelation 2048^2", F2, ms, 3 * B, "read frame + write PSD + write autocorrelation")
# C4: tracking against a broadcast reference
tr = engine.PhaseTracker(stack[0], (n, n), y0=0, x0=0, device=0)
ms = timed(lambda: tr.track(stack, return_device=True))
line("C4 phase-correlation tracking 2048^2", F2, ms, B, "frame read once; reference spectrum L2-resident")
del stack, psd, ac
# C5: temporal moments at 1024^2 with fused flat field
m = 1024
raw, flat, dark = synth.flatfield_case(8, m, seed=5, dead_frac=1e-4)
d_flat, d_dark = torch.from_numpy(flat).to(dev), torch.from_numpy(dark).to(dev)
den = flat - dark
eps = 1e-6 * float(np.median(den))
gain = engine.flat_gain(d_flat, d_dark, eps=eps, scale_value=float(np.median(den[den > eps])))
st1 = torch.from_numpy(raw).to(dev).repeat(F1 // 8, 1, 1)
ms = timed(lambda: engine.temporal_moments(st1, gain=gain, dark=d_dark, return_device=True))
line("C5 temporal moments 1024^2 + flat field", st1.shape[0], ms, m * m * 4, "one streaming read per frame")
del st1
# C1: one 2048^2 frame through the barc4dip-speckles path (speckle_stats: stats, amplitude, grain, bandwidth), full frame
# and with the reference's default 9x9 sub-tile grid (227 / 228 px tiles: Bluestein FFT path); latency, host array in,
# result dict out. CPU: the oracle port of the same call on this box (one pass each).
import time
import barc4dip_b200 as dip
from oracle import ref_numpy as orc
frame = synth.speckle_frame(2048, grain=6.0, seed=0)
for tiles in (False, True):
    for _ in range(2):
        dip.metrics.speckle_stats(frame, tiles=tiles, verbose=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        dip.metrics.speckle_stats(frame, tiles=tiles, verbose=False)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    flipped = frame[::-1, :]
    orc.amplitude(flipped); orc.grain(flipped); orc.distribution_moments(flipped); orc.bandwidth(flipped)
    if tiles:
        orc.speckle_tiles(frame)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"config": f"C1 speckle_stats 2048^2 tiles={tiles}", "gpu_ms": round(gpu_ms, 2), "cpu_port_ms": round(cpu_ms, 1),
                      "speedup": round(cpu_ms / gpu_ms, 1), "note": "latency of one call, host numpy in, result dict out; CPU = oracle port, 1 thread"}), flush=True)
