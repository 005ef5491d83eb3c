#!/usr/bin/env python
"""Where the end-to-end time of StackAnalyzer.run goes: wall time per call for several chunk sizes, next to the bare
pinned host->device copy of the same stack."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from barc4dip_b200 import synth
from barc4dip_b200.pipeline import StackAnalyzer

F, n = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 2048
dev = torch.device("cuda:0")
base = torch.from_numpy(synth.speckle_frame(n, grain=6.0, seed=0)).to(dev)
stack = base[None].repeat(F, 1, 1) + 10.0 * torch.randn((F, n, n), device=dev)
host = torch.empty((F, n, n), dtype=torch.float32, pin_memory=True)
host.copy_(stack)
del stack
torch.cuda.synchronize()
buf = torch.empty((16, n, n), device=dev)
for rep in range(3):
    t0 = time.perf_counter()
    for a in range(0, F, 16):
        buf.copy_(host[a:a + 16], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"bare H2D: {dt * 1e3:.1f} ms  {F * n * n * 4 / dt / 1e9:.1f} GB/s")
for chunk in (16, 8, 4):
    an = StackAnalyzer((n, n), device=0, chunk_frames=chunk, want_maps=True, want_contrast=True)
    an.set_reference(host[0].clone())
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = an.run(host, keep_maps_on_device=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"chunk {chunk:3d} run {rep}: {dt * 1e3:7.1f} ms  {F / dt:7.0f} frames/s")
        del out
    del an
