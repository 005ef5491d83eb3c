import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import barc4dip_b200 as dip
from barc4dip_b200 import engine, stack as blocks, synth
frame = synth.speckle_frame(2048, grain=6.0, seed=0)
def T(name, fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); print(f"{name:40s} {(time.perf_counter() - t0) / reps * 1e3:8.3f} ms", flush=True)
    return r
d = T("as_stack (H2D)", lambda: engine.as_stack(np.ascontiguousarray(frame[::-1])))
fb = T("FusedBlocks(keep_map)", lambda: blocks.FusedBlocks(d, saturation_value=65535.0, eps=1e-6, keep_map=True))
T("FusedBlocks(no map)", lambda: blocks.FusedBlocks(d, saturation_value=65535.0, eps=1e-6, keep_map=False))
T("fb.amplitude()", lambda: fb.amplitude())
T("fb.bandwidth()+grain+moments", lambda: (fb.bandwidth(), fb.grain(), fb.moments()))
T("map .cpu().numpy().astype(f64)", lambda: fb.ac[0].cpu().numpy().astype(np.float64))
def composed():
    table = engine.frame_reductions(d, saturation_value=65535.0, eps=1e-6)
    a = blocks.amplitude_block(d, table); g, ac = blocks.grain_block(d, table=table, return_map=True)
    m = blocks.moments_block(table, 65535.0); b = blocks.bandwidth_block(d, table=table)
    return a, g, m, b
T("composed blocks (round 1 path)", composed)
T("speckle_stats(tiles=False)", lambda: dip.metrics.speckle_stats(frame, tiles=False, verbose=False))
T("sharpness_stats(tiles=False)", lambda: dip.metrics.sharpness_stats(frame, tiles=False, verbose=False))
