"""What the first next() of DeviceInflater spends its ~20 ms on: the steps of __iter__ replayed by hand with timers."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from barc4dip_b200 import synth
    from barc4dip_b200._lib import ptr
    from barc4dip_b200.io import hdf5, stream
    n, T = 2048, 64
    rng = np.random.default_rng(0)
    base = synth.speckle_frame(n, grain=6.0, seed=0)
    stack = np.stack([rng.poisson(np.roll(base, t, axis=1)).clip(0, 65535).astype(np.uint16) for t in range(T)])
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "s.h5")
        hdf5.write_stack(path, stack)
        for rep in range(3):
            with hdf5.H5File(path) as f:
                marks = [("start", time.perf_counter())]

                def mark(name):
                    marks.append((name, time.perf_counter()))
                d = f["entry_0000/measurement/data"]
                mark("dataset")
                src = stream.DeviceInflater(d, block_frames=32)
                mark("inflater")
                bufs = src._buffers()
                mark("buffers")
                pins = [b["pin"].numpy() for b in bufs]
                s_in = torch.cuda.Stream()
                mark("stream")
                ctx, lib = src.ctx, src.ctx.lib
                for k in range(2):
                    lo, hi, recs = src.blocks[k]
                    offs, sizes, total = src._pack(recs, pins[k])
                    mark(f"pack{k} ({hi - lo} frames)")
                    b = bufs[k]
                    with torch.cuda.stream(s_in):
                        b["comp"][:total].copy_(b["pin"][:total], non_blocking=True)
                        mark(f"h2d{k} issued")
                        ctx.check(lib.b4d_inflate_batch(ctx.handle, ptr(b["comp"]), offs.ctypes.data, sizes.ctypes.data, ptr(b["raw"]),
                                                        src.chunk_bytes, ptr(b["act"]), len(recs)), "inflate")
                        mark(f"inflate{k} submitted")
                        ctx.check(lib.b4d_unchunk_to_f32(ctx.handle, ptr(b["raw"]), src.code, 0, hi - lo, n, n, src.c0, src.cy, src.cx, 0,
                                                         ptr(b["frames"])), "unchunk")
                        mark(f"unchunk{k} launched")
                torch.cuda.synchronize()
                mark("gpu drained")
                src._release()
                src.close()
                mark("closed")
            if rep:
                print(f"rep {rep}: " + ", ".join(f"{name} +{(t - marks[i][1]) * 1e3:.2f}" for i, (name, t) in enumerate(marks[1:])), flush=True)


main()
