"""Where the device ingestion path spends its time: host packing, upload, decompression engine, unchunk kernel."""
import json
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
    from barc4dip_b200.io import hdf5
    from barc4dip_b200.io.stream import DeviceInflater
    n, T, blk = 2048, 32, 16
    rng = np.random.default_rng(0)
    base = synth.speckle_frame(n, grain=6.0, seed=0)
    stack = np.stack([rng.poisson(np.roll(base, t, axis=1)).clip(0, 65535).astype(np.uint16) for t in range(T)])
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for label, chunks, shuffle in (("1MiB", (1, 256, 2048), False), ("4MiB", (1, 1024, 2048), False), ("1MiB_shuffle", (1, 256, 2048), True),
                                       ("256KiB", (1, 64, 2048), False)):
            path = os.path.join(tmp, f"{label}.h5")
            hdf5.write_stack(path, stack, chunks=chunks, shuffle=shuffle)
            r = {"file_mb": os.path.getsize(path) / 1e6}
            with hdf5.H5File(path) as f:
                d = f["entry_0000/measurement/data"]
                inf = DeviceInflater(d, block_frames=blk)
                bufs = inf._buffers()
                b = bufs[0]
                pin = b["pin"].numpy()
                lo, hi, recs = inf.blocks[0]
                inf._pack(recs, pin)
                t0 = time.perf_counter()
                for _ in range(3):
                    offs, sizes, total = inf._pack(recs, pin)
                r["pack_ms_per_frame"] = (time.perf_counter() - t0) / 3 / (hi - lo) * 1e3
                ctx, lib = inf.ctx, inf.ctx.lib
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

                def step():
                    ev[0].record()
                    b["comp"][:total].copy_(b["pin"][:total], non_blocking=True)
                    ev[1].record()
                    ctx.check(lib.b4d_inflate_batch(ctx.handle, ptr(b["comp"]), offs.ctypes.data, sizes.ctypes.data, ptr(b["raw"]),
                                                    inf.chunk_bytes, ptr(b["act"]), len(recs)), "inflate")
                    ev[2].record()
                    ctx.check(lib.b4d_unchunk_to_f32(ctx.handle, ptr(b["raw"]), inf.code, int(inf.shuffled), hi - lo, n, n,
                                                     inf.c0, inf.cy, inf.cx, 0, ptr(b["frames"])), "unchunk")
                    ev[3].record()
                    torch.cuda.synchronize()
                    return [ev[i].elapsed_time(ev[i + 1]) / (hi - lo) for i in range(3)]
                step()
                t = np.mean([step() for _ in range(3)], axis=0)
                r["h2d_ms_per_frame"], r["inflate_ms_per_frame"], r["unchunk_ms_per_frame"] = (float(v) for v in t)
                r["inflate_out_gb_s"] = n * n * 2 / r["inflate_ms_per_frame"] / 1e6
                r["ok"] = bool(np.array_equal(b["frames"][:hi - lo].cpu().numpy(), stack[lo:hi].astype(np.float32)))
            res[label] = r
            print(label, r, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/ingest_diag.json", "w").write(json.dumps(res) + "\n")


main()
