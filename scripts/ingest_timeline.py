"""Host-side timeline of the device ingestion path, per block: wait for the packer, enqueue, run, byte-count check."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from barc4dip_b200 import synth
    from barc4dip_b200.io import hdf5
    from barc4dip_b200.io import stream
    from barc4dip_b200.pipeline import StackAnalyzer
    n, T = 2048, 128
    rng = np.random.default_rng(0)
    base = synth.speckle_frame(n, grain=6.0, seed=0)
    stack = np.stack([rng.poisson(np.roll(base, t, axis=1)).clip(0, 65535).astype(np.uint16) for t in range(T)])
    an = StackAnalyzer((n, n), reference=stack[0], want_maps=False)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "s.h5")
        hdf5.write_stack(path, stack)
        for blk in (32, 64):
            for rep in range(2):
                with hdf5.H5File(path) as f:
                    d = f["entry_0000/measurement/data"]
                    t_init = time.perf_counter()
                    src = stream.DeviceInflater(d, block_frames=blk)
                    it = iter(src)
                    marks = []
                    t0 = time.perf_counter()
                    while True:
                        ta = time.perf_counter()
                        try:
                            lo, frames = next(it)
                        except StopIteration:
                            break
                        tb = time.perf_counter()
                        an.run(frames)
                        tc = time.perf_counter()
                        marks.append((tb - ta, tc - tb))
                    torch.cuda.synchronize()
                    total = time.perf_counter() - t0
                    src.close()
                if rep:
                    print(f"block {blk}: {T / total:.0f} frames/s, setup {t0 - t_init:.4f} s; per block (next(), run()) ms:",
                          [(round(a * 1e3, 2), round(b * 1e3, 2)) for a, b in marks], flush=True)
        # the packer alone, as the reader thread runs it
        with hdf5.H5File(path) as f:
            src = stream.DeviceInflater(f["entry_0000/measurement/data"], block_frames=32)
            bufs = src._buffers()
            pin = bufs[0]["pin"].numpy()
            for _ in range(2):
                t0 = time.perf_counter()
                src._pack(src.blocks[1][2], pin)
                print("pack of one 32-frame block, ms:", round((time.perf_counter() - t0) * 1e3, 2))
            t0 = time.perf_counter()
            res = an.run(bufs[0]["frames"][:32])
            print("run() on 32 resident frames, ms:", round((time.perf_counter() - t0) * 1e3, 2))
            t0 = time.perf_counter()
            res = an.run(bufs[0]["frames"][:32])
            print("run() again, ms:", round((time.perf_counter() - t0) * 1e3, 2))


main()
