"""
File -> GPU ingestion figures (not the headline bench): a gzip-4 chunked HDF5 stack of 2048^2 uint16 frames, written by
io.h5.save_h5, analysed by io.stream.analyze_h5_stack. Reports, in frames/s,
  decode      the block reader alone (inflate on the host threads into pinned staging),
  file_to_gpu_host    analyze_h5_stack(inflate="host"): inflate of block k+1 on the host threads overlapped with upload +
                      kernels of block k,
  file_to_gpu_device  analyze_h5_stack(inflate="device"): compressed chunks over PCIe, inflated by the GPU's
                      decompression engine, un-tiled and widened by b4d_unchunk_to_f32,
  in_memory   StackAnalyzer.run on the same frames already in pinned host memory (what the decode is measured against).

    python scripts/ingest_bench.py [--frames 48] [--n 2048] [--out gpurun_out/ingest.json]
"""

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--block", type=int, default=16)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import torch
    from barc4dip_b200 import synth
    from barc4dip_b200.io import h5 as h5io
    from barc4dip_b200.io.stream import H5StackSource, analyze_h5_stack
    from barc4dip_b200.pipeline import StackAnalyzer

    n, T = args.n, args.frames
    rng = np.random.default_rng(0)
    base = synth.speckle_frame(n, grain=6.0, seed=0)                    # mean 1000 counts, fully developed speckle
    stack = np.empty((T, n, n), np.uint16)
    for t in range(T):                                                  # shot noise on a slowly drifting pattern
        stack[t] = rng.poisson(np.roll(base, t, axis=1)).clip(0, 65535).astype(np.uint16)
    res = {"frames": T, "frame": [n, n], "dtype": "uint16", "host_threads": os.cpu_count(), "backend": h5io.backend()}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "scan.h5")
        t0 = time.perf_counter()
        h5io.save_h5(stack, path)
        res["write_s"] = time.perf_counter() - t0
        res["file_mb"], res["raw_mb"] = os.path.getsize(path) / 1e6, stack.nbytes / 1e6

        with H5StackSource(path, block_frames=args.block) as src:
            t0 = time.perf_counter()
            for _ in src:
                pass
            res["decode_frames_s"] = T / (time.perf_counter() - t0)

        an = StackAnalyzer((n, n), reference=stack[0], want_maps=False)
        from barc4dip_b200._lib import inflate_caps
        res["engine"] = dict(zip(("algo_mask", "max_bytes"), inflate_caps()))
        got = None
        for mode in ("host", "device"):
            if mode == "device" and not res["engine"]["algo_mask"] & 1:
                continue
            analyze_h5_stack(path, analyzer=an, block_frames=args.block, inflate=mode)   # warm: staging, plans, page cache
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = analyze_h5_stack(path, analyzer=an, block_frames=args.block, inflate=mode)
            torch.cuda.synchronize()
            res[f"file_to_gpu_{mode}_frames_s"] = T / (time.perf_counter() - t0)

        keep = torch.empty((stack.nbytes,), dtype=torch.uint8, pin_memory=True)
        pinned = keep.numpy().view(np.uint16).reshape(stack.shape)
        pinned[...] = stack
        an.run(pinned)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        want = an.run(pinned)
        torch.cuda.synchronize()
        res["in_memory_frames_s"] = T / (time.perf_counter() - t0)
    res["tables_identical"] = bool(np.array_equal(got["table"], want["table"])
                                   and np.array_equal(got["tracking"]["dx"], want["tracking"]["dx"]))
    line = json.dumps(res)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as fh:
            fh.write(line + "\n")


if __name__ == "__main__":
    main()
