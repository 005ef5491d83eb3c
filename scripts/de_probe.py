"""
Probe of the hardware decompression engine (cuMemBatchDecompressAsync, CUDA 12.8+): is it exposed on this GPU, which
allocations qualify, does it take zlib-wrapped or raw deflate streams, how fast is it on detector-like data.

    python scripts/de_probe.py [--out gpurun_out/de_probe.json]
"""

import argparse
import ctypes
import json
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Params(ctypes.Structure):
    _fields_ = [("srcNumBytes", ctypes.c_size_t), ("dstNumBytes", ctypes.c_size_t), ("dstActBytes", ctypes.c_void_p),
                ("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("algo", ctypes.c_int), ("padding", ctypes.c_ubyte * 20)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--unsafe-wrapped", action="store_true",
                    help="also hand the engine a zlib-WRAPPED stream: on B200 this ends in a sticky 'unspecified launch "
                         "failure' (the engine does not validate its input) and the CUDA context is lost -- run it last, alone")
    args = ap.parse_args()
    import torch
    res = {"sizeof_params": ctypes.sizeof(Params)}
    cu = ctypes.CDLL("libcuda.so.1")
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    dev = ctypes.c_int()
    cu.cuCtxGetDevice(ctypes.byref(dev))
    v = ctypes.c_int()
    for name, attr in (("algo_mask", 136), ("max_length", 137)):
        rc = cu.cuDeviceGetAttribute(ctypes.byref(v), attr, dev)
        res[name] = [rc, v.value]
    print(res, flush=True)

    def capable(ptr):
        b = ctypes.c_int(0)
        rc = cu.cuPointerGetAttribute(ctypes.byref(b), 21, ctypes.c_uint64(ptr))
        return [rc, b.value]

    t = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    res["torch_alloc_capable"] = capable(t.data_ptr())
    raw_ptr = ctypes.c_uint64()
    cu.cuMemAlloc_v2(ctypes.byref(raw_ptr), ctypes.c_size_t(1 << 20))
    res["cuMemAlloc_capable"] = capable(raw_ptr.value)
    pin = torch.empty(1 << 20, dtype=torch.uint8, pin_memory=True)
    res["pinned_host_capable"] = capable(pin.data_ptr())
    print(res, flush=True)

    try:
        fn = cu.cuMemBatchDecompressAsync
    except AttributeError:
        res["entry_point"] = "missing"
        fn = None
    if fn is not None and res["algo_mask"][1] & 1:
        fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
        rng = np.random.default_rng(0)
        from barc4dip_b200 import synth
        base = synth.speckle_frame(2048, grain=6.0, seed=0)
        frame = rng.poisson(base).clip(0, 65535).astype(np.uint16)
        rows = 256
        chunks = [frame[r:r + rows].tobytes() for r in range(0, 2048, rows)]          # 8 chunks of 1 MiB

        def run(streams, n_rep=1, label=""):
            n = len(streams)
            src_sizes = [len(s) for s in streams]
            blob = b"".join(s + b"\x00" * (-len(s) % 16) for s in streams)
            offs = np.cumsum([0] + [len(s) + (-len(s) % 16) for s in streams])[:-1]
            d_src = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
            d_dst = torch.zeros(n * (1 << 20), dtype=torch.uint8, device="cuda")
            d_act = torch.zeros(n, dtype=torch.int32, device="cuda")
            arr = (Params * n)()
            for i in range(n):
                arr[i].srcNumBytes = src_sizes[i]
                arr[i].dstNumBytes = 1 << 20
                arr[i].dstActBytes = d_act.data_ptr() + 4 * i
                arr[i].src = d_src.data_ptr() + int(offs[i])
                arr[i].dst = d_dst.data_ptr() + i * (1 << 20)
                arr[i].algo = 1
            err = ctypes.c_size_t(0)
            stream = torch.cuda.current_stream().cuda_stream
            torch.cuda.synchronize()
            rc = fn(arr, n, 0, ctypes.byref(err), ctypes.c_void_p(stream))
            torch.cuda.synchronize()
            out = {"rc": rc, "err_index": err.value if rc else None, "act": d_act[:4].tolist()}
            if rc == 0:
                out["src_capable"] = capable(d_src.data_ptr())
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(n_rep):
                    fn(arr, n, 0, ctypes.byref(err), ctypes.c_void_p(stream))
                ev1.record()
                torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / n_rep
                out["ms"] = ms
                out["out_gb_s"] = n * (1 << 20) / ms / 1e6
                out["in_gb_s"] = sum(src_sizes) / ms / 1e6
            return out, d_dst

        def probe(label, wbits, strip):
            streams = []
            for c in chunks:
                co = zlib.compressobj(4, zlib.DEFLATED, wbits)
                s = co.compress(c) + co.flush()
                streams.append(s[2:-4] if strip else s)
            try:
                out, d_dst = run(streams, label=label)
                got = d_dst.cpu().numpy().tobytes()
                out["correct"] = got == b"".join(chunks)
                out["ratio"] = sum(len(s) for s in streams) / (len(chunks) << 20)
            except Exception as e:                                    # noqa: BLE001
                out = {"exception": repr(e)}
            res[label] = out
            print(label, out, flush=True)

        probe("raw_deflate", -15, False)
        probe("zlib_stripped", 15, True)
        # throughput: 64 frames' worth of chunks in one batch
        if res.get("zlib_stripped", {}).get("correct") or res.get("raw_deflate", {}).get("correct"):
            co_streams = []
            for c in chunks:
                co = zlib.compressobj(4, zlib.DEFLATED, -15)
                co_streams.append(co.compress(c) + co.flush())
            big = co_streams * 64
            out, _ = run(big, n_rep=5)
            out.pop("act", None)
            res["batch_512_chunks"] = out
            print("batch", out, flush=True)
            t0 = time.perf_counter()
            for s in co_streams:
                zlib.decompress(s, -15)
            res["host_inflate_one_thread_mb_s"] = 8 * (1 << 20) / (time.perf_counter() - t0) / 1e6
        if args.unsafe_wrapped:
            probe("zlib_wrapped", 15, False)
    line = json.dumps(res)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        open(args.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
