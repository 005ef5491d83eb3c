"""
File -> GPU ingestion of a detector stack: the caller side of the stack pipeline.

The reference loads the whole stack into host memory (`read_image` -> `read_h5` -> `dset[()]`, io/rw.py:129,
io/h5.py:80) and only then starts its per-frame loop (metrics/speckles.py:300-325). Here the file is walked in blocks of
frames: a reader thread inflates the chunks of block k+1 (thread pool, GIL released) straight into one of two pinned
staging buffers while `StackAnalyzer.run` uploads and analyses block k; integer detector frames stay in their own width
until they are on the device (`b4d_cast_to_f32`). Host memory in use: two blocks, whatever the stack's length.

Where the GPU has a hardware decompression engine (B200: `b4d_inflate_caps`) the host does not inflate at all
(`DeviceInflater`): the deflate streams of a block are copied as stored into pinned staging, cross PCIe compressed, are
inflated by the engine (`b4d_inflate_batch`, one batch per block) and one kernel (`b4d_unchunk_to_f32`) undoes the byte
shuffle and the chunk tiling on the way to float32 frames. `inflate="auto"` takes that path whenever the file qualifies
(builtin codec, (shuffle +) deflate pipeline, native element type, chunks within the engine's limit).

    res = analyze_h5_stack("scan_0001.h5")                       # reference frame = frame 0 of the file
    res = analyze_h5_stack(path, frames=parallel.frame_range(T, rank, world))   # one rank's share of a stack
"""

from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .._lib import UNCHUNK_CODES, get_context, inflate_caps, native_int_code, ptr, require_cuda
from . import h5 as _h5
from . import hdf5 as _hdf5


class H5StackSource:
    """Frames [a, b) of the stack in `path`, delivered block by block into pinned host buffers.

    Iterating yields (first frame index, array view of the block); the view is valid until the next iteration step.
    Blocks are decoded one ahead of the consumer on a background thread."""

    def __init__(self, path, *, frames: tuple[int, int] | None = None, block_frames: int = 32,
                 decode_threads: int | None = None, pinned: bool = True):
        self._file, self.dset = _h5.open_dataset(path)
        try:
            if self.dset.ndim != 3:
                raise ValueError(f"expected a (N, H, W) stack at '{_h5.DATASET_PATH}', got shape {self.dset.shape} in '{path}'")
            T = int(self.dset.shape[0])
            a, b = (0, T) if frames is None else (int(frames[0]), int(frames[1]))
            if not (0 <= a <= b <= T):
                raise ValueError(f"frames {frames} outside a stack of {T} frames")
        except BaseException:
            self._file.close()
            raise
        self.range = (a, b)
        self.frame_shape = tuple(int(s) for s in self.dset.shape[1:])
        src = np.dtype(self.dset.dtype)
        # integer detector types and float32 are staged as stored; everything else is converted to float32 on the host
        self.dtype = src.newbyteorder("=") if (native_int_code(src.newbyteorder("=")) is not None or src == np.float32) \
            else np.dtype(np.float32)
        self.block = max(1, min(int(block_frames), max(1, b - a)))
        self.threads = decode_threads
        self._pinned = bool(pinned)
        self._bufs = None

    def __len__(self):
        return self.range[1] - self.range[0]

    def close(self):
        if self._file is not None:
            self._file.close()
            self._file = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _staging(self):
        if self._bufs is None:
            shape = (self.block,) + self.frame_shape
            if self._pinned:
                torch = require_cuda()
                nbytes = int(np.prod(shape)) * self.dtype.itemsize
                self._keep = [torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
                self._bufs = [t.numpy().view(self.dtype).reshape(shape) for t in self._keep]
            else:
                self._bufs = [np.empty(shape, self.dtype) for _ in range(2)]
        return self._bufs

    def read_block(self, a: int, b: int, out: np.ndarray) -> np.ndarray:
        """Frames [a, b) of the file into out[:b-a] (staging dtype)."""
        view = out[:b - a]
        if hasattr(self.dset, "read") and self.dtype == np.dtype(self.dset.dtype).newbyteorder("="):
            self.dset.read(a, b, out=view, threads=self.threads)      # builtin codec: chunks land in the buffer
        elif hasattr(self.dset, "read_direct") and self.dtype == self.dset.dtype:
            self.dset.read_direct(view, np.s_[a:b])                   # h5py
        else:
            view[...] = self.dset[a:b]
        return view

    def __iter__(self):
        bufs = self._staging()
        a0, b0 = self.range
        starts = list(range(a0, b0, self.block))
        free: queue.Queue = queue.Queue()
        ready: queue.Queue = queue.Queue()
        for i in range(2):
            free.put(i)
        stop = threading.Event()

        def producer():
            try:
                for a in starts:
                    i = free.get()
                    if stop.is_set():
                        return
                    b = min(b0, a + self.block)
                    ready.put((a, i, self.read_block(a, b, bufs[i])))
                ready.put(None)
            except BaseException as e:                                 # surfaces in the consumer
                ready.put(e)

        th = threading.Thread(target=producer, name="b4d-h5-reader", daemon=True)
        th.start()
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                a, i, view = item
                yield a, view
                free.put(i)
        finally:
            stop.set()
            free.put(0)                                                # wake a producer waiting for a buffer
            th.join()


_CACHE: dict = {}            # device index -> {"busy", "caps", "flat"}: the staging buffers kept between calls
_CACHE_LOCK = threading.Lock()


def release_buffers():
    """Drop the cached staging buffers of the device ingestion path (pinned host + device memory)."""
    with _CACHE_LOCK:
        for key in [k for k, e in _CACHE.items() if not e["busy"]]:
            del _CACHE[key]


class DeviceInflater:
    """Frames [a, b) of a deflate-chunked dataset -> float32 CUDA frames, inflated by the GPU's decompression engine.

    Per block of frames: the chunks' raw deflate streams (zlib header and Adler-32 trailer checked / dropped on the host)
    are packed into a pinned buffer, uploaded, inflated in one batch and un-tiled + widened by one kernel. Two blocks
    are in flight: a reader thread packs block k+1 while the GPU works on block k."""

    def __init__(self, dset, *, frames: tuple[int, int] | None = None, block_frames: int = 32, device: int | None = None,
                 pack_threads: int | None = None, ramp: bool = True):
        torch = require_cuda()
        why = self.unsupported(dset, device)
        if why:
            raise _hdf5.H5Unsupported(why)
        self.dset, self.ctx = dset, get_context(device)
        self.device = torch.device(f"cuda:{self.ctx.device}")
        T, self.ny, self.nx = dset.shape
        a, b = (0, T) if frames is None else (int(frames[0]), int(frames[1]))
        if not (0 <= a <= b <= T):
            raise ValueError(f"frames {frames} outside a stack of {T} frames")
        self.range, self.frame_shape = (a, b), (self.ny, self.nx)
        self.c0, self.cy, self.cx = dset.chunks
        self.gy, self.gx = -(-self.ny // self.cy), -(-self.nx // self.cx)
        self.code = UNCHUNK_CODES[str(dset.dtype.newbyteorder("="))]
        self.shuffled = dset.filters[0][0] == _hdf5.FILTER_SHUFFLE
        self.chunk_bytes = self.c0 * self.cy * self.cx * dset.dtype.itemsize
        blk = max(1, min(int(block_frames), max(1, b - a)))
        if blk > self.c0:
            blk -= blk % self.c0                                        # whole chunk rows per block where possible
        self.block = blk
        index = dset._chunk_index()                                     # sorted by (frame, y, x) offsets
        per_fb = self.gy * self.gx
        # the first blocks are short (a quarter, a half of a block, in whole chunk rows): the consumer gets its first
        # frames after a quarter of a block's pack + upload + inflate time instead of a whole one
        self.blocks = []
        lo, size = a, max(self.c0, (blk // 4) - (blk // 4) % self.c0) if ramp else blk
        while lo < b:
            hi = min(b, lo + size)
            fb0, fb1 = lo // self.c0, (hi - 1) // self.c0
            recs = index[fb0 * per_fb:(fb1 + 1) * per_fb]
            self.blocks.append((lo, hi, recs))
            lo, size = hi, min(blk, 2 * size)
        self._cap_chunks = max((len(r) for _, _, r in self.blocks), default=0)
        self._cap_bytes = max((sum(16 + n + (-n % 16) for _, _, n, _ in r) for _, _, r in self.blocks), default=0)
        self._pack_threads = max(1, min(12, os.cpu_count() or 1)) if pack_threads is None else max(1, int(pack_threads))
        self._pool = ThreadPoolExecutor(self._pack_threads) if self._pack_threads > 1 else None

    @staticmethod
    def unsupported(dset, device: int | None = None) -> str | None:
        """Why this dataset cannot take the device path (None: it can)."""
        if not isinstance(dset, _hdf5.H5Dataset):
            return "the device path reads chunks through the builtin codec"
        if dset.ndim != 3 or dset.chunks is None:
            return "not a chunked (N, H, W) dataset"
        ids = [f[0] for f in dset.filters]
        if ids not in ([_hdf5.FILTER_DEFLATE], [_hdf5.FILTER_SHUFFLE, _hdf5.FILTER_DEFLATE]):
            return f"filter pipeline {ids} is not (shuffle +) deflate"
        if not dset.dtype.isnative or str(dset.dtype) not in UNCHUNK_CODES:
            return f"element type {dset.dtype} has no device cast"
        if ids[0] == _hdf5.FILTER_SHUFFLE and dset.filters[0][1] and int(dset.filters[0][1][0]) != dset.dtype.itemsize:
            return "shuffle filter over a different element size"
        mask, max_bytes = inflate_caps(device)
        if not mask & 1:
            return "no deflate decompression engine on this device / driver"
        if int(np.prod(dset.chunks)) * dset.dtype.itemsize > max_bytes:
            return f"chunks exceed the engine's limit of {max_bytes} bytes"
        index = dset._chunk_index()
        grid = int(np.prod([-(-s // c) for s, c in zip(dset.shape, dset.chunks)]))
        if len(index) != grid:
            return "dataset with unallocated chunks"
        if any(r[3] for r in index):
            return "chunks stored with filters masked out"
        if any(r[2] < 7 for r in index):
            return "chunk too short to be a zlib stream"
        return None

    def _buffers(self):
        """Two sets of staging buffers. Pinning a few hundred MB costs tens of milliseconds, more than a short stack takes
        to analyse, so one pair per device is kept between calls (release_buffers() drops it)."""
        torch = require_cuda()
        fb_max = max(((hi - 1) // self.c0 - lo // self.c0 + 1) for lo, hi, _ in self.blocks)
        need = {"bytes": max(16, self._cap_bytes), "raw": fb_max * self.gy * self.gx * self.chunk_bytes,
                "act": max(1, self._cap_chunks), "frames": self.block * self.ny * self.nx}
        dev, key = self.device, self.ctx.device
        with _CACHE_LOCK:
            ent = _CACHE.get(key)
            if ent is not None and not ent["busy"] and all(ent["caps"][k] >= v for k, v in need.items()):
                ent["busy"] = True
            else:
                ent = None
        if ent is None:
            flat = [{"pin": torch.empty((need["bytes"],), dtype=torch.uint8, pin_memory=True),
                     "comp": torch.empty((need["bytes"],), dtype=torch.uint8, device=dev),
                     "raw": torch.empty((need["raw"],), dtype=torch.uint8, device=dev),
                     "act": torch.zeros((need["act"],), dtype=torch.int32, device=dev),
                     "frames": torch.empty((need["frames"],), dtype=torch.float32, device=dev)} for _ in range(2)]
            ent = {"busy": True, "caps": need, "flat": flat}
            with _CACHE_LOCK:
                old = _CACHE.get(key)
                if old is None or not old["busy"]:
                    _CACHE[key] = ent                                   # (a pair in use elsewhere keeps its place)
        else:
            torch.cuda.synchronize(dev)                                # whatever last used the pair has drained
        self._cache_entry = ent
        n = self.block * self.ny * self.nx
        return [dict(f, frames=f["frames"][:n].view(self.block, self.ny, self.nx)) for f in ent["flat"]]

    def _release(self):
        ent, self._cache_entry = getattr(self, "_cache_entry", None), None
        if ent is not None:
            with _CACHE_LOCK:
                ent["busy"] = False

    def _pack(self, recs, pin: np.ndarray):
        """Stored chunks of `recs` into the pinned buffer, read with pread on a few threads (no page faults of a mapping,
        GIL released); each chunk lands so that its raw deflate stream -- the zlib stream minus the 2-byte header, the
        Adler-32 trailer is simply not handed to the engine -- starts on a 16-byte boundary -> (offsets, sizes, bytes)."""
        n_rec = len(recs)
        offs, sizes = np.empty(n_rec, np.int64), np.empty(n_rec, np.int64)
        pos = 0
        for i, (_, _, nbytes, _) in enumerate(recs):
            offs[i], sizes[i] = pos + 16, nbytes - 6
            pos += 16 + nbytes + (-nbytes % 16)
        fd = self.dset._f._fh.fileno()

        def load(span):
            for i in range(*span):
                _, addr, nbytes, _ = recs[i]
                at = int(offs[i]) - 2
                got = os.preadv(fd, [pin[at:at + nbytes]], addr)
                cmf, flg = int(pin[at]), int(pin[at + 1])
                if got != nbytes or (cmf & 0x0F) != 8 or ((cmf << 8) | flg) % 31 or (flg & 0x20):
                    raise OSError(f"chunk of '{self.dset.name}' at {addr} is not a zlib stream")

        nt = max(1, min(self._pack_threads, n_rec))
        spans = [(n_rec * t // nt, n_rec * (t + 1) // nt) for t in range(nt)]
        if nt == 1:
            load(spans[0])
        else:
            list(self._pool.map(load, spans))
        return offs, sizes, pos

    def __iter__(self):
        """Yields (first frame index, float32 CUDA tensor (n, ny, nx)) per block; the tensor is valid, and ordered after
        its production on the current stream, until the next iteration step."""
        torch = require_cuda()
        if not self.blocks:
            return
        bufs = self._buffers()
        pins = [b["pin"].numpy() for b in bufs]
        s_up, s_in = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)   # upload; inflate + unchunk
        free: queue.Queue = queue.Queue()
        ready: queue.Queue = queue.Queue()
        for i in range(2):
            free.put((i, None))
        stop = threading.Event()

        def producer():
            try:
                for k, (_, _, recs) in enumerate(self.blocks):
                    i, ev = free.get()
                    if stop.is_set():
                        return
                    if ev is not None:
                        ev.synchronize()                               # the upload that read this pinned buffer is done
                    ready.put((k, i) + self._pack(recs, pins[i]))
                ready.put(None)
            except BaseException as e:
                ready.put(e)

        th = threading.Thread(target=producer, name="b4d-h5-packer", daemon=True)
        th.start()
        ctx, lib = self.ctx, self.ctx.lib
        done = [torch.cuda.Event() for _ in range(2)]

        def enqueue():
            item = ready.get()
            if item is None:
                return None
            if isinstance(item, BaseException):
                raise item
            k, i, offs, sizes, total = item
            lo, hi, recs = self.blocks[k]
            b = bufs[i]
            cur_stream = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(s_up):
                s_up.wait_stream(cur_stream)                               # set i's previous consumer is behind us
                b["comp"][:total].copy_(b["pin"][:total], non_blocking=True)
                up = torch.cuda.Event()
                up.record(s_up)
                free.put((i, up))
            # the decompression engine and the copy engine are different units: on streams of their own the upload of
            # block k+1 runs under the inflation of block k (one stream for both: 0.22 ms per frame, the sum of the two)
            with torch.cuda.stream(s_in):
                s_in.wait_event(up)
                s_in.wait_stream(cur_stream)
                ctx.check(lib.b4d_inflate_batch(ctx.handle, ptr(b["comp"]), offs.ctypes.data, sizes.ctypes.data, ptr(b["raw"]),
                                                self.chunk_bytes, ptr(b["act"]), len(recs)), "b4d_inflate_batch")
                ctx.check(lib.b4d_unchunk_to_f32(ctx.handle, ptr(b["raw"]), self.code, int(self.shuffled), hi - lo, self.ny,
                                                 self.nx, self.c0, self.cy, self.cx, lo - (lo // self.c0) * self.c0,
                                                 ptr(b["frames"])), "b4d_unchunk_to_f32")
                done[i].record(s_in)
            return k, i, lo, hi, len(recs)

        try:
            nxt = enqueue()
            while nxt is not None:
                cur = nxt
                k, i, lo, hi, n = cur
                # block k+1 goes to the copy / decompression engines before block k is handed out: its upload and
                # inflation overlap the consumer's kernels on block k (the default stream there has nothing pending
                # from block k-1: consumers synchronise before they return, StackAnalyzer.run does)
                nxt = enqueue()
                torch.cuda.current_stream(self.device).wait_event(done[i])
                yield lo, bufs[i]["frames"][:hi - lo]
                if not bool((bufs[i]["act"][:n] == self.chunk_bytes).all()):
                    raise OSError(f"chunks of '{self.dset.name}' in frames [{lo}, {hi}) did not inflate to {self.chunk_bytes} bytes")
        finally:
            stop.set()
            free.put((0, None))
            th.join()
            torch.cuda.current_stream(self.device).wait_stream(s_up)
            torch.cuda.current_stream(self.device).wait_stream(s_in)
            self._release()

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def concat_results(parts: list[dict]) -> dict:
    """Per-block result dicts of StackAnalyzer.run -> one dict: every array leaf concatenated along the frame axis."""
    if len(parts) == 1:
        return parts[0]
    first = parts[0]
    out = {}
    for k, v in first.items():
        vals = [p[k] for p in parts]
        if isinstance(v, dict):
            out[k] = concat_results(vals)
        elif isinstance(v, np.ndarray):
            out[k] = np.concatenate(vals, axis=0)
        else:                                                          # device tensors (keep_maps_on_device=True)
            torch = require_cuda()
            out[k] = torch.cat(vals, dim=0)
    return out


def analyze_h5_stack(path, *, reference=None, frames: tuple[int, int] | None = None, block_frames: int = 32,
                     decode_threads: int | None = None, keep_maps_on_device: bool = False, analyzer=None,
                     inflate: str = "auto", **analyzer_kw) -> dict:
    """Analyse the stack stored in an HDF5 file with the fused pipeline; returns what `StackAnalyzer.run` returns for the
    same frames held in memory.

    reference: the tracker's reference frame (default: frame 0 of the FILE, also for a rank that reads a later range);
    frames: the frame range to analyse; analyzer: an existing StackAnalyzer to reuse (then `analyzer_kw` / `reference`
    are ignored); analyzer_kw go to StackAnalyzer (want_maps defaults to False here: the maps of a long stack do not
    belong in host memory); inflate: "device" = the GPU's decompression engine (raises if the file or the device does
    not qualify), "host" = zlib on the host threads, "auto" = the device when it qualifies."""
    from ..pipeline import StackAnalyzer

    if inflate not in ("auto", "device", "host"):
        raise ValueError(f"inflate must be 'auto', 'device' or 'host', got {inflate!r}")

    def make_analyzer(frame_shape, first_frame):
        analyzer_kw.setdefault("want_maps", False)
        ref = np.asarray(first_frame()) if reference is None else reference
        return StackAnalyzer(frame_shape, reference=ref, **analyzer_kw)

    if inflate != "host":
        f = None
        try:
            f = _hdf5.H5File(path)
            dset = f[_h5.DATASET_PATH] if _h5.DATASET_PATH in f else None
            dev = analyzer.dev if analyzer is not None else analyzer_kw.get("device")
            why = "dataset not found" if dset is None else DeviceInflater.unsupported(dset, dev)
            if why is None:
                src = DeviceInflater(dset, frames=frames, block_frames=block_frames, device=dev)
                if analyzer is None:
                    analyzer = make_analyzer(src.frame_shape, lambda: dset[0])
                parts = [analyzer.run(block, keep_maps_on_device=keep_maps_on_device) for _, block in src]
                if not parts:
                    raise ValueError("no frames to analyse")
                return concat_results(parts)
            if inflate == "device":
                raise _hdf5.H5Unsupported(f"'{path}' cannot be inflated on the device: {why}")
        except _hdf5.H5Unsupported:
            if inflate == "device":
                raise
        finally:
            if f is not None:
                f.close()

    with H5StackSource(path, frames=frames, block_frames=block_frames, decode_threads=decode_threads) as src:
        if analyzer is None:
            analyzer = make_analyzer(src.frame_shape, lambda: src.dset[0])
        parts = [analyzer.run(block, keep_maps_on_device=keep_maps_on_device) for _, block in src]
    if not parts:
        raise ValueError("no frames to analyse")
    return concat_results(parts)
