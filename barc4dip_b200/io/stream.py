"""
File -> GPU ingestion of a detector stack: the caller side of the stack pipeline.

The reference loads the whole stack into host memory (`read_image` -> `read_h5` -> `dset[()]`, io/rw.py:129,
io/h5.py:80) and only then starts its per-frame loop (metrics/speckles.py:300-325). Here the file is walked in blocks of
frames: a reader thread inflates the chunks of block k+1 (thread pool, GIL released) straight into one of two pinned
staging buffers while `StackAnalyzer.run` uploads and analyses block k; integer detector frames stay in their own width
until they are on the device (`b4d_cast_to_f32`). Host memory in use: two blocks, whatever the stack's length.

    res = analyze_h5_stack("scan_0001.h5")                       # reference frame = frame 0 of the file
    res = analyze_h5_stack(path, frames=parallel.frame_range(T, rank, world))   # one rank's share of a stack
"""

from __future__ import annotations

import queue
import threading

import numpy as np

from .._lib import native_int_code, require_cuda
from . import h5 as _h5


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
                     **analyzer_kw) -> dict:
    """Analyse the stack stored in an HDF5 file with the fused pipeline; returns what `StackAnalyzer.run` returns for the
    same frames held in memory.

    reference: the tracker's reference frame (default: frame 0 of the FILE, also for a rank that reads a later range);
    frames: the frame range to analyse; analyzer: an existing StackAnalyzer to reuse (then `analyzer_kw` / `reference`
    are ignored); analyzer_kw go to StackAnalyzer (want_maps defaults to False here: the maps of a long stack do not
    belong in host memory)."""
    from ..pipeline import StackAnalyzer

    with H5StackSource(path, frames=frames, block_frames=block_frames, decode_threads=decode_threads) as src:
        if analyzer is None:
            analyzer_kw.setdefault("want_maps", False)
            if reference is None:
                reference = np.asarray(src.dset[0])
            analyzer = StackAnalyzer(src.frame_shape, reference=reference, **analyzer_kw)
        parts = [analyzer.run(block, keep_maps_on_device=keep_maps_on_device) for _, block in src]
    if not parts:
        raise ValueError("no frames to analyse")
    return concat_results(parts)
