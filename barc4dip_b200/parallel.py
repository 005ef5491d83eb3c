"""
Frame-wise sharding of a stack over the GPUs of one box (one process per GPU, torch.distributed).

The stack-analysis path shards naturally: frames are independent units for every per-frame metric,
PSD, autocorrelation and tracking.  Collectives appear only where the path has a
real exchange:

  * ``broadcast_reference``  -- the tracker's reference frame (the owner of frame 0 broadcasts it; every
                                rank then builds the identical conjugate spectrum locally);
  * ``allreduce_temporal``   -- per-pixel shifted power sums of the temporal moments (sum, float64), after
                                the shift map itself was broadcast so that the sums are addable;
  * ``gather_rows``          -- (T_local, K) scalar tables back to a full (T, K) table (KBs).

The reference has no counterpart (single process, joblib threads: metrics/speckles.py:323,409,
metrics/sharpness.py:361).  Backend "nccl" on GPUs; the same code runs on "gloo" with CPU tensors, which is
how the host logic is tested without a GPU.
"""

from __future__ import annotations

import numpy as np


def dist_info():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def frame_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [r*T/G, (r+1)*T/G) of rank r; covers 0..T exactly once."""
    if n_frames < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def owner_of(frame: int, n_frames: int, world: int) -> int:
    """Rank whose shard contains `frame`."""
    if not (0 <= frame < n_frames):
        raise ValueError("frame out of range")
    for r in range(world):
        lo, hi = frame_range(n_frames, r, world)
        if lo <= frame < hi:
            return r
    raise AssertionError("unreachable")


def inc_halo_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Shard widened by the one-frame halo incremental tracking needs (frame t vs t-1, speckles.py:349,373)."""
    lo, hi = frame_range(n_frames, rank, world)
    return max(lo - 1, 0), hi


def broadcast_reference(frame, src: int = 0):
    """Broadcast the reference frame tensor in place from `src`; returns it. No-op on one rank."""
    rank, world = dist_info()
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(frame, src=src)
    return frame


def allreduce_sum_(tensor):
    rank, world = dist_info()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def allreduce_temporal(acc, total_frames: int | None = None):
    """All-reduce a TemporalAccumulator's power sums (and frame count); every rank must share acc.shift."""
    rank, world = dist_info()
    if world > 1:
        import torch
        allreduce_sum_(acc.sums)
        n = torch.tensor([acc.count], dtype=torch.int64, device=acc.sums.device)
        allreduce_sum_(n)
        acc.count = int(n.item())
    if total_frames is not None and acc.count != total_frames:
        raise RuntimeError(f"temporal accumulator holds {acc.count} frames, expected {total_frames}")
    return acc


def gather_rows(local, n_frames: int):
    """All-gather per-frame tables along axis 0 in rank order -> (n_frames, K) tensor on every rank."""
    rank, world = dist_info()
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    counts = [frame_range(n_frames, r, world)[1] - frame_range(n_frames, r, world)[0] for r in range(world)]
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def _comm_device():
    """Device the collectives of the current process group run on: the current CUDA device for NCCL, the host for gloo."""
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_tree(tree, n_frames: int):
    """All-gather every per-frame leaf of a nested result dict along axis 0 in rank order.

    Leaves are numpy arrays or tensors whose leading axis is this rank's frame range (frame_range(n_frames, rank,
    world)); anything else (scalars, strings, meta dicts, arrays with another leading length) is left as it is -- such
    leaves must already be identical on every rank. Returns the same structure with (n_frames, ...) leaves, numpy
    leaves staying numpy. No-op on one rank."""
    rank, world = dist_info()
    if world == 1:
        return tree
    import torch
    lo, hi = frame_range(n_frames, rank, world)
    dev = _comm_device()

    def walk(node):
        if isinstance(node, dict):
            return {k: walk(node[k]) for k in node}                     # (insertion order is the same on every rank)
        if isinstance(node, np.ndarray) and node.ndim >= 1 and node.shape[0] == hi - lo and node.dtype != object:
            t = torch.from_numpy(np.ascontiguousarray(node)).to(dev)
            return gather_rows(t, n_frames).cpu().numpy()
        if isinstance(node, torch.Tensor) and node.ndim >= 1 and node.shape[0] == hi - lo:
            return gather_rows(node.contiguous().to(dev), n_frames)
        return node

    return walk(tree)


def analyze_stack_sharded(stack, *, reference=None, slices_yx=None, **analyzer_kw) -> dict:
    """StackAnalyzer over the ranks of the current process group (one process per GPU, torchrun).

    Every rank passes the same (T, ny, nx) host stack (an array, a memmap, anything sliceable along axis 0 that
    returns a numpy array): rank r uploads and analyses only its contiguous frame range, the tracker's reference
    frame (default: frame 0) is broadcast from its owner so that every rank builds the identical spectrum, and the
    per-frame tables are all-gathered, so every rank returns the full (T, ...) tables. The PSD / autocorrelation maps
    stay on the device of the rank that computed them: out["psd"], out["autocorr"] hold this rank's frames
    out["frame_range"] = (lo, hi). On one rank this is StackAnalyzer.run with maps kept on the device.
    Replaces the per-frame loop of metrics/speckles.py:300-325,347-386 at stack scale."""
    import torch
    from . import engine
    from .pipeline import StackAnalyzer
    rank, world = dist_info()
    T = int(stack.shape[0])
    ny, nx = int(stack.shape[1]), int(stack.shape[2])
    lo, hi = frame_range(T, rank, world)
    dev = torch.cuda.current_device()
    analyzer = StackAnalyzer((ny, nx), device=dev, **analyzer_kw)
    if reference is None:
        owner = owner_of(0, T, world)
        ref = engine.as_stack(np.asarray(stack[0:1]), dev)[0] if rank == owner else \
            torch.empty((ny, nx), dtype=torch.float32, device=f"cuda:{dev}")
        broadcast_reference(ref, src=owner)
    else:
        ref = engine.as_stack(np.asarray(reference)[None], dev)[0]
    analyzer.set_reference(ref, slices_yx=slices_yx)
    if hi > lo:
        res = analyzer.run(np.asarray(stack[lo:hi]), keep_maps_on_device=True)
    else:
        res = analyzer.run(np.asarray(stack[0:1]), keep_maps_on_device=True)     # a rank without frames still joins the gathers
        res = {k: (_take0(v)) for k, v in res.items()}
    maps = {k: res.pop(k) for k in ("psd", "autocorr") if k in res}
    out = gather_tree(res, T)
    out.update(maps)
    out["frame_range"] = (lo, hi)
    return out


def analyze_h5_stack_sharded(path, *, reference=None, **kw) -> dict:
    """The stack stored in an HDF5 file, analysed by the ranks of the current process group straight from the file.

    Rank r reads and analyses only its contiguous frame range (io.stream.analyze_h5_stack(frames=...): compressed chunks
    -> GPU, inflated there where the device can); no rank ever holds the stack, and nothing is broadcast: every rank
    reads the tracker's reference (default: frame 0 of the file) itself. Per-frame tables are all-gathered, so every
    rank returns the full (T, ...) tables; maps (want_maps=True with keep_maps_on_device=True) stay with the rank that
    computed them; out["frame_range"] = (lo, hi). The reference loads the whole file on one host first
    (io/rw.py:129 -> metrics/speckles.py:300-325)."""
    from .io import h5 as h5io
    from .io.stream import analyze_h5_stack
    rank, world = dist_info()
    f, dset = h5io.open_dataset(path)
    try:
        if dset.ndim != 3:
            raise ValueError(f"expected a (N, H, W) stack in '{path}', got shape {dset.shape}")
        T = int(dset.shape[0])
        if reference is None:
            reference = np.asarray(dset[0])
    finally:
        f.close()
    lo, hi = frame_range(T, rank, world)
    if hi > lo:
        res = analyze_h5_stack(path, reference=reference, frames=(lo, hi), **kw)
    else:                                                              # a rank without frames still joins the gathers
        res = {k: _take0(v) for k, v in analyze_h5_stack(path, reference=reference, frames=(0, 1), **kw).items()}
    maps = {k: res.pop(k) for k in ("psd", "autocorr") if k in res}
    out = gather_tree(res, T)
    out.update(maps)
    out["frame_range"] = (lo, hi)
    return out


def _take0(v):
    if isinstance(v, dict):
        return {k: _take0(x) for k, x in v.items()}
    return v[:0]


def sharded_temporal_moments(local_stack, n_total: int, *, gain=None, dark=None, pilot_frames: int = 16):
    """Per-pixel temporal moments of a frame-sharded stack.

    Rank 0 computes the shift map from its first frames and broadcasts it (NCCL); every rank accumulates
    its shard against that common shift; the float64 sums are all-reduced; every rank finalises.
    """
    from . import engine
    rank, world = dist_info()
    T, ny, nx = local_stack.shape
    acc = engine.TemporalAccumulator(ny, nx, device=local_stack.device.index or 0, gain=gain, dark=dark)
    import torch
    if rank == 0:
        acc.pilot(local_stack, n_frames=pilot_frames)
    else:
        acc.shift = torch.empty((ny, nx), dtype=torch.float32, device=local_stack.device)
    broadcast_reference(acc.shift, src=0)
    acc.update(local_stack)
    allreduce_temporal(acc, n_total)
    return acc.finalize()


def merge_power_sums(parts: list[np.ndarray]) -> np.ndarray:
    """Host model of the all-reduce: shifted power sums with a common shift simply add."""
    return np.sum(np.stack(parts, axis=0), axis=0)
