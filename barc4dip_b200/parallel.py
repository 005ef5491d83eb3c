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
