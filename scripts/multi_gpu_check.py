#!/usr/bin/env python
"""N-rank check of the sharded paths against the single-GPU result (tier 4). Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/multi_gpu_check.py
Every rank holds its contiguous frame range; rank 0 additionally computes the whole stack alone and compares:
  * per-frame outputs of the fused pipeline (reductions, grain, tracking, PSD / autocorrelation maps): bitwise;
  * temporal moments after the NCCL all-reduce of the power sums: <= 1e-6 relative (a different shift map and summation order).
Prints "multi_gpu_check ok" from rank 0 and exits 0, or raises."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from barc4dip_b200 import engine, parallel, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)

n, T = 512, 4 * world + 3
stack, _ = synth.tracking_stack(T, n, grain=5.0, seed=17, integer_every=3)
lo, hi = parallel.frame_range(T, rank, world)
mine = engine.as_stack(stack[lo:hi], local)

# tracker reference = frame 0, broadcast from its owner; every rank builds the same spectrum
ref = torch.from_numpy(stack[0]).to(dev) if rank == parallel.owner_of(0, T, world) else torch.empty((n, n), device=dev)
parallel.broadcast_reference(ref, src=parallel.owner_of(0, T, world))
engine.PhaseTracker(ref, (n, n), y0=0, x0=0, device=local)
res = engine.stack_pipeline(mine, tail_quantiles=(0.0005, 0.9995))
tables = {k: parallel.gather_rows(res[k], T) for k in ("reductions", "grain", "tracking", "quantiles")}
maps_sum = torch.stack([res["psd"].double().sum(), res["autocorr"].double().abs().sum()])
local_maps = (res["psd"].clone(), res["autocorr"].clone())

# temporal moments: shards accumulate against a common shift, power sums all-reduced over NCCL
tm = parallel.sharded_temporal_moments(mine, T)

if rank == 0:
    full = engine.as_stack(stack, local)
    engine.PhaseTracker(torch.from_numpy(stack[0]).to(dev), (n, n), y0=0, x0=0, device=local)
    one = engine.stack_pipeline(full, tail_quantiles=(0.0005, 0.9995))
    for k, v in tables.items():
        assert torch.equal(v, one[k]), f"{k}: sharded table differs from the single-GPU one"
    assert torch.equal(local_maps[0], one["psd"][lo:hi]) and torch.equal(local_maps[1], one["autocorr"][lo:hi])
    tm1 = engine.temporal_moments(full)
    # the ranks share rank 0's shift map, which differs from the one the single-GPU run picks (other pilot frames): the
    # fp32-over-8-frames partial sums round differently. Dimensioned maps to 1e-6 relative; the dimensionless ones
    # (skewness, excess kurtosis; O(1), zero crossings) to 1e-5 absolute.
    for k in ("mean", "std", "variance"):
        np.testing.assert_allclose(tm[k], tm1[k], rtol=1e-6, atol=1e-9, err_msg=k)
    for k in ("skewness", "kurtosis"):
        np.testing.assert_allclose(tm[k], tm1[k], rtol=1e-6, atol=1e-5, err_msg=k)
    print(f"multi_gpu_check ok: {world} ranks, {T} frames of {n}^2; per-frame outputs bitwise equal, temporal moments within 1e-6 (1e-5 absolute for skewness / kurtosis)")
dist.barrier()
dist.destroy_process_group()
