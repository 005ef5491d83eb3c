#!/usr/bin/env python
"""N-rank check of the sharded paths against the single-GPU result (tier 4). Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/multi_gpu_check.py
Every rank holds its contiguous frame range; rank 0 additionally computes the whole stack alone and compares:
  * per-frame outputs of the fused pipeline (reductions, grain, tracking, PSD / autocorrelation maps): bitwise;
  * temporal moments after the NCCL all-reduce of the power sums: <= 1e-6 relative (a different shift map and summation order);
  * the public sharded entries: parallel.analyze_stack_sharded (StackAnalyzer per rank, reference broadcast, tables gathered)
    and metrics.speckle_stack_stats under the process group (frame ranges with the one-frame halo of incremental
    tracking): every leaf, temporal["inc"] across the shard seams included, equals the single-GPU result bitwise.
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
trk = engine.PhaseTracker(ref, (n, n), y0=0, x0=0, device=local)
res = engine.stack_pipeline(mine, tracker=trk, tail_quantiles=(0.0005, 0.9995))
tables = {k: parallel.gather_rows(res[k], T) for k in ("reductions", "grain", "tracking", "quantiles")}
maps_sum = torch.stack([res["psd"].double().sum(), res["autocorr"].double().abs().sum()])
local_maps = (res["psd"].clone(), res["autocorr"].clone())

# temporal moments: shards accumulate against a common shift, power sums all-reduced over NCCL
tm = parallel.sharded_temporal_moments(mine, T)

if rank == 0:
    full = engine.as_stack(stack, local)
    trk1 = engine.PhaseTracker(torch.from_numpy(stack[0]).to(dev), (n, n), y0=0, x0=0, device=local)
    one = engine.stack_pipeline(full, tracker=trk1, tail_quantiles=(0.0005, 0.9995))
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

# ---- public sharded entries -------------------------------------------------------------------------------------
import barc4dip_b200 as dip
from barc4dip_b200.pipeline import StackAnalyzer


def same_tree(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and list(a) == list(b), f"{path}: keys differ: {list(a)} vs {list(b)}"
        for k in a:
            same_tree(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (np.ndarray, torch.Tensor)):
        x = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
        y = b.cpu().numpy() if isinstance(b, torch.Tensor) else b
        if x.dtype == object:
            assert (x == y).all(), path
        else:
            np.testing.assert_array_equal(x, y, err_msg=path)
    else:
        assert a == b or (a != a and b != b), f"{path}: {a!r} != {b!r}"


sh = parallel.analyze_stack_sharded(stack, chunk_frames=3)
# the same stack as a gzip-4 HDF5 file of uint16 frames, every rank reading its own frame range from the file
import tempfile
from barc4dip_b200.io import h5 as h5io
u16 = np.clip(np.rint(stack / stack.max() * 60000.0), 0, 60000).astype(np.uint16)
h5_path = os.path.join(tempfile.gettempdir(), f"b4d_mgc_{os.environ.get('MASTER_PORT', '0')}.h5")
if rank == 0:
    if os.path.exists(h5_path):
        os.remove(h5_path)
    h5io.save_h5(u16, h5_path)
dist.barrier()
sh_file = parallel.analyze_h5_stack_sharded(h5_path, block_frames=4, chunk_frames=3)
dist.barrier()
sp = {m: dip.metrics.speckle_stack_stats(stack, metrics="all", tiles=True, tracking_method=m, tracking_backend=b, verbose=False)
      for m, b in (("template", "opencv"), ("phase", "internal"))}
if rank == 0:
    lo_, hi_ = sh["frame_range"]
    one_a = StackAnalyzer((n, n), device=local, reference=stack[0], chunk_frames=3).run(stack, keep_maps_on_device=True)
    for k in ("psd", "autocorr"):
        assert torch.equal(sh[k], one_a[k][lo_:hi_]), k
    same_tree({k: v for k, v in sh.items() if k not in ("psd", "autocorr", "frame_range")},
              {k: v for k, v in one_a.items() if k not in ("psd", "autocorr")}, "analyze_stack_sharded")
    one_f = StackAnalyzer((n, n), device=local, reference=u16[0], chunk_frames=3, want_maps=False).run(u16)
    same_tree({k: v for k, v in sh_file.items() if k != "frame_range"}, one_f, "analyze_h5_stack_sharded")
    os.remove(h5_path)
    print(f"file-sharded entry ok: analyze_h5_stack_sharded over {world} ranks equals the single-GPU analysis of the same uint16 frames")
    for m, b in (("template", "opencv"), ("phase", "internal")):
        one_s = dip.metrics.speckle_stack_stats(stack, metrics="all", tiles=True, tracking_method=m, tracking_backend=b,
                                                verbose=False, sharded=False)
        got = dict(sp[m])
        got["meta"] = {k: v for k, v in got["meta"].items() if k != "sharding"}
        same_tree(got, one_s, f"speckle_stack_stats[{m}]")
        seams = [parallel.frame_range(T, r, world)[0] for r in range(1, world)]
        assert all(np.isfinite(got["temporal"]["inc"]["dx"][t]) for t in seams)
    print(f"sharded entries ok: analyze_stack_sharded and speckle_stack_stats (template, phase) over {world} ranks equal the "
          f"single-GPU results leaf by leaf, temporal['inc'] across the seams at frames {seams} included")
    print(f"multi_gpu_check ok: {world} ranks, {T} frames of {n}^2; per-frame outputs bitwise equal, temporal moments within 1e-6 (1e-5 absolute for skewness / kurtosis)")
dist.barrier()
dist.destroy_process_group()
