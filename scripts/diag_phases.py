"""Where do begin_stack / end_stack of bench.py spend their time? (torchrun, N ranks)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from barc4dip_b200 import engine, parallel, synth
from barc4dip_b200.pipeline import StackAnalyzer
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, F = 2048, 16
stack = torch.from_numpy(synth.speckle_stack(2, n)).to(dev).repeat(F // 2, 1, 1).contiguous()
ref = stack[0].clone()
an = StackAnalyzer((n, n), device=local, chunk_frames=F)
def T(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:34s} {(time.perf_counter() - t0) * 1e3:8.3f} ms", flush=True)
    return r
for it in range(3):
    if rank == 0: print("--- iteration", it)
    T("broadcast ref", lambda: parallel.broadcast_reference(ref, src=0))
    T("set_reference", lambda: an.set_reference(ref))
    acc = T("TemporalAccumulator()", lambda: engine.TemporalAccumulator(n, n, device=local))
    if rank == 0: T("pilot", lambda: acc.pilot(stack))
    else: acc.shift = torch.empty((n, n), dtype=torch.float32, device=dev)
    T("broadcast shift", lambda: parallel.broadcast_reference(acc.shift, src=0))
    T("update", lambda: acc.update(stack))
    T("allreduce sums", lambda: parallel.allreduce_sum_(acc.sums))
    T("allreduce_temporal (sums+count)", lambda: parallel.allreduce_temporal(acc))
    T("finalize", lambda: acc.finalize(return_device=True))
if world > 1: dist.destroy_process_group()
