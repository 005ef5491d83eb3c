import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from barc4dip_b200 import engine, synth
from barc4dip_b200.pipeline import StackAnalyzer
dev = torch.device('cuda:0')
n, F = 2048, 32
base = synth.speckle_frame(n, grain=6.0, seed=0)
shifts = bench.make_shifts(F, 2)
stack = bench.device_stack(base, shifts, 1234, dev)
u16 = stack.clamp(0, 65535).round()
an = StackAnalyzer((n, n), device=0, chunk_frames=8, want_maps=True, want_contrast=True)
an.set_reference(u16[0])
res = an.run_device(u16, resolve_tails=False)
print('tails unresolved', int((res['n_valid'] < 0).sum()), 'snr nan', int(torch.isnan(res['tracking'][:, 3]).sum()))
res = an.run_device(stack, resolve_tails=False)
print('f32: tails unresolved', int((res['n_valid'] < 0).sum()), 'snr nan', int(torch.isnan(res['tracking'][:, 3]).sum()))
host = torch.empty((F, n, n), dtype=torch.uint16, pin_memory=True); host.copy_(u16.to(torch.uint16))
hostf = torch.empty((F, n, n), dtype=torch.float32, pin_memory=True); hostf.copy_(stack)
for name, h in (('u16', host), ('f32', hostf)):
    for _ in range(2): an.run(h, keep_maps_on_device=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    an.run(h, keep_maps_on_device=True); torch.cuda.synchronize()
    print(name, 'run', (time.perf_counter() - t0) * 1e3, 'ms for', F)
