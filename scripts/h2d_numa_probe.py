#!/usr/bin/env python
"""Host->device bandwidth of a pinned buffer under the default CPU affinity and bound to each NUMA node (diagnostic for
the end-to-end figure of bench.py: pinned pages land on the node of the allocating thread)."""
import glob, os, subprocess, time
import torch

def cpus_of(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

def h2d_gbs(nbytes=1 << 30, reps=5):
    host = torch.empty(nbytes // 4, dtype=torch.float32, pin_memory=True)
    host.fill_(1.0)
    dev = torch.empty_like(host, device="cuda:0")
    dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return nbytes * reps / (a.elapsed_time(b) / 1e3) / 1e9

torch.cuda.init()
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "numa nodes", nodes)
try:
    bus = torch.cuda.get_device_properties(0).pci_bus_id
except Exception:
    bus = None
q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
print("gpu0 bus", bus, q)
for cand in (q.lower(), q.lower()[4:] if len(q) > 12 else q.lower()):
    p = f"/sys/bus/pci/devices/{cand}/numa_node"
    if os.path.exists(p):
        print("gpu numa_node", open(p).read().strip(), "from", p)
print("default affinity: %.1f GB/s" % h2d_gbs())
full = os.sched_getaffinity(0)
for n in nodes:
    try:
        cp = set(cpus_of(n)) & full
        if not cp:
            print("node", n, "no allowed cpus"); continue
        os.sched_setaffinity(0, cp)
        print("bound to node %d (%d cpus): %.1f GB/s" % (n, len(cp), h2d_gbs()))
    except Exception as e:
        print("node", n, "failed:", e)
os.sched_setaffinity(0, full)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
