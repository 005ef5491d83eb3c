#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
usage: python scripts/make_profiles.py <tag> <launches.csv> <full.ncu-rep> "<command line that was profiled>"
writes profiles/<tag>_launches.txt (per-kernel totals of the launch list), profiles/<tag>_kernels.txt (one block per
profiled launch: time, DRAM bytes, throughputs, stalls) and profiles/ncu_traffic.json (DRAM bytes per launch and per
frame of each kernel, read by bench.py for roofline.traffic)."""
import collections, csv, io, json, os, re, subprocess, sys
tag, launches, rep, cmd = sys.argv[1:5]
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 128
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
short = lambda n: re.sub(r"\(.*", "", n).replace("<unnamed>::", "").replace("void ", "")

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr, body = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in body:
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ix["Metric Unit"]], 1)
    k = short(r[ix["Kernel Name"]])
    agg[k][0] += 1; agg[k][1] += us
tot = sum(v[1] for v in agg.values())
with open(os.path.join(root, "profiles", f"{tag}_launches.txt"), "w") as fh:
    fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {cmd}\n")
    fh.write("# our kernels only (the synthetic stack is generated with torch before the timed region); cold-cache, serialised: compare shares.\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{v[1]:12.1f} us {v[0]:5d}x {100 * v[1] / tot:6.2f}%  {k}\n")
    fh.write(f"total {tot:.1f} us over {len(body)} launches\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return float("nan")
def scaled(r, k, table):
    return g(r, k) * table.get(units[ix[k]], 1.0)
BY = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TM = {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
traffic = {}
with open(os.path.join(root, "profiles", f"{tag}_kernels.txt"), "w") as fh:
    fh.write(f"# ncu --set full --clock-control none, {cmd}\n# one block per profiled launch ({frames} frames per launch unless the grid says otherwise)\n")
    for r in data:
        name = short(r[ix["Kernel Name"]])
        t_us = scaled(r, "gpu__time_duration.sum", TM)
        rd, wr = scaled(r, "dram__bytes_read.sum", BY), scaled(r, "dram__bytes_write.sum", BY)
        st = {h.split("issue_stalled_")[1].split("_per_")[0]: g(r, h) for h in hdr
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
        top = ", ".join(f"{k}:{v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
        fh.write(f"{name}  grid {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}  regs {r[ix['launch__registers_per_thread']]}\n")
        fh.write(f"   time {t_us:.1f} us   dram read {rd / 1e6:.1f} MB  write {wr / 1e6:.1f} MB  -> {(rd + wr) / t_us / 1e3:.0f} GB/s\n")
        fh.write(f"   dram {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f}%  L2 {g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f}%"
                 f"  L1/smem pipe {g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f}%  issue {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}%"
                 f"  fma pipe {g(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.0f}%  warps {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}%\n")
        fh.write(f"   warp-instructions {g(r, 'smsp__inst_executed.sum') / 1e6:.1f} M   stall cycles per issue: {top}\n")
        if t_us > 50:   # the batch launches, not the one-frame reference launches
            traffic[name.split("<")[0]] = {"dram_bytes_per_launch": rd + wr, "frames_per_launch": frames,
                                           "dram_bytes_per_frame": (rd + wr) / frames, "time_us": t_us}
json.dump({"source": f"profiles/{tag}_kernels.txt", "command": cmd, "kernels": traffic},
          open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(root, "profiles", f"{tag}_launches.txt")).read())
print(open(os.path.join(root, "profiles", f"{tag}_kernels.txt")).read())
