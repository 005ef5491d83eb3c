#!/usr/bin/env python
"""Per-launch summary of an .ncu-rep: time, DRAM bytes, throughputs, occupancy, top stalls, opcode mix.
usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep [kernel-substring] [--ops]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return float("nan")
for r in data:
    name = r[ix["Kernel Name"]]
    if filt and filt not in name: continue
    t = g(r, "gpu__time_duration.sum")
    tu = units[ix["gpu__time_duration.sum"]]
    t_us = t * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(tu, 1.0)
    def gb(k):
        v = g(r, k); u = units[ix[k]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    print(f"== {name[:70]}  grid {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}  regs {r[ix['launch__registers_per_thread']]}")
    print(f"   time {t_us:.1f} us  dram rd {rd/1e6:.1f} MB wr {wr/1e6:.1f} MB -> {(rd+wr)/t_us/1e3:.0f} GB/s"
          f"  | dram% {g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} lts% {g(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f}"
          f" l1tex% {g(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} sm% {g(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f}"
          f" issue% {g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} warps% {g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}")
    st = {h.split("issue_stalled_")[1].split("_per_")[0]: g(r, h) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print("   stalls/issue: " + ", ".join(f"{k}:{v:.2f}" for k, v in top)
          + f" | inst {g(r,'smsp__inst_executed.sum')/1e6:.1f}M  smem-wavefronts {g(r,'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')/1e6:.1f}M"
          + f" conflicts {g(r,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')/1e6:.2f}M")
if "--ops" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur)
        elif r and r[0] == "Address": cur["hdr"] = r
        elif r and cur is not None: cur["rows"].append(r)
    for b in blocks:
        if filt and filt not in b["name"]: continue
        ix2 = {h: i for i, h in enumerate(b["hdr"])}
        S = lambda r, k: float(r[ix2[k]] or 0)
        inst = sum(S(r, "Instructions Executed") for r in b["rows"])
        samp = sum(S(r, "# Samples") for r in b["rows"])
        ops, sto = collections.Counter(), collections.Counter()
        for r in b["rows"]:
            src_ = r[ix2["Source"]].split()
            if not src_: continue
            op = (src_[1] if src_[0].startswith("@") else src_[0]).rstrip(";")
            ops[op] += S(r, "Instructions Executed"); sto[op] += S(r, "# Samples")
        print(f"== ops {b['name'][:60]}  {inst/1e6:.1f}M warp-inst, {len(b['rows'])} SASS")
        print("   " + "  ".join(f"{k}:{100*v/inst:.1f}%" for k, v in ops.most_common(28)))
        print("   samples by op: " + "  ".join(f"{k}:{100*v/max(samp,1):.1f}%" for k, v in sto.most_common(14)))
