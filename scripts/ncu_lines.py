#!/usr/bin/env python
"""Join an ncu SASS profile with nvdisasm line info: executed warp-instructions and stall samples per source line.
usage: python scripts/ncu_lines.py X.ncu-rep <kernel-substring (mangled ok)> [top] [--lib path/to/lib.so]
The kernel substring is matched against the demangled name in the report and, with spaces/punctuation stripped, is
used to pick the function in the cubin by its template arguments (give e.g. 'cols_kernel<2048, 4, 0, 1, 1>')."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, filt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else os.path.join(os.path.dirname(__file__), "..", "barc4dip_b200", "libb4d.so")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif r and r[0] == "Address": cur["hdr"] = r
    elif r and cur is not None: cur["rows"].append(r)
norm = lambda s: re.sub(r"[^A-Za-z0-9<>,]", "", s.replace("(int)", "").replace("(bool)", ""))
nth = int(sys.argv[sys.argv.index("--nth") + 1]) if "--nth" in sys.argv else 0
cands = [b for b in blocks if norm(filt) in norm(b["name"])]
if "--biggest" in sys.argv:
    def _tot(b):
        i = b["hdr"].index("Instructions Executed")
        return sum(float(r[i] or 0) for r in b["rows"])
    blk = max(cands, key=_tot)
else:
    blk = cands[nth]
ix = {h: i for i, h in enumerate(blk["hdr"])}
# mangled pattern from 'name<args>'
m = re.match(r"(\w+)<(.*)>", filt.strip())
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
pat = None
if m:
    args = [a.strip() for a in m.group(2).split(",")]
    enc = "".join(("Li%sE" % a) if (a.isdigit() and int(a) > 1) else ("Lb%sE" % a if a in ("0", "1") else "Li%sE" % a) for a in args)
    pat = m.group(1) + "I" + enc
lines_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, curline = False, None
    for ln in out.splitlines():
        if ln.startswith("//--------------------- .text."):
            infn = (pat or m.group(1) if m else filt) in ln
            curline = None
            continue
        if not infn: continue
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if mm:
            if "inlined at" not in ln or curline is None: curline = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
        if mm: lines_of[int(mm.group(1), 16)] = curline
    if lines_of: break
S = lambda r, k: float(r[ix[k]] or 0)
base = min(int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]]) for r in blk["rows"])
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
tot_i = tot_s = 0.0
for r in blk["rows"]:
    a = r[ix["Address"]]
    off = (int(a, 16) if a.startswith("0x") else int(a)) - base
    key = lines_of.get(off)
    agg[key][0] += S(r, "Instructions Executed"); agg[key][1] += S(r, "# Samples")
    agg[key][2] += S(r, "L1 Wavefronts Shared Excessive"); agg[key][3] += S(r, "L1 Wavefronts Shared")
    tot_i += S(r, "Instructions Executed"); tot_s += S(r, "# Samples")
print(f"{blk['name'][:80]}: {tot_i/1e6:.1f}M warp-inst, {tot_s:.0f} samples, {len(lines_of)} SASS mapped")
srcs = {}
def text(key):
    if not key: return ""
    f, l = key
    for d in ("barc4dip_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(__file__), "..", d, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][l - 1].strip()[:100] if l - 1 < len(srcs[p]) else ""
    return ""
sortcol = 2 if "--smem" in sys.argv else (1 if "--samples" in sys.argv else 0)
for key, (i, s, ex, wf) in sorted(agg.items(), key=lambda kv: -kv[1][sortcol])[:top]:
    print(f"  inst {100*i/tot_i:5.1f}%  samp {100*s/max(tot_s,1):5.1f}%  smem-wf {wf/1e6:6.2f}M excess {ex/1e6:5.2f}M  {key}  {text(key)}")
