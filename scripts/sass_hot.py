#!/usr/bin/env python
"""Summarise an ncu source-page CSV: stall samples per SASS region and the hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K | python scripts/sass_hot.py [nbuckets]"""
import csv, sys
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rows = list(csv.reader(sys.stdin))
# the dump may hold several kernels / launches: take the first block
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r, k: float(r[ix[k]] or 0)
tot = sum(S(r, "# Samples") for r in body)
inst = sum(S(r, "Instructions Executed") for r in body)
print(f"{len(body)} SASS instructions, {tot:.0f} samples, {inst:.3g} warp-instructions executed")
size = max(1, len(body) // nb)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for b in range(0, len(body), size):
    chunk = body[b:b + size]
    s = sum(S(r, "# Samples") for r in chunk)
    e = sum(S(r, "Instructions Executed") for r in chunk)
    top = sorted(((sum(S(r, k) for r in chunk), k[6:]) for k in stalls), reverse=True)[:3]
    ops = {}
    for r in chunk:
        op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[ix["Source"]].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + S(r, "Instructions Executed")
    topops = ", ".join(f"{k}:{v / max(e, 1) * 100:.0f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print(f"  sass[{b:5d}:{b + len(chunk):5d}] samples {100 * s / tot:5.1f}%  exec {100 * e / inst:5.1f}%  "
          f"stalls {', '.join(f'{k}:{100 * v / max(s, 1):.0f}%' for v, k in top)} | {topops}")
print("hottest instructions:")
for r in sorted(body, key=lambda r: -S(r, "# Samples"))[:14]:
    print(f"  {100 * S(r, '# Samples') / tot:5.2f}%  {r[ix['Source']].strip()[:90]}")
