"""
Build libb4d.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m barc4dip_b200.build          # incremental
    python -m barc4dip_b200.build --force
    python -m barc4dip_b200.build --variant NAME [-DFLAG ...]   # A/B builds: barc4dip_b200/variants/libb4d_NAME.so
                                                                # (picked up with B4D_LIB=<path>, see _lib.py)

The .so lands next to this file (barc4dip_b200/libb4d.so) so that it travels to the GPU box
with the repository snapshot; it is git-ignored.
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb4d.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> str:
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, variant: str | None = None, defines: list[str] | None = None) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    OBJDIR, LIB, NVCC_FLAGS = globals()["OBJDIR"], globals()["LIB"], list(globals()["NVCC_FLAGS"])
    if variant:
        OBJDIR = os.path.join(HERE, "variants", "obj_" + variant)
        LIB = os.path.join(HERE, "variants", f"libb4d_{variant}.so")
        NVCC_FLAGS += list(defines or [])
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(OBJDIR, "headers.sha1")
    digest = _headers_digest() + " ".join(defines or [])
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if old != digest:
        force = True
    objs, rebuilt = [], False
    procs = []
    for src in _sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var,
                 defines=[a for a in sys.argv[1:] if a.startswith("-D")])
    print(path)
