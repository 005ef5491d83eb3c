#!/usr/bin/env python
"""
Schedule sweep of the fused stack pipeline (b4d_set_schedule): frames per pipelined step, ring slots, cache policy.
One process, one resident stack; every schedule is checked against the whole-batch schedule (all outputs) and timed
with CUDA events; the per-kernel-class event spans are printed beside it.

    python scripts/sched_sweep.py [--frames 128] [--steps 10] [--configs "sub:lanes:slots:keep:graphs,..."]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--frames", type=int, default=128)
    p.add_argument("--size", type=int, default=2048)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--configs", default="0:1:1:0:0,1:2:1:1:1,2:2:1:1:1,1:3:1:1:1,2:3:1:1:1")
    p.add_argument("--out", default="")
    p.add_argument("--no-track", action="store_true", help="PSD + autocorrelation + reductions only (no tracker chain)")
    p.add_argument("--no-psd", action="store_true", help="no PSD map (autocorrelation map + tracker + reductions)")
    p.add_argument("--no-prof", action="store_true", help="no per-kernel event spans (fewer driver calls per launch)")
    a = p.parse_args()
    import torch
    from barc4dip_b200 import synth
    from barc4dip_b200._lib import get_context
    from barc4dip_b200.pipeline import StackAnalyzer
    sys.path.insert(0, ROOT)
    import bench

    dev = torch.device("cuda:0")
    n, F = a.size, a.frames
    ctx = get_context(0)
    base = synth.speckle_frame(n, grain=6.0, seed=0)
    shifts = bench.make_shifts(F, seed=2)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    base_d = torch.from_numpy(base).to(dev)
    stack = torch.empty((F, n, n), dtype=torch.float32, device=dev)
    sigma = 0.01 * float(base.mean())
    for t in range(F):
        stack[t] = torch.roll(base_d, (int(shifts[t, 0]), int(shifts[t, 1])), dims=(0, 1)) + \
            sigma * torch.randn((n, n), generator=g, device=dev)
    analyzer = StackAnalyzer((n, n), device=0, chunk_frames=F, want_maps=not a.no_psd, want_contrast=True)
    if not a.no_track:
        analyzer.set_reference(stack[0].clone())
    psd = torch.empty((F, n, n), dtype=torch.float32, device=dev)
    ac = torch.empty((F, n, n), dtype=torch.float32, device=dev)

    def step():
        return analyzer.run_device(stack, psd_out=None if a.no_psd else psd, ac_out=ac, resolve_tails=False)

    def snapshot(res):
        out = {k: v.clone() for k, v in res.items() if isinstance(v, torch.Tensor) and v.numel() < 10_000_000}
        out["psd_sum"] = psd.double().sum(dim=(1, 2))
        out["ac_sum"] = ac.double().abs().sum(dim=(1, 2))
        out["psd_probe"] = psd[:, ::97, ::89].clone()
        out["ac_probe"] = ac[:, ::97, ::89].clone()
        return out

    ref = None
    rows = []
    for cfg in a.configs.split(","):
        sub, lanes, slots, keep, graphs, pair = (int(x) for x in (cfg.split(":") + ["0"])[:6])
        ctx.set_schedule(sub, lanes, slots, keep, graphs)
        ctx.set_pairing(pair)
        psd.zero_(); ac.zero_()
        for _ in range(3):
            res = step()
        torch.cuda.synchronize()
        snap = snapshot(res)
        ok = "ref"
        if ref is None:
            ref = snap
        else:
            bad = []
            for k, v in ref.items():
                w = snap[k]
                same = torch.equal(torch.nan_to_num(v.double(), nan=-7.0), torch.nan_to_num(w.double(), nan=-7.0))
                if not same:
                    d = (torch.nan_to_num(v.double()) - torch.nan_to_num(w.double())).abs().max().item()
                    bad.append(f"{k}:{d:.3g}")
            ok = "same" if not bad else "DIFF " + " ".join(bad)
        if not a.no_prof:
            ctx.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        prof = ctx.profile_end() if not a.no_prof else {}
        ms = e0.elapsed_time(e1) / a.steps
        kern = {k: round(v[0] / a.steps, 3) for k, v in prof.items() if v[1] > 0 and v[0] / a.steps > 0.02}
        row = {"sub": sub, "lanes": lanes, "slots": slots, "keep": keep, "graphs": graphs, "pair": pair, "ms_per_step": round(ms, 3), "fps": round(F / ms * 1e3), "check": ok, "kernels": kern}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
