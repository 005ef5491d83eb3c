"""One small invocation of the hot path on cuda:0, checked against the numpy oracle (used by __graft_entry__.smoke)."""

from __future__ import annotations

import numpy as np


def run(verbose: bool = False) -> None:
    import torch
    from oracle import ref_numpy as orc          # the checker, never the product
    from . import engine, synth
    from .metrics.statistics import moments_from_row

    n, T = 256, 3
    stack, _ = synth.tracking_stack(T, n, grain=5.0, seed=7, step_sigma=0.6)
    dev = engine.as_stack(stack)
    engine.PhaseTracker(stack[0], (n, n), y0=0, x0=0)
    out = engine.stack_pipeline(dev)
    torch.cuda.synchronize()
    fr = out["reductions"].cpu().numpy()
    psd = out["psd"].cpu().numpy()
    ac = out["autocorr"].cpu().numpy()
    grain = out["grain"].cpu().numpy()
    track = out["tracking"].cpu().numpy()
    full = (slice(0, n), slice(0, n))
    for t in range(T):
        want = orc.distribution_moments(stack[t])
        got = moments_from_row(fr[t], 65535.0)
        for k in ("mean", "std", "skewness", "kurtosis"):
            assert abs(got[k] - want[k]) <= 1e-4 * abs(want[k]), (t, k, got[k], want[k])
        ten = orc.tenengrad(stack[t])
        assert abs(fr[t, 7] / fr[t, 0] - ten["ex"]) <= 1e-4 * ten["ex"]
        P, _, _ = orc.psd2d(stack[t])
        assert np.max(np.abs(psd[t] - P)) <= 1e-5 * P.max(), "psd2d parity"
        g = orc.grain(stack[t])
        assert np.max(np.abs(ac[t] - g["autocorr"])) <= 1e-5, "autocorr2d parity"
        for i, k in enumerate(("lx", "ly", "leq")):
            assert abs(grain[t, i] - g[k]) <= 1e-4 * g[k], (t, k, grain[t, i], g[k])
        dy, dx, peak, snr = orc.phase_correlation(stack[0], stack[t], slices_yx=full)
        assert abs(track[t, 0] - dy) <= 0.01 and abs(track[t, 1] - dx) <= 0.01, (t, track[t], dy, dx)
        assert abs(track[t, 2] - peak) <= 1e-3 * abs(peak)
    if verbose:
        print(f"smoke ok: fused stack pipeline on {T} x {n}^2 frames matches the oracle "
              f"(launches so far: {engine.get_context(0).launches})")
