"""
FFT-based circular cross/auto-correlation on the B200 path.

Drop-in for barc4dip.signal.corr (xcorr2d :169-253, autocorr2d :256-320).  The reference computes
in float64; the device computes in float32 (1.3e-7 of peak, inside the 1e-5 bar) and the result is
returned as float64 to keep the reference's output dtype.  xcorr2d always returns a real array
(the reference may return complex128 with a ~1e-5 imaginary residue).
"""

from __future__ import annotations

import numpy as np

from .. import engine
from .common import lag_axis, resolve_step_1d, resolve_steps_2d


def _check_norm(normalize: str):
    if normalize not in ("none", "peak"):
        raise ValueError(f"Invalid normalize='{normalize}'. Use 'none' or 'peak'.")


def xcorr2d(a, b, *, x=None, y=None, dx: float = 1.0, dy: float = 1.0, remove_mean: bool = True,
            standardize: bool = False, normalize: str = "peak"):
    aa, bb = np.asarray(a), np.asarray(b)
    if aa.ndim != 2 or bb.ndim != 2:
        raise ValueError("a and b must be 2D arrays.")
    if aa.shape != bb.shape:
        raise ValueError("a and b must have the same shape.")
    ny, nx = aa.shape
    sx, sy = resolve_steps_2d(shape=(ny, nx), x=x, y=y, dx=dx, dy=dy)
    _check_norm(normalize)
    c = engine.xcorr2d(engine.as_stack(aa), engine.as_stack(bb), remove_mean=remove_mean, standardize=standardize,
                       normalize_peak=normalize == "peak")
    return c[0].cpu().numpy().astype(np.float64), lag_axis(nx, sx), lag_axis(ny, sy)


def autocorr2d(a, *, x=None, y=None, dx: float = 1.0, dy: float = 1.0, remove_mean: bool = True,
               standardize: bool = False, normalize: str = "peak"):
    aa = np.asarray(a)
    if aa.ndim != 2:
        raise ValueError("a and b must be 2D arrays.")
    ny, nx = aa.shape
    sx, sy = resolve_steps_2d(shape=(ny, nx), x=x, y=y, dx=dx, dy=dy)
    _check_norm(normalize)
    ac, _ = engine.autocorr2d(engine.as_stack(aa), remove_mean=remove_mean, standardize=standardize,
                              normalize_peak=normalize == "peak")
    return ac[0].cpu().numpy().astype(np.float64), lag_axis(nx, sx), lag_axis(ny, sy)


# ---- 1-D signals: the same kernels on (1, n) frames (signal/corr.py:45-166 of the reference) -----------------------------

def xcorr1d(a, b, *, x=None, dx: float = 1.0, remove_mean: bool = True, standardize: bool = False, normalize: str = "peak"):
    """Circular cross-correlation of two 1-D signals, zero lag at n // 2; (corr, xlag)."""
    aa, bb = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if aa.ndim != 1 or bb.ndim != 1:
        raise ValueError("a and b must be 1D arrays.")
    if aa.size != bb.size:
        raise ValueError("a and b must have the same length.")
    n = int(aa.size)
    xlag = lag_axis(n, resolve_step_1d(n=n, x=x, dx=dx))
    _check_norm(normalize)
    c = engine.xcorr2d(engine.as_stack(aa[None, :]), engine.as_stack(bb[None, :]), remove_mean=remove_mean,
                       standardize=standardize, normalize_peak=normalize == "peak")
    return c[0, 0].cpu().numpy().astype(np.float64), xlag


def autocorr1d(a, *, x=None, dx: float = 1.0, remove_mean: bool = True, standardize: bool = False, normalize: str = "peak"):
    """Circular auto-correlation of a 1-D signal; (corr, xlag)."""
    return xcorr1d(a, a, x=x, dx=dx, remove_mean=remove_mean, standardize=standardize, normalize=normalize)
