"""Axis / step validation shared by fft.py and corr.py (host-side; mirrors signal/common.py of the reference)."""

from __future__ import annotations

import numpy as np


def _uniform_step(axis, name: str) -> float:
    a = np.asarray(axis, dtype=float)
    if a.ndim != 1 or a.size < 2:
        raise ValueError(f"{name} must be a 1D array with at least 2 samples.")
    d = np.diff(a)
    if not np.all(np.isfinite(d)):
        raise ValueError(f"{name} contains non-finite values.")
    if not (np.all(d > 0) or np.all(d < 0)):
        raise ValueError(f"{name} must be strictly monotonic (uniform sampling assumed).")
    mag = np.abs(d)
    step = float(np.median(mag))
    if step <= 0:
        raise ValueError(f"{name} has non-positive sampling step.")
    rel = float(np.max(np.abs(mag - step)) / step)
    if rel > 1e-6:
        raise ValueError(f"{name} appears non-uniform (max relative deviation {rel:.2e}). "
                         "Provide uniformly sampled axes.")
    return step


def resolve_steps_2d(*, shape, x, y, dx: float, dy: float) -> tuple[float, float]:
    """(dx, dy) from either explicit steps or uniformly sampled axes (ref: signal/common.py:58-87)."""
    ny, nx = shape
    if (x is None) ^ (y is None):
        raise ValueError("Provide both x and y axes, or neither.")
    if (x is not None and dx != 1.0) or (y is not None and dy != 1.0):
        raise ValueError("Provide either (x, y) or (dx, dy), not both.")
    if x is None:
        if dx <= 0 or dy <= 0:
            raise ValueError("dx and dy must be > 0.")
        return float(dx), float(dy)
    xa, ya = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    if xa.ndim != 1 or ya.ndim != 1:
        raise ValueError("x and y must be 1D arrays.")
    if xa.size != nx or ya.size != ny:
        raise ValueError("x/y sizes must match (nx, ny) of the image.")
    return _uniform_step(xa, "x"), _uniform_step(ya, "y")


def lag_axis(n: int, step: float) -> np.ndarray:
    """(arange(n) - n//2) * step (ref: signal/common.py:89-90)."""
    return (np.arange(n, dtype=float) - (n // 2)) * float(step)
