"""
Sampling steps, shifted frequency axes and lag axes for the spectral wrappers (host side, scalar and 1-D work only).

The callers may describe the sampling grid either by steps (dx, dy) or by coordinate axes (x, y); the reference accepts
one or the other and insists on uniformly sampled, strictly monotonic axes (signal/common.py:13-87 there). The rules and
the wording of the errors are the reference's, because user code and its tests see them.
"""

from __future__ import annotations

import numpy as np

_UNIFORM_TOL = 1e-6      # largest relative deviation of any spacing from the median spacing


def step_of_axis(values, label: str) -> float:
    """Spacing of a uniformly sampled coordinate axis; ValueError when it is not one."""
    coords = np.asarray(values, dtype=float)
    if coords.ndim != 1 or coords.size < 2:
        raise ValueError(f"{label} must be a 1D array with at least 2 samples.")
    gaps = coords[1:] - coords[:-1]
    if not np.isfinite(gaps).all():
        raise ValueError(f"{label} contains non-finite values.")
    rising, falling = bool((gaps > 0).all()), bool((gaps < 0).all())
    if not (rising or falling):
        raise ValueError(f"{label} must be strictly monotonic (uniform sampling assumed).")
    widths = np.abs(gaps)
    h = float(np.median(widths))
    if h <= 0:
        raise ValueError(f"{label} has non-positive sampling step.")
    spread = float(np.abs(widths - h).max() / h)
    if spread > _UNIFORM_TOL:
        raise ValueError(f"{label} appears non-uniform (max relative deviation {spread:.2e}). "
                         "Provide uniformly sampled axes.")
    return h


def resolve_steps_2d(*, shape, x, y, dx: float, dy: float) -> tuple[float, float]:
    """(step along x, step along y) of an image of `shape` = (ny, nx), from steps or from axes, never both."""
    with_axes = (x is not None, y is not None)
    if with_axes[0] != with_axes[1]:
        raise ValueError("Provide both x and y axes, or neither.")
    if (with_axes[0] and dx != 1.0) or (with_axes[1] and dy != 1.0):
        raise ValueError("Provide either (x, y) or (dx, dy), not both.")
    if not with_axes[0]:
        if min(dx, dy) <= 0:
            raise ValueError("dx and dy must be > 0.")
        return float(dx), float(dy)
    axes = [np.asarray(v, dtype=float) for v in (x, y)]
    if any(a.ndim != 1 for a in axes):
        raise ValueError("x and y must be 1D arrays.")
    if (axes[1].size, axes[0].size) != tuple(int(n) for n in shape):
        raise ValueError("x/y sizes must match (nx, ny) of the image.")
    return step_of_axis(axes[0], "x"), step_of_axis(axes[1], "y")


def resolve_step_1d(*, n: int, x, dx: float, name: str = "x") -> float:
    """Step of a 1-D signal of length n, from dx or from a uniformly sampled axis (never both)."""
    if x is None:
        if dx <= 0:
            raise ValueError(f"d{name} must be > 0.")
        return float(dx)
    if dx != 1.0:
        raise ValueError(f"Provide either {name} or d{name}, not both.")
    axis = np.asarray(x, dtype=float)
    if axis.ndim != 1:
        raise ValueError(f"{name} must be a 1D array.")
    if axis.size != n:
        raise ValueError(f"{name}.size must match the signal length ({n}).")
    return step_of_axis(axis, name)


def freq_axis1d(*, n: int, x=None, dx: float = 1.0) -> np.ndarray:
    """Shifted frequency axis of a 1-D signal, cycles per unit."""
    if n < 1:
        raise ValueError("n must be >= 1.")
    return np.fft.fftshift(np.fft.fftfreq(int(n), d=resolve_step_1d(n=n, x=x, dx=dx)))


def freq_axes2d(*, shape, x=None, y=None, dx: float = 1.0, dy: float = 1.0):
    ny, nx = shape
    if ny < 1 or nx < 1:
        raise ValueError("shape must contain positive integers.")
    sx, sy = resolve_steps_2d(shape=shape, x=x, y=y, dx=dx, dy=dy)
    return (np.fft.fftshift(np.fft.fftfreq(int(nx), d=sx)), np.fft.fftshift(np.fft.fftfreq(int(ny), d=sy)))


def lag_axis(n: int, step: float) -> np.ndarray:
    """Lags of a shifted correlation of length n: zero lag at index n // 2."""
    centre = n // 2
    return float(step) * (np.arange(n, dtype=float) - centre)
