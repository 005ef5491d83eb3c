"""
Shifted 2-D FFT and power spectral density on the B200 path.

Drop-in for barc4dip.signal.fft (fft2d :198-237, psd2d :261-309, freq_axes2d :58-96): same
signatures, DC-centred outputs, shifted frequency axes.  float32 input gives complex64 / float32
output like the reference on numpy >= 2; other input dtypes are computed in float32 on the device
and returned in the dtype the reference would return (complex128 / float64).
"""

from __future__ import annotations

import numpy as np

from .. import engine
from .common import freq_axes2d, freq_axis1d, resolve_step_1d, resolve_steps_2d


def _out_real_dtype(img: np.ndarray):
    return np.float32 if img.dtype == np.float32 else np.float64


def fft2d(image, *, x=None, y=None, dx: float = 1.0, dy: float = 1.0):
    """F = fftshift(fft2(image)) with shifted (fx, fy)."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    fx, fy = freq_axes2d(shape=img.shape, x=x, y=y, dx=dx, dy=dy)
    F = engine.fft2d(engine.as_stack(img))[0].cpu().numpy()
    if img.dtype != np.float32:
        F = F.astype(np.complex128)
    return F, fx, fy


def ifft2d(F):
    """Inverse of fft2d: ifft2(ifftshift(F)) of a shifted complex spectrum (signal/fft.py:240-258)."""
    F = np.asarray(F)
    if F.ndim != 2:
        raise ValueError("F must be a 2D array.")
    torch = engine.require_cuda()
    dev = torch.from_numpy(np.ascontiguousarray(F, dtype=np.complex64)).to(f"cuda:{engine._lib.default_device()}")
    out = engine.ifft2d(dev[None])[0].cpu().numpy()
    return out if F.dtype == np.complex64 else out.astype(np.complex128)


def psd2d(image, *, x=None, y=None, dx: float = 1.0, dy: float = 1.0, scale: bool = True):
    """P = |fftshift(fft2(image))|^2, times dx*dy/(nx*ny) when scale is True. No mean removal."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    ny, nx = img.shape
    sx, sy = resolve_steps_2d(shape=(ny, nx), x=x, y=y, dx=dx, dy=dy)
    fx, fy = freq_axes2d(shape=(ny, nx), x=x, y=y, dx=dx, dy=dy)
    factor = (sx * sy) / (float(nx) * float(ny)) if scale else 1.0
    P, _ = engine.psd2d(engine.as_stack(img), scale_factor=factor)
    return P[0].cpu().numpy().astype(_out_real_dtype(img), copy=False), fx, fy


# ---- 1-D signals: the same kernels on (1, n) frames (signal/fft.py:31-196 of the reference) ------------------------------

def _signal1d(signal, name: str = "signal") -> np.ndarray:
    s = np.asarray(signal)
    if s.ndim != 1:
        raise ValueError(f"{name} must be a 1D array.")
    return s


def fft1d(signal, *, x=None, dx: float = 1.0):
    """fftshift(fft(signal)) and the shifted frequency axis."""
    s = _signal1d(signal)
    fx = freq_axis1d(n=int(s.size), x=x, dx=dx)
    F, _, _ = fft2d(s[None, :])
    return F[0], fx


def ifft1d(F):
    """ifft(ifftshift(F)) of a shifted 1-D spectrum."""
    return ifft2d(_signal1d(F, "F")[None, :])[0]


def psd1d(signal, *, x=None, dx: float = 1.0, scale: bool = True):
    """|fft1d(signal)|^2, times dx / n when scale is True."""
    s = _signal1d(signal)
    n = int(s.size)
    step = resolve_step_1d(n=n, x=x, dx=dx)
    fx = freq_axis1d(n=n, x=x, dx=dx)
    P, _ = engine.psd2d(engine.as_stack(s[None, :]), scale_factor=(step / float(n)) if scale else 1.0)
    return P[0, 0].cpu().numpy().astype(_out_real_dtype(s), copy=False), fx
