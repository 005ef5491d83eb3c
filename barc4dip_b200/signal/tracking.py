"""
Translation tracking by FFT phase correlation on the B200 path.

Drop-in for barc4dip.signal.tracking: track_translation (:25-78) and phase_correlation (:192-297)
with backend="internal".  The registry seam of the reference is kept (`_TRACKERS`, `_register`);
the "template" method (cv2 / skimage matchTemplate) is a "next" row of SURVEY.md 8(f) and raises.
"""

from __future__ import annotations

from typing import Callable

import numpy as np

from .. import engine

_Tracker = Callable[..., tuple[float, float, float, float]]
_TRACKERS: dict[str, _Tracker] = {}


def _register(method: str):
    key = method.strip().lower()

    def deco(fn: _Tracker) -> _Tracker:
        _TRACKERS[key] = fn
        return fn

    return deco


def track_translation(template, image, *, slices_yx=None, method: str = "phase", backend: str = "internal",
                      subpixel: bool = True, eps: float = 1e-9):
    fn = _TRACKERS.get(method.strip().lower())
    if fn is None:
        raise ValueError(f"Unsupported tracking method: {method!r}. Supported: {', '.join(sorted(_TRACKERS))}")
    return fn(template, image, slices_yx=slices_yx, backend=backend, subpixel=subpixel, eps=eps)


def _as_float2d(a, *, name: str) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError(f"{name} must be a 2D array.")
    return a if np.issubdtype(a.dtype, np.floating) else a.astype(np.float32, copy=False)


def centered_slices(image_shape, size_yx):
    """Centred ROI slices; sizes must be odd (ref: geometry/roi.py:44-106)."""
    H, W = image_shape
    sy, sx = size_yx
    if sy <= 0 or sx <= 0:
        raise ValueError("ROI sizes must be positive.")
    if sy % 2 == 0 or sx % 2 == 0:
        raise ValueError("ROI sizes must be odd for symmetry.")
    y0, x0 = H // 2 - sy // 2, W // 2 - sx // 2
    if y0 < 0 or y0 + sy > H or x0 < 0 or x0 + sx > W:
        raise ValueError("ROI exceeds image bounds.")
    return slice(y0, y0 + sy), slice(x0, x0 + sx)


@_register("phase")
def phase_correlation(template, image, *, slices_yx=None, backend: str = "internal", subpixel: bool = True,
                      eps: float = 1e-9):
    """(dy, dx, peak, snr): shift to apply to the template to align it with the image (+dy down, +dx right)."""
    tpl = _as_float2d(template, name="template")
    img = _as_float2d(image, name="image")
    H, W = img.shape
    if slices_yx is None:
        slices_yx = centered_slices((H, W), tpl.shape)
    sy, sx = slices_yx
    if tpl.shape != (sy.stop - sy.start, sx.stop - sx.start):
        raise ValueError("ROI shape does not match target slice dimensions.")
    if backend == "skimage":
        raise ImportError("backend='skimage' requires scikit-image; the B200 path implements backend='internal'.")
    if backend != "internal":
        raise ValueError("backend must be 'internal' or 'skimage'.")
    tracker = engine.PhaseTracker(tpl, (H, W), y0=int(sy.start), x0=int(sx.start), eps=eps)
    dy, dx, peak, snr = tracker.track(engine.as_stack(img), subpixel=subpixel)[0]
    return float(dy), float(dx), float(peak), float(snr)


@_register("template")
def template_matching(template, image, *, slices_yx=None, backend: str = "opencv", subpixel: bool = True,
                      eps: float = 1e-9):
    """(dy, dx, peak, snr) by normalised cross-correlation of the z-scored template against the image.

    Both of the reference's backends evaluate the same quantity -- cv2.matchTemplate(TM_CCOEFF_NORMED) and
    skimage.feature.match_template(pad_input=False) -- and the B200 path computes it once for either name (FFT
    correlation + double-precision window sums, csrc/spectral.cu); parity is pinned against the opencv backend, the
    only one runnable where the goldens were recorded."""
    tpl = _as_float2d(template, name="template")
    img = _as_float2d(image, name="image")
    H, W = img.shape
    h, w = tpl.shape
    if h > H or w > W:
        raise ValueError(f"template shape {(h, w)} must fit inside image shape {(H, W)}")
    if backend not in ("opencv", "skimage"):
        raise ValueError("backend must be 'opencv' or 'skimage'.")
    if slices_yx is None:
        slices_yx = centered_slices((H, W), (h, w))
    sy, sx = slices_yx
    y0 = (sy.start + sy.stop - 1) / 2.0
    x0 = (sx.start + sx.stop - 1) / 2.0
    dy, dx, peak, snr = engine.template_match(tpl, engine.as_stack(img), ref_center_yx=(y0, x0), subpixel=subpixel, eps=eps)[0]
    return float(dy), float(dx), float(peak), float(snr)
