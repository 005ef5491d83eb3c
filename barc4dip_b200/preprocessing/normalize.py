"""
flat_field_correction on the B200 path -- drop-in for barc4dip.preprocessing.normalize (:12-145).

(I - D) / (F - D) * scale with the reference's float32 operation order (bit-identical output), bad-pixel
mask den <= eps -> 0.  The two medians (default eps and the "flat_median" scale) are exact radix selects on
the device.  `bad_pixel_removal=True` repairs the bad pixels with the reference's 3x3 median (:134-140), in place on the
device, bit-exact (b4d_bad_pixel_repair).
"""

from __future__ import annotations

import numpy as np

from .. import engine
from .._lib import B4DUnsupported, require_cuda


def _collapse(a, dev):
    """flats/darks: 2-D as is, 3-D -> float32 mean over axis 0 (normalize.py:83-93)."""
    if a is None:
        return None
    a = np.asarray(a)
    if a.ndim not in (2, 3):
        raise ValueError("flats/darks must be 2D or 3D")
    d = engine.as_stack(np.ascontiguousarray(a, dtype=np.float32), dev)
    if a.ndim == 2:
        return d[0]
    # float32 sum over axis 0 in frame order, then / n: the same operation order as numpy's mean(axis=0)
    acc = engine.TemporalAccumulator(d.shape[1], d.shape[2], device=dev)
    return acc.pilot(d, n_frames=d.shape[0])


def _median_f32(x_dev) -> np.float32:
    """np.median of a float32 device vector (exact: mean of the two middle values in float32)."""
    vals, nv = engine.select_quantiles(x_dev.reshape(1, 1, -1), [0.5])
    lo, hi = np.float32(vals[0, 0]), np.float32(vals[0, 1])
    return lo if int(nv[0]) % 2 == 1 else np.float32((lo + hi) / np.float32(2))


def resolve_flat_field(flat_dev, dark_dev, *, scale: str, eps):
    """(eps, scale_value, apply_scale) exactly as the reference resolves them (normalize.py:104-127)."""
    torch = require_cuda()
    den = engine.sub(flat_dev, dark_dev)
    if eps is None:
        med = _median_f32(den)
        eps = np.float32(1e-6) * med if med > 0 else 1e-6
    eps = float(eps)
    scale_value, apply_scale = 1.0, scale != "none"
    if apply_scale:
        valid = den[den > eps]                      # device-side compaction (plumbing), statistics below on the device
        if scale == "flat_mean":
            tab = engine.frame_reductions(valid.reshape(1, 1, -1), saturation_value=None)
            # np.mean of a float32 array accumulates pairwise in float32 and returns float32
            scale_value = float(np.float32(tab[0, 1]))
        else:
            scale_value = float(_median_f32(valid))
    return eps, scale_value, apply_scale


def flat_field_correction(images, *, flats=None, darks=None, scale: str = "flat_median", bad_pixel_removal: bool = False,
                          eps: float | None = None, verbose: bool = False) -> np.ndarray:
    """Flat-field (gain) correction of one frame (H, W) or a stack (N, H, W); float32 result of the same shape."""
    if scale not in {"none", "flat_mean", "flat_median"}:
        raise ValueError(f"Invalid scale option: {scale}")
    images = np.asarray(images)
    if images.ndim not in (2, 3):
        raise ValueError("images must be 2D or 3D")
    img = images.astype(np.float32, copy=False)
    if flats is None and darks is None:
        return img.copy()
    torch = require_cuda()
    dev_img = engine.as_stack(img)
    dev = dev_img.device.index or 0
    flat, dark = _collapse(flats, dev), _collapse(darks, dev)
    if flat is not None and tuple(flat.shape) != tuple(dev_img.shape[1:]) or dark is not None and tuple(dark.shape) != tuple(dev_img.shape[1:]):
        raise ValueError("flats/darks shape does not match the images")
    if flat is None:
        out = engine.sub(dev_img, dark.expand_as(dev_img).contiguous())
    else:
        eps_v, s, apply_scale = resolve_flat_field(flat, dark, scale=scale, eps=eps)
        out = engine.flat_field(dev_img, flat, dark, eps=eps_v, scale_value=s, apply_scale=apply_scale)
        if bad_pixel_removal:
            engine.bad_pixel_repair(out, flat, dark, eps=eps_v)
    res = out.cpu().numpy()
    return res[0] if images.ndim == 2 else res
