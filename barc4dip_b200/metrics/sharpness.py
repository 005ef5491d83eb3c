"""
Sharpness metrics on the B200 path -- drop-in for barc4dip.metrics.sharpness.

tenengrad (:405-476), laplacian_variance (:482-530), spectral_entropy (:536-629),
inverse_autocorr_width (:635-746), sharpness_stats (:89-288), sharpness_stack_stats (:290-399).
`eigenvalues` (:752-861, a dense symmetric eigenproblem) is outside the hot path: it runs on the device
through library code (cuBLAS Gram matrix + cuSOLVER symmetric eigensolver via torch.linalg), see stack.eigenvalues_block.
"""

from __future__ import annotations

import logging
import math
from typing import Sequence

import numpy as np

from .. import engine, stack as blocks
from .._lib import B4DUnsupported
from .common import (apply_display_origin, choose_tiling_mode, normalize_display_origin, normalize_groups, tiled_blocks,
                     tiles_meta)

logger = logging.getLogger(__name__)

_SHARPNESS_UNITS: dict[str, dict[str, str]] = {
    "stats": {"mean": "a.u.", "std": "a.u.", "variance": "a.u.^2", "skewness": "", "kurtosis": "",
              "frac_zero": "", "frac_sat": "", "SNRdB": "dB"},
    "gradient": {"tenengrad": "a.u.^2", "ex": "a.u.^2", "ey": "a.u.^2", "re": ""},
    "laplacian": {"laplacian_variance": "a.u.^2"},
    "spectral": {"spectral_entropy": ""},
    "autocorrelation": {"sx": "1/px", "sy": "1/px", "seq": "1/px", "r": ""},
    "eigenvalues": {"eigenvalues": "", "e1": "", "e2": "", "re": ""},
}
_GROUP_ORDER = ("stats", "gradient", "laplacian", "spectral", "autocorrelation", "eigenvalues")
_ALL_SHARPNESS_GROUPS = {"stats", "gradient", "laplacian", "spectral", "autocorrelation", "eigenvalues"}


def _scalar(block: dict, t: int = 0) -> dict:
    return {k: float(v[t]) for k, v in block.items()}


def _check2d(image, who: str) -> np.ndarray:
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError(f"{who} received an empty image.")
    return data


def tenengrad(image, *, eps: float = 1e-12, verbose: bool = False) -> dict:
    """Mean squared Sobel gradients: tenengrad = ex + ey, re = ex / (ey + eps)."""
    data = _check2d(image, "tenengrad")
    table = engine.frame_reductions(engine.as_stack(data), saturation_value=None)
    out = _scalar(blocks.gradient_block(table, eps))
    if verbose:
        logger.info("> tenengrad: %.6g | ex: %.6g | ey: %.6g | ex/ey: %.3f", out["tenengrad"], out["ex"], out["ey"], out["re"])
    return out


def laplacian_variance(image, *, verbose: bool = False) -> float:
    """Population variance of the 5-point Laplacian (reflect borders) over finite pixels."""
    data = _check2d(image, "laplacian_variance")
    table = engine.frame_reductions(engine.as_stack(data), saturation_value=None)
    var = float(blocks.laplacian_block(table)["laplacian_variance"][0])
    if verbose:
        logger.info("> laplacian variance: %.6g", var)
    return var


def spectral_entropy(image, *, remove_mean: bool = True, remove_dc: bool = True, eps: float = 1e-30,
                     verbose: bool = False) -> float:
    """Normalised Shannon entropy of the PSD (DC excluded)."""
    data = _check2d(image, "spectral_entropy")
    if not (remove_mean and remove_dc):
        raise B4DUnsupported("spectral_entropy on the B200 path is built for the defaults remove_mean=True, remove_dc=True")
    if np.issubdtype(data.dtype, np.floating) and not np.all(np.isfinite(data)):
        raise ValueError("spectral_entropy requires all values to be finite.")
    hn = float(blocks.spectral_entropy_block(engine.as_stack(data))["spectral_entropy"][0])
    if verbose:
        logger.info("> spectral_entropy: %.6g", hn)
    return hn


def inverse_autocorr_width(image, *, fraction: float = 1.0 / math.e, radial_method: str = "interpolated",
                           min_size_px: int = 32, verbose: bool = False) -> dict:
    """Inverse 1/e widths of the standardised, peak-normalised autocorrelation (sx, sy, seq, r = lx/ly)."""
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError("image must be a 2D array.")
    if data.size == 0:
        raise ValueError("inverse_autocorr_width received an empty image.")
    if min(data.shape) < int(min_size_px):
        raise ValueError(f"image too small for inverse autocorrelation width (min dimension < {int(min_size_px)}).")
    if radial_method not in ("binned", "interpolated"):      # both use the interpolated estimator (quirk 9)
        raise ValueError("radial_method must be 'binned' or 'interpolated'.")
    if not (0.0 < fraction < 1.0):
        raise ValueError("fraction must be in (0, 1).")
    out = _scalar(blocks.inverse_autocorr_block(engine.as_stack(data), fraction=fraction))
    if verbose:
        logger.info("> inv_ac_width: sx=%.4g | sy=%.4g | seq=%.4g | r(lx/ly)=%.3g", out["sx"], out["sy"], out["seq"], out["r"])
    return out


def eigenvalues(image, *, k: int = 5, eps: float = 1e-30, verbose: bool = False) -> dict:
    """STA2: sum of the k leading eigenvalues of the covariance of the energy-normalised, mean-removed image, e1, e2, e1/e2."""
    data = _check2d(image, "eigenvalues")
    if not np.all(np.isfinite(data)):
        raise ValueError("eigenvalues requires all values to be finite.")
    out = _scalar(blocks.eigenvalues_block(engine.as_stack(np.ascontiguousarray(data)), k=k, eps=eps))
    if verbose:
        logger.info("> eigenvalues: %.6g | e1: %.6g | e2: %.6g | e1/e2: %.3f | k=%d", out["eigenvalues"], out["e1"], out["e2"],
                    out["re"], min(int(k), min(data.shape)))
    return out


def _resolve_groups(metrics) -> set[str]:
    return normalize_groups(metrics, all_groups=_ALL_SHARPNESS_GROUPS, context="sharpness", param_name="metrics")


def _full_blocks(dev_stack, groups, saturation_value, eps) -> dict:
    """Requested groups for an HBM-resident stack -> {group: {metric: (T,) array}}."""
    out: dict = {}
    table = None
    if blocks.fused_available(dev_stack) and (groups & {"spectral", "autocorrelation"}):
        # one fused pass: a single forward FFT per frame feeds the spectral entropy and the autocorrelation widths
        fb = blocks.FusedBlocks(dev_stack, saturation_value=saturation_value, eps=eps, want_tails=False,
                                want_spectral="spectral" in groups)
        table = fb.table
        if "stats" in groups:
            out["stats"] = blocks.moments_block(table, saturation_value)
        if "gradient" in groups:
            out["gradient"] = blocks.gradient_block(table)
        if "laplacian" in groups:
            out["laplacian"] = blocks.laplacian_block(table)
        if "spectral" in groups:
            out["spectral"] = fb.entropy()
        if "autocorrelation" in groups:
            out["autocorrelation"] = fb.inverse_autocorr()
        if "eigenvalues" in groups:
            out["eigenvalues"] = blocks.eigenvalues_block(dev_stack)
        return out
    if groups & {"stats", "gradient", "laplacian", "autocorrelation"}:
        table = engine.frame_reductions(dev_stack, saturation_value=saturation_value, eps=eps)
    if "stats" in groups:
        out["stats"] = blocks.moments_block(table, saturation_value)
    if "gradient" in groups:
        out["gradient"] = blocks.gradient_block(table)
    if "laplacian" in groups:
        out["laplacian"] = blocks.laplacian_block(table)
    if "spectral" in groups:
        out["spectral"] = blocks.spectral_entropy_block(dev_stack)
    if "autocorrelation" in groups:
        out["autocorrelation"] = blocks.inverse_autocorr_block(dev_stack, table=table)
    if "eigenvalues" in groups:
        out["eigenvalues"] = blocks.eigenvalues_block(dev_stack)
    return out


def sharpness_stats(image, *, metrics="all", tiles: bool = True, display_origin: str = "lower",
                    saturation_value: float | None = 65535.0, eps: float = 1e-6, verbose: bool = True) -> dict:
    """Sharpness metrics of one frame; result schema of the reference ({"meta", "full"[, "tiles"]})."""
    if not isinstance(image, np.ndarray):
        raise TypeError("sharpness_stats expects a numpy.ndarray")
    if image.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={image.ndim}")
    h, w = image.shape
    groups = _resolve_groups(metrics)
    dev = engine.as_stack(np.ascontiguousarray(image))
    if normalize_display_origin(display_origin) == "lower":      # (common.py:70-71; flipped on the device)
        dev = dev.flip(1).contiguous()
    full = _full_blocks(dev, groups, saturation_value, eps)
    out = {"meta": {"kind": "sharpness", "display_origin": display_origin, "input_shape": (int(h), int(w)),
                    "requested_groups": sorted(groups), "units": _SHARPNESS_UNITS, "tile_mode": "off"},
           "full": {}}
    order = _GROUP_ORDER
    for grp in order:
        if grp in full:
            out["full"][grp] = _scalar(full[grp])
    mode, tile_shape_px = choose_tiling_mode(h, w, tiles=tiles)
    if mode != "off":
        out["meta"].update(tiles_meta(h, w, tile_mode=mode, tile_shape_px=tile_shape_px))
        tl = _tiles(dev, mode, groups, saturation_value, eps)
        if tl:
            out["tiles"] = {g: {k: {"mean": v["mean"][0], "std": v["std"][0]} for k, v in f.items()} for g, f in tl.items()}
    if verbose:
        logger.info("\nsharpness stats for a (h x w: %.0f x %.0f) image: %s", h, w, sorted(groups))
    return out


def _tiles(dev_oriented, mode, groups, saturation_value, eps) -> dict:
    """{group: {field: {"mean": (T, 3, 3), "std": (T, 3, 3)}}} of a display-oriented device stack (sharpness.py:213-282)."""
    res = tiled_blocks(dev_oriented, tile_mode=mode, block_fn=lambda tl: _full_blocks(tl, groups, saturation_value, eps))
    return {g: res[g] for g in _GROUP_ORDER if g in res}


def sharpness_stack_stats(stack, *, metrics="all", tiles: bool = True, display_origin: str = "lower",
                          saturation_value: float | None = 65535.0, eps: float = 1e-6, verbose: bool = True,
                          parallel: bool = True, n_jobs: int | None = None) -> dict:
    """Per-frame sharpness metrics of a (T, H, W) stack, every leaf with a leading T axis.

    The whole stack is processed in batched kernels on the GPU; `parallel` / `n_jobs` are accepted for
    signature compatibility and only recorded in the metadata.
    """
    if not isinstance(stack, np.ndarray):
        raise TypeError("sharpness_stack_stats expects a numpy.ndarray")
    if stack.ndim != 3:
        raise ValueError(f"stack must be a 3D array with shape (T, H, W); got ndim={stack.ndim}")
    T, H, W = (int(v) for v in stack.shape)
    if T < 1:
        raise ValueError("stack must contain at least one frame.")
    normalize_display_origin(display_origin)
    groups = _resolve_groups(metrics)
    # full-frame scalars do not depend on the row flip of display_origin="lower" 
    dev = engine.as_stack(stack)
    full = _full_blocks(dev, groups, saturation_value, eps)
    serial = (not parallel) or (n_jobs is not None and int(n_jobs) <= 1)
    meta = {"kind": "sharpness_stack_stats", "input_shape": (H, W), "stack_shape": (T, H, W), "n_frames": T,
            "display_origin": display_origin, "requested_groups": sorted(groups), "units": _SHARPNESS_UNITS,
            "parallel": {"enabled": bool(not serial), "n_jobs": None if serial else (-1 if n_jobs is None else n_jobs)},
            }
    out_full = {grp: full[grp] for grp in _GROUP_ORDER if grp in full}
    out = {"meta": meta, "full": out_full}
    mode, tile_shape_px = choose_tiling_mode(H, W, tiles=tiles)
    meta.update(tiles_meta(H, W, tile_mode=mode, tile_shape_px=tile_shape_px))      # (sharpness.py:384)
    if mode != "off":
        oriented = dev.flip(1) if normalize_display_origin(display_origin) == "lower" else dev
        tl = _tiles(oriented, mode, groups, saturation_value, eps)
        if tl:
            out["tiles"] = tl
    if verbose:
        logger.info("> sharpness_stack_stats | frames=%d | device=cuda", T)
    return out
