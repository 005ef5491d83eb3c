"""
Speckle-field metrics on the B200 path -- drop-in for barc4dip.metrics.speckles.

amplitude (:602-663), grain (:497-596), bandwidth (:669-817), speckle_stats (:83-255),
speckle_stack_stats (:258-490, with tracking_method="phase", tracking_backend="internal").
"""

from __future__ import annotations

import logging
import math

import numpy as np

from .. import engine, parallel as _par, stack as blocks
from .._lib import B4DUnsupported
from ..signal.common import lag_axis
from .common import (apply_display_origin, choose_tiling_mode, normalize_display_origin, normalize_groups, tiled_blocks,
                     tiles_meta)

logger = logging.getLogger(__name__)

_SPECKLE_UNITS: dict[str, dict[str, str]] = {
    "amplitude": {"visibility": "", "contrast": ""},
    "stats": {"mean": "a.u.", "std": "a.u.", "variance": "a.u.^2", "skewness": "", "kurtosis": "",
              "frac_zero": "", "frac_sat": "", "SNRdB": "dB"},
    "grain": {"lx": "px", "ly": "px", "leq": "px", "r": "", "xlag": "px", "ylag": "px", "autocorr": ""},
    "bandwidth": {"spr": "", "feq": "1/px", "f95": "1/px", "sig_fx": "1/px", "sig_fy": "1/px", "rf": ""},
    "temporal": {"dx": "px", "dy": "px", "r": "px", "std_dx": "px", "std_dy": "px", "std_r": "px"},
}
_ALL_SPECKLE_GROUPS = {"amplitude", "grain", "bandwidth", "stats"}


def _scalar(block: dict, t: int = 0) -> dict:
    return {k: float(v[t]) for k, v in block.items()}


def amplitude(image, verbose: bool = False) -> dict:
    """visibility = std/mean and the robust Michelson contrast from the 0.05 / 99.95 percentiles."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    out = _scalar(blocks.amplitude_block(engine.as_stack(img)))
    if verbose:
        logger.info("> visibility: %.2f | contrast: %.2f", out["visibility"], out["contrast"])
    return out


def grain(image, *, fraction: float = 1.0 / math.e, radial_method: str = "interpolated", verbose: bool = False) -> dict:
    """1/e widths (lx, ly, leq, r = lx/ly) of the peak-normalised autocorrelation, plus the map and lag axes."""
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError("image must be a 2D array.")
    if min(data.shape) < 128:
        raise ValueError("image too small for speckle grain metrics (min dimension < 128).")
    if radial_method == "binned":
        raise B4DUnsupported("grain(radial_method='binned') is not built on the B200 path; use 'interpolated'")
    if radial_method != "interpolated":
        raise ValueError("radial_method must be 'binned' or 'interpolated'.")
    if not (0.0 < fraction < 1.0):
        raise ValueError("fraction must be in (0, 1).")
    g, ac = blocks.grain_block(engine.as_stack(data), fraction=fraction, return_map=True)
    n = int(ac.shape[-1])
    out = _scalar(g)
    metrics = {"lx": out["lx"], "ly": out["ly"], "leq": out["leq"], "r": out["r"],
               "autocorr": ac[0].cpu().numpy().astype(np.float64), "xlag": lag_axis(n, 1.0), "ylag": lag_axis(n, 1.0)}
    if verbose:
        logger.info("> grain: lx=%.2f | ly=%.2f | lx/ly=%.2f | leq=%.2f ", metrics["lx"], metrics["ly"], metrics["r"], metrics["leq"])
    return metrics


def bandwidth(image, verbose: bool = False) -> dict:
    """PSD bandwidth metrics inside the inscribed circle: feq, f95, sig_fx, sig_fy, rf, spr."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    spectral = _scalar(blocks.bandwidth_block(engine.as_stack(img)))
    if verbose:
        logger.info("> bandwidth: fx=%.4f | fy=%.4f | fx/fy=%.2f | feq=%.4f | f95=%.4f | spr=%.0f", spectral["sig_fx"],
                    spectral["sig_fy"], spectral["rf"], spectral["feq"], spectral["f95"], spectral["spr"])
    return spectral


def _full_blocks(dev_stack, groups, saturation_value, eps, keep_maps: bool):
    out: dict = {}
    if "grain" in groups and min(dev_stack.shape[1:]) < 128:
        raise ValueError("image too small for speckle grain metrics (min dimension < 128).")
    if blocks.fused_available(dev_stack) and (groups & {"grain", "bandwidth"}):
        # one fused pass: a single forward FFT per frame feeds grain (autocorrelation) and bandwidth (spectral sums)
        fb = blocks.FusedBlocks(dev_stack, saturation_value=saturation_value, eps=eps, keep_map=keep_maps and "grain" in groups,
                                want_tails="amplitude" in groups, want_spectral="bandwidth" in groups)
        if "amplitude" in groups:
            out["amplitude"] = fb.amplitude()
        if "grain" in groups:
            g = fb.grain()
            if keep_maps:
                n = int(fb.ac.shape[-1])
                g["autocorr"], g["xlag"], g["ylag"] = fb.ac, lag_axis(n, 1.0), lag_axis(n, 1.0)
            out["grain"] = g
        if "stats" in groups:
            out["stats"] = fb.moments()
        if "bandwidth" in groups:
            out["bandwidth"] = fb.bandwidth()
        return out
    table = engine.frame_reductions(dev_stack, saturation_value=saturation_value, eps=eps)
    if "amplitude" in groups:
        out["amplitude"] = blocks.amplitude_block(dev_stack, table)
    if "grain" in groups:
        if keep_maps:
            g, ac = blocks.grain_block(dev_stack, table=table, return_map=True)
            n = int(ac.shape[-1])
            g = dict(g)
            g["autocorr"], g["xlag"], g["ylag"] = ac, lag_axis(n, 1.0), lag_axis(n, 1.0)
            out["grain"] = g
        else:
            out["grain"] = blocks.grain_block(dev_stack, table=table)
    if "stats" in groups:
        out["stats"] = blocks.moments_block(table, saturation_value)
    if "bandwidth" in groups:
        out["bandwidth"] = blocks.bandwidth_block(dev_stack, table=table)
    return out


def speckle_stats(image, *, metrics="all", tiles: bool = True, display_origin: str = "lower",
                  saturation_value: float | None = 65535.0, eps: float = 1e-6, verbose: bool = True) -> dict:
    """Speckle metrics of one frame; result schema of the reference ({"meta", "full"[, "tiles"]})."""
    if not isinstance(image, np.ndarray):
        raise TypeError("speckle_stats expects a numpy.ndarray")
    if image.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={image.ndim}")
    h, w = image.shape
    groups = normalize_groups(metrics, all_groups=_ALL_SPECKLE_GROUPS, context="speckles", param_name="metrics")
    # display_origin="lower" analyses the row-flipped frame (common.py:70-71): flipped on the device (a 17 MB host copy of a
    # 2048^2 frame costs more than every kernel of this call together)
    dev = engine.as_stack(np.ascontiguousarray(image))
    if normalize_display_origin(display_origin) == "lower":
        dev = dev.flip(1).contiguous()
    full = _full_blocks(dev, groups, saturation_value, eps, keep_maps=True)
    out = {"meta": {"kind": "speckles", "display_origin": display_origin, "input_shape": (int(h), int(w)),
                    "requested_groups": sorted(groups), "units": _SPECKLE_UNITS, "tile_mode": "off"}, "full": {}}
    for grp in ("amplitude", "grain", "stats", "bandwidth"):
        if grp not in full:
            continue
        blk = {}
        for k, v in full[grp].items():
            if k == "autocorr":
                blk[k] = _map_to_host_f64(v[:1])[0]
            elif k in ("xlag", "ylag"):
                blk[k] = v
            else:
                blk[k] = float(v[0])
        out["full"][grp] = blk
    mode, tile_shape_px = choose_tiling_mode(h, w, tiles=tiles)
    if mode != "off":
        out["meta"].update(tiles_meta(h, w, tile_mode=mode, tile_shape_px=tile_shape_px))
        tl = _tiles(dev, mode, groups, saturation_value, eps)
        if tl:
            out["tiles"] = {g: {k: {"mean": v["mean"][0], "std": v["std"][0]} for k, v in f.items()} for g, f in tl.items()}
    if verbose:
        logger.info("\nspeckle stats for a (h x w: %.0f x %.0f) image: %s", h, w, sorted(groups))
    return out


def _tiles(dev_oriented, mode, groups, saturation_value, eps) -> dict:
    """{group: {field: {"mean": (T, 3, 3), "std": (T, 3, 3)}}} of a display-oriented device stack (speckles.py:192-250)."""
    res = tiled_blocks(dev_oriented, tile_mode=mode,
                       block_fn=lambda tl: _full_blocks(tl, groups, saturation_value, eps, keep_maps=False))
    return {g: res[g] for g in ("amplitude", "grain", "stats", "bandwidth") if g in res}


def _map_to_host_f64(maps, chunk: int = 8) -> np.ndarray:
    """(T, N, N) float32 device maps -> float64 numpy (the reference returns float64 autocorrelation maps): widened on the
    device a few frames at a time and copied straight into the result array (no float32 host copy, no host-side cast)."""
    import torch
    T = int(maps.shape[0])
    out = np.empty(tuple(maps.shape), dtype=np.float64)
    dst = torch.from_numpy(out)
    for a in range(0, T, chunk):
        dst[a:a + chunk].copy_(maps[a:a + chunk].double())
    return out


def _empty_like_block(block):
    """The same nested block with zero frames (a rank that owns no frame still takes part in the gathers)."""
    if isinstance(block, dict):
        return {k: _empty_like_block(v) for k, v in block.items()}
    return block[:0]


def _odd_size(n: float, min_size: int = 3) -> int:
    size = max(int(math.ceil(n)), min_size)
    return size + 1 if size % 2 == 0 else size


def _roi_grid_3x3(shape, roi, step):
    """3x3 grid of centred odd ROIs (geometry/roi.py:109-172): row-major NW..SE, raises when out of bounds."""
    H, W = shape
    cy, cx = H // 2, W // 2
    half = roi // 2
    grid = []
    for dy in (-step, 0, step):
        row = []
        for dx in (-step, 0, step):
            y0, x0 = cy + dy - half, cx + dx - half
            if y0 < 0 or y0 + roi > H or x0 < 0 or x0 + roi > W:
                raise ValueError("ROI exceeds image bounds.")
            row.append((slice(y0, y0 + roi), slice(x0, x0 + roi)))
        grid.append(row)
    return grid


def speckle_stack_stats(stack, *, metrics="all", tiles: bool = True, display_origin: str = "lower",
                        roi_grain_factor: float = 3.0, roi_step_factor: float = 0.5, tracking_method: str = "template",
                        tracking_backend: str = "skimage", subpixel: bool = True,
                        saturation_value: float | None = 65535.0, eps: float = 1e-6, verbose: bool = True,
                        parallel: bool = True, n_jobs: int | None = None, keep_autocorr: bool = True,
                        sharded: bool | None = None) -> dict:
    """Per-frame speckle metrics of a (T, H, W) stack plus 3x3-ROI translation tracking (abs / inc).

    Under torch.distributed (one process per GPU, torchrun; sharded=None picks it up, False ignores it) every rank
    passes the same host stack and works on its contiguous frame range only -- widened by the one-frame halo the
    incremental tracker needs (frame t against frame t-1, metrics/speckles.py:349,373) -- and every per-frame leaf is
    all-gathered: all ranks return the full result, identical to the single-GPU one.

    Trackers: "template" (the reference's default; both of its backend names are served by the same normalised
    cross-correlation kernels, pinned against opencv) and "phase" / "internal". Like the reference, the result carries
    the per-frame (T, N, N) float64 autocorrelation stack in full["grain"] (32 MB per 2048^2 frame on the host, SURVEY.md
    section 7 "hard parts"); keep_autocorr=False, the one keyword the reference does not have, leaves it out.
    """
    if not isinstance(stack, np.ndarray):
        raise TypeError("speckle_stack_stats expects a numpy.ndarray")
    if stack.ndim != 3:
        raise ValueError(f"stack must be a 3D array with shape (T, H, W); got ndim={stack.ndim}")
    T, H, W = (int(v) for v in stack.shape)
    if T < 1:
        raise ValueError("stack must contain at least one frame.")
    normalize_display_origin(display_origin)
    groups = normalize_groups(metrics, all_groups=_ALL_SPECKLE_GROUPS, context="speckles", param_name="metrics")
    method = str(tracking_method).strip().lower()
    if method not in ("phase", "template"):
        raise ValueError(f"Unsupported tracking method: {tracking_method!r}. Supported: phase, template")
    if method == "phase" and tracking_backend != "internal":
        raise B4DUnsupported("phase tracking on the B200 path implements tracking_backend='internal' only")
    if method == "template" and tracking_backend not in ("opencv", "skimage"):
        raise ValueError("backend must be 'opencv' or 'skimage'.")
    rank, world = _par.dist_info() if sharded is not False else (0, 1)
    lo, hi = _par.frame_range(T, rank, world)
    hlo, _ = _par.inc_halo_range(T, rank, world)
    Tl = hi - lo
    dev_h = engine.as_stack(stack[hlo:max(hi, hlo + 1)])          # this rank's frames + the halo frame lo - 1
    dev = dev_h[lo - hlo:lo - hlo + Tl]
    dev0 = dev_h[:1] if hlo == 0 else engine.as_stack(stack[:1])  # frame 0: ROI size, absolute template (every rank)
    # the reference runs speckle_stats(frame, display_origin=...) per frame, i.e. on the row-flipped frame for "lower":
    # scalars do not see the flip, the returned autocorrelation maps do
    flip_maps = keep_autocorr and normalize_display_origin(display_origin) == "lower"
    if Tl > 0:
        full = _full_blocks(dev.flip(1).contiguous() if flip_maps else dev, groups, saturation_value, eps, keep_maps=keep_autocorr)
    else:
        full = {k: _empty_like_block(v) for k, v in _full_blocks(dev0, groups, saturation_value, eps, keep_maps=keep_autocorr).items()}
    if "grain" in full and keep_autocorr:
        full["grain"]["autocorr"] = _map_to_host_f64(full["grain"]["autocorr"])
        n = full["grain"]["autocorr"].shape[-1]
        full["grain"]["xlag"] = np.tile(full["grain"]["xlag"], (Tl, 1))
        full["grain"]["ylag"] = np.tile(full["grain"]["ylag"], (Tl, 1))

    g0 = blocks.grain_block(dev0)
    l = float(np.nanmax([g0["lx"][0], g0["ly"][0], g0["leq"][0]]))
    if not np.isfinite(l) or l <= 0:
        raise ValueError("Could not infer a valid grain size from frame 0 (lx/ly/leq).")
    roi = _odd_size(int(math.ceil(roi_grain_factor * l)))
    step = int(max(1, round(roi_step_factor * roi)))
    grid = _roi_grid_3x3((H, W), roi, step)

    dx_abs = np.empty((Tl, 3, 3), np.float32)
    dy_abs = np.empty((Tl, 3, 3), np.float32)
    dx_inc = np.empty((Tl, 3, 3), np.float32)
    dy_inc = np.empty((Tl, 3, 3), np.float32)
    prev = np.maximum(np.arange(lo, hi) - 1, 0) - hlo              # index of frame t-1 (frame 0 for t = 0) inside dev_h
    for iy in range(3):
        for ix in range(3):
            if Tl == 0:
                break
            sy, sx = grid[iy][ix]
            if method == "template":
                # the reference's default tracker: normalised cross-correlation of the ROI against the full frame; absolute
                # = ROI of frame 0 against every frame, incremental = ROI of frame t-1 against frame t, both batched
                centre = ((sy.start + sy.stop - 1) / 2.0, (sx.start + sx.stop - 1) / 2.0)
                tab = engine.template_match(dev0[0, sy, sx], dev, ref_center_yx=centre, subpixel=subpixel, eps=1e-9)
                dy_abs[:, iy, ix], dx_abs[:, iy, ix] = tab[:, 0], tab[:, 1]
                tab = engine.template_match(dev_h[prev][:, sy, sx].contiguous(), dev, ref_center_yx=centre, subpixel=subpixel, eps=1e-9)
                dy_inc[:, iy, ix], dx_inc[:, iy, ix] = tab[:, 0], tab[:, 1]
                continue
            # absolute: one reference (ROI of frame 0) against the whole stack
            tr = engine.PhaseTracker(dev0[0, sy, sx].contiguous(), (H, W), y0=sy.start, x0=sx.start, eps=1e-9)
            tab = tr.track(dev, subpixel=subpixel)
            tr.close()
            dy_abs[:, iy, ix], dx_abs[:, iy, ix] = tab[:, 0], tab[:, 1]
            # incremental: the ROI of frame t-1 (frame 0 for t = 0) against frame t
            for t in range(Tl):
                tri = engine.PhaseTracker(dev_h[int(prev[t]), sy, sx].contiguous(), (H, W), y0=sy.start, x0=sx.start, eps=1e-9)
                r = tri.track(dev[t:t + 1], subpixel=subpixel)[0]
                tri.close()
                dy_inc[t, iy, ix], dx_inc[t, iy, ix] = r[0], r[1]

    def summarise(dx, dy):
        r = np.sqrt(dx ** 2 + dy ** 2)
        f = lambda a, fn: fn(a, axis=(1, 2)).astype(np.float32)
        return {"dx": f(dx, np.nanmean), "dy": f(dy, np.nanmean), "r": f(r, np.nanmean),
                "std_dx": f(dx, np.nanstd), "std_dy": f(dy, np.nanstd), "std_r": f(r, np.nanstd)}

    temporal = {"abs": summarise(dx_abs, dy_abs), "inc": summarise(dx_inc, dy_inc), "qc": {"roi_grid_shape": (3, 3)}}
    serial = (not parallel) or (n_jobs is not None and int(n_jobs) <= 1)
    meta = {
        "kind": "speckle_stack_stats", "input_shape": (H, W), "stack_shape": (T, H, W), "n_frames": T,
        "display_origin": display_origin, "units": _SPECKLE_UNITS,
        "grain0": {k: float(g0[k][0]) for k in ("lx", "ly", "leq", "r")},
        "tracking": {"method": str(tracking_method), "backend": str(tracking_backend), "subpixel": bool(subpixel),
                     "peak_mode": "abs", "search_area": "full_frame",
                     "normalization": {"template": "zscore_local", "search": "zscore_global"},
                     "roi_grain_factor": float(roi_grain_factor), "roi_size_yx": (roi, roi),
                     "roi_step_factor": float(roi_step_factor), "roi_step_yx": (step, step),
                     "roi_labels": np.array([["NW", "N", "NE"], ["W", "C", "E"], ["SW", "S", "SE"]], dtype=object),
                     "roi_order": "row-major"},
        "parallel": {"enabled": bool(not serial), "joblib_verbose": 0},
    }
    out_full = {grp: full[grp] for grp in ("amplitude", "grain", "stats", "bandwidth") if grp in full}
    out = {"meta": meta, "full": out_full, "temporal": temporal}
    # tiles: the reference evaluates speckle_stats(frame, tiles=...) per frame, i.e. on the display-oriented frame
    mode, _ = choose_tiling_mode(H, W, tiles=tiles)
    if mode != "off":
        src = dev if Tl > 0 else dev0
        oriented = src.flip(1) if normalize_display_origin(display_origin) == "lower" else src
        tl = _tiles(oriented, mode, groups, saturation_value, eps)
        if Tl == 0:
            tl = {g: _empty_like_block(f) for g, f in tl.items()}
        if tl:
            out["tiles"] = tl
    if world > 1:
        # every per-frame leaf back to the full stack length, in rank order; meta is identical on every rank already
        qc = out["temporal"].pop("qc")
        for part in ("full", "temporal", "tiles"):
            if part in out:
                out[part] = _par.gather_tree(out[part], T)
        out["temporal"]["qc"] = qc
        meta["sharding"] = {"world_size": world, "frame_range_of_rank": [list(_par.frame_range(T, r, world)) for r in range(world)]}
    if verbose:
        logger.info("> speckle_stack_stats | frames=%d | roi=%dx%d | step=%d | device=cuda", T, roi, roi, step)
    return out
