"""
Host-side glue shared by the metric aggregators (result schema, group selection, display origin).

Mirrors the behaviour of the reference's metrics/common.py for the parts the hot path needs:
apply_display_origin (:44-72), normalize_groups (:411-464), stack_time_series (:381-408).
The 3x3 / 9x9 tiling executor (split_edges :75-106, choose_tiling_mode :109-170, tiles_meta :173-217,
aggregate_subtiles_9x9_to_3x3 :248-275, tiled_scalar_fields :278-378) is mirrored by `tiled_blocks`: the (sub)tiles of
every frame are gathered on the device, grouped by shape, and each shape class goes through the same batched metric
blocks as full frames (non power-of-two tiles take the Bluestein FFT path, csrc/generic_dft.cuh).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np


def normalize_display_origin(display_origin: str) -> str:
    origin = str(display_origin).strip().lower()
    if origin not in ("upper", "lower"):
        raise ValueError("display_origin must be 'upper' or 'lower'.")
    return origin


def apply_display_origin(image: np.ndarray, *, display_origin: str) -> np.ndarray:
    """Row-flipped view for display_origin="lower" (the reference's default), identity for "upper"."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError(f"apply_display_origin expects a 2D array, got ndim={img.ndim}")
    return img[::-1, :] if normalize_display_origin(display_origin) == "lower" else img


def normalize_groups(groups, *, all_groups: set[str], context: str, param_name: str = "metrics") -> set[str]:
    """'all' | 'a' | 'a,b' | sequence of names -> validated set of group keys."""
    if isinstance(groups, str):
        keys = {g.strip() for g in groups.split(",")}
    elif isinstance(groups, Sequence):
        keys = set()
        for g in groups:
            if not isinstance(g, str):
                raise TypeError(f"{context}: {param_name} must be str or a sequence of str")
            keys.add(g.strip())
    else:
        raise TypeError(f"{context}: {param_name} must be str or a sequence of str")
    if "all" in keys:
        return set(all_groups)
    unknown = sorted(k for k in keys if k not in all_groups)
    if unknown:
        raise ValueError(f"{context}: unknown {param_name} group(s): {', '.join(unknown)}. "
                         f"Allowed: {', '.join(sorted(all_groups))}")
    return keys


def stack_time_series(values: list):
    """Stack per-frame results along a new leading time axis (dicts recursively, arrays, scalars)."""
    if not values:
        raise ValueError("No values provided for stacking.")
    first = values[0]
    if isinstance(first, dict):
        return {k: stack_time_series([v[k] for v in values]) for k in first}
    if isinstance(first, np.ndarray):
        return np.stack([np.asarray(v) for v in values], axis=0)
    if isinstance(first, (float, int, np.floating, np.integer, bool, np.bool_)):
        return np.asarray(values)
    return list(values)


TILE_GRID_SHAPE_3X3 = (3, 3)
TILE_ORDER = "row-major"
TILE_LABELS_3X3 = np.array([["NW", "N", "NE"], ["W", "C", "E"], ["SW", "S", "SE"]], dtype=object)
MIN_TILE_PX = 128


def split_edges(length: int, n_parts: int) -> list[tuple[int, int]]:
    """[0, length) in n_parts contiguous slices, edges at round(linspace) (metrics/common.py:75-106)."""
    if length < 1:
        raise ValueError("length must be >= 1.")
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1.")
    edges = np.linspace(0, length, n_parts + 1)
    out = []
    for i in range(n_parts):
        a = int(round(float(edges[i])))
        b = max(int(round(float(edges[i + 1]))), a + 1)
        out.append((a, b))
    out[-1] = (out[-1][0], length)
    return out


def choose_tiling_mode(h: int, w: int, *, tiles: bool = False, min_tile_px: int = MIN_TILE_PX):
    """("off" | "tiles_3x3" | "subtiles_9x9", evaluation tile shape or None), the reference's policy (:109-170)."""
    if h < 1 or w < 1:
        raise ValueError("Invalid image shape (h and w must be >= 1).")
    if min_tile_px < 1:
        raise ValueError("min_tile_px must be >= 1.")
    if not bool(tiles):
        return "off", None
    if (h // 9) >= min_tile_px and (w // 9) >= min_tile_px:
        return "subtiles_9x9", (h // 9, w // 9)
    if (h // 3) >= min_tile_px and (w // 3) >= min_tile_px:
        return "tiles_3x3", (h // 3, w // 3)
    import warnings
    warnings.warn(f"Image too small for tiling: shape=({h}, {w}), min_tile_px={min_tile_px}.", RuntimeWarning,
                  stacklevel=3)
    return "off", None


def tiles_meta(h: int, w: int, *, tile_mode: str, tile_shape_px=None) -> dict:
    meta = {"tile_mode": tile_mode}
    if tile_mode == "off":
        return meta
    if tile_shape_px is None:
        raise ValueError("tile_shape_px must be provided when tile_mode is not 'off'.")
    meta.update({"tile_grid_shape": TILE_GRID_SHAPE_3X3, "tile_labels": TILE_LABELS_3X3, "tile_order": TILE_ORDER,
                 "tile_shape_px": (int(tile_shape_px[0]), int(tile_shape_px[1])),
                 "used_subtiles": bool(tile_mode == "subtiles_9x9")})
    return meta


def tiled_blocks(dev_stack, *, tile_mode: str, block_fn, frames_per_chunk: int = 16) -> dict:
    """tiled_blocks_chunk over chunks of frames (the gathered tiles are a second copy of the chunk in HBM)."""
    T = int(dev_stack.shape[0])
    parts = [tiled_blocks_chunk(dev_stack[a:a + frames_per_chunk], tile_mode=tile_mode, block_fn=block_fn)
             for a in range(0, T, frames_per_chunk)]
    if len(parts) == 1:
        return parts[0]
    return {grp: {key: {s: np.concatenate([p[grp][key][s] for p in parts], axis=0) for s in ("mean", "std")}
                  for key in parts[0][grp]} for grp in parts[0]}


def tiled_blocks_chunk(dev_stack, *, tile_mode: str, block_fn) -> dict:
    """The reference's tiling executor for an HBM-resident, display-oriented (T, H, W) stack.

    block_fn(tiles) maps a device stack of equally shaped tiles (n, h, w) to {group: {field: (n,) array}} -- the same
    batched blocks the full frames go through. Returns {group: {field: {"mean": (T, 3, 3), "std": (T, 3, 3)}}}:
    tiles_3x3 evaluates the nine tiles directly (std = NaN), subtiles_9x9 evaluates 81 sub-tiles and aggregates each
    3 x 3 block with np.mean / np.std(ddof=0) (metrics/common.py:248-275).
    """
    import torch
    if tile_mode not in ("tiles_3x3", "subtiles_9x9"):
        raise ValueError("tile_mode must be 'tiles_3x3' or 'subtiles_9x9'.")
    T, h, w = (int(v) for v in dev_stack.shape)
    n = 3 if tile_mode == "tiles_3x3" else 9
    y_edges, x_edges = split_edges(h, n), split_edges(w, n)
    classes: dict = {}
    for r, (y0, y1) in enumerate(y_edges):
        for c, (x0, x1) in enumerate(x_edges):
            classes.setdefault((y1 - y0, x1 - x0), []).append((r, c, y0, x0))
    grids: dict = {}
    for (th, tw), members in classes.items():
        tiles = torch.stack([dev_stack[:, y0:y0 + th, x0:x0 + tw] for (_, _, y0, x0) in members], dim=1)
        res = block_fn(tiles.reshape(len(members) * T, th, tw).contiguous())      # index = t * len(members) + m
        for grp, fields in res.items():
            for key, vals in fields.items():
                g = grids.setdefault(grp, {}).setdefault(key, np.empty((T, n, n), dtype=float))
                vals = np.asarray(vals, dtype=float).reshape(T, len(members))
                for m, (r, c, _, _) in enumerate(members):
                    g[:, r, c] = vals[:, m]
    out: dict = {}
    for grp, fields in grids.items():
        out[grp] = {}
        for key, g in fields.items():
            if n == 3:
                out[grp][key] = {"mean": g, "std": np.full((T, 3, 3), np.nan)}
            else:
                blocks9 = g.reshape(T, 3, 3, 3, 3).transpose(0, 1, 3, 2, 4).reshape(T, 3, 3, 9)
                out[grp][key] = {"mean": blocks9.mean(axis=-1), "std": blocks9.std(axis=-1, ddof=0)}
    return out
