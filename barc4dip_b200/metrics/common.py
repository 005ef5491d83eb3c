"""
Host-side glue shared by the metric aggregators (result schema, group selection, display origin).

Mirrors the behaviour of the reference's metrics/common.py for the parts the hot path needs:
apply_display_origin (:44-72), normalize_groups (:411-464), stack_time_series (:381-408).
The 3x3 / 9x9 tiling executor (:278-378) is a "next" row of SURVEY.md 8(f) and is not built:
asking for tiles raises instead of silently falling back to a CPU path.
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


def reject_tiles(tiles: bool, h: int, w: int, min_tile_px: int = 128):
    """tiles=True is served only when the reference itself would have switched tiling off."""
    if not tiles:
        return
    if (h // 3) >= min_tile_px and (w // 3) >= min_tile_px:
        from .._lib import B4DUnsupported
        raise B4DUnsupported(
            "tiles=True (3x3 / 9x9 tile grids, metrics/common.py:278-378) is not built on the B200 path yet "
            "(non power-of-two tile FFTs); pass tiles=False. No CPU fallback is taken.")
    import warnings
    warnings.warn(f"Image too small for tiling: shape=({h}, {w}), min_tile_px={min_tile_px}.", RuntimeWarning,
                  stacklevel=3)
