"""
Schema of the aggregators' result dicts -- TEST INFRASTRUCTURE ONLY.

The reference's consumers (report/markdown.py:37 ``logbook_report``, plotting/stack.py) walk the nested dicts that
``speckle_stats`` / ``sharpness_stats`` and their stack variants return (metrics/speckles.py:157-166,440-476;
metrics/sharpness.py:172-181,371-388). ``schema_tree`` flattens such a dict into (path, kind, dtype, shape) rows;
``make_golden.py --only-schema`` records the rows of the REAL reference's outputs in tests/golden/schema.json and the
GPU tier checks the drop-in's outputs against them row by row.
"""

from __future__ import annotations

import numpy as np

from barc4dip_b200 import synth


def schema_inputs():
    """Seeded inputs: one 384^2 frame (tiles_3x3 with 128 px tiles) and a 5-frame 384^2 drifting stack."""
    frame = synth.speckle_frame(384, grain=6.0, seed=71)
    stack, _ = synth.tracking_stack(5, 384, grain=8.0, seed=72, step_sigma=0.8)
    return frame, stack


def schema_calls(met):
    """name -> zero-argument callable on `met` (the reference's metrics package or the drop-in's)."""
    frame, stack = schema_inputs()
    return {
        "speckle_stats": lambda: met.speckle_stats(frame, metrics="all", tiles=True, verbose=False),
        "sharpness_stats": lambda: met.sharpness_stats(frame, metrics="all", tiles=True, verbose=False),
        "speckle_stack_stats": lambda: met.speckle_stack_stats(stack, metrics="all", tiles=True, tracking_method="template",
                                                               tracking_backend="opencv", verbose=False),
        "sharpness_stack_stats": lambda: met.sharpness_stack_stats(stack, metrics="all", tiles=True, verbose=False),
    }


def schema_tree(node, path=""):
    """Flatten a result dict into sorted rows [path, kind, dtype, shape]."""
    rows = []
    if isinstance(node, dict):
        for k in node:
            rows += schema_tree(node[k], f"{path}/{k}")
        return rows
    if isinstance(node, np.ndarray):
        return [[path, "ndarray", str(node.dtype), list(node.shape)]]
    if isinstance(node, (tuple, list)):
        return [[path, type(node).__name__, "", [len(node)]]]
    if isinstance(node, (bool, np.bool_)):
        return [[path, "bool", "", []]]
    if isinstance(node, (float, np.floating)):
        return [[path, "float", "", []]]
    if isinstance(node, (int, np.integer)):
        return [[path, "int", "", []]]
    if isinstance(node, str):
        return [[path, "str", "", []]]
    if node is None:
        return [[path, "none", "", []]]
    return [[path, type(node).__name__, "", []]]


def strip_big_arrays(node, limit: int = 4096):
    """Copy of a result dict with every array above `limit` elements replaced by zeros of its dtype and shape (the
    report does not read maps; keeps the committed fixture small)."""
    if isinstance(node, dict):
        return {k: strip_big_arrays(v, limit) for k, v in node.items()}
    if isinstance(node, np.ndarray) and node.size > limit and node.dtype != object:
        return np.zeros(node.shape, node.dtype)
    return node


def markdown_skeleton(text: str) -> list[str]:
    """The report with every number masked: headings, labels and table structure only."""
    import re
    out = []
    for ln in text.splitlines():
        out.append(re.sub(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|nan|inf", "#", ln).rstrip())
    return out
