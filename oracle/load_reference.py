"""
Import the real barc4dip hot-path sub-packages from /root/reference  --  TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by
the optional ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).

``import barc4dip`` itself fails here because the top-level ``__init__`` pulls in ``io``
(h5py) and ``plotting`` (matplotlib), which are not installed; the hot-path sub-packages
import fine once a stub parent package is registered.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src/barc4dip"


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def load_reference() -> types.SimpleNamespace:
    """Return a namespace with the reference's signal / metrics / preprocessing.normalize / maths / geometry."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found at {REFERENCE_SRC}")
    if "barc4dip" not in sys.modules:
        stub = types.ModuleType("barc4dip")
        stub.__path__ = [REFERENCE_SRC]
        sys.modules["barc4dip"] = stub
    mods = {}
    for name in ("signal", "metrics", "maths", "geometry", "preprocessing.normalize",
                 "metrics.common", "utils.range"):
        mods[name.replace(".", "_")] = importlib.import_module(f"barc4dip.{name}")
    return types.SimpleNamespace(**mods)
