"""Result-dict schema of the aggregators against the reference's.

GPU tier: key tree, kinds, dtypes and shapes of the drop-in's dicts == those recorded from the real reference
(tests/golden/schema.json, oracle/make_golden.py --only-schema).
Container tier: the reference's own consumer, report.markdown.logbook_report (report/markdown.py:37), runs on the
drop-in's dicts (tests/golden/dropin_dicts.pkl.gz, produced on a B200 by scripts/dump_dropin_dicts.py) and renders
the same report skeleton as on the reference's own dicts. Skipped where /root/reference does not exist.
"""

import gzip
import json
import os
import pickle

import numpy as np
import pytest

from oracle import schema as sc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden():
    with open(os.path.join(GOLDEN, "schema.json")) as fh:
        return json.load(fh)


def _rows(tree):
    return {r[0]: tuple(r[1:3]) + (tuple(r[3]),) for r in tree}


# leaves the drop-in adds or types differently on purpose
ALLOWED_EXTRA = ()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["speckle_stats", "sharpness_stats", "speckle_stack_stats", "sharpness_stack_stats"])
def test_result_schema_equals_reference(name):
    import barc4dip_b200 as dip
    want = _rows(_golden()[name]["tree"])
    got = _rows(sc.schema_tree(sc.schema_calls(dip.metrics)[name]()))
    missing = sorted(set(want) - set(got))
    extra = sorted(p for p in set(got) - set(want) if not p.startswith(ALLOWED_EXTRA or ("\0",)))
    assert not missing, f"leaves of the reference's dict missing from the drop-in's: {missing}"
    assert not extra, f"leaves the reference does not have: {extra}"
    diff = {p: (got[p], want[p]) for p in want if got[p] != want[p]}
    assert not diff, f"kind / dtype / shape differ (drop-in, reference): {diff}"


@pytest.mark.parametrize("name", ["speckle_stats", "sharpness_stats"])
def test_reference_logbook_report_accepts_dropin_dicts(name):
    from oracle.load_reference import load_reference, reference_available
    if not reference_available():
        pytest.skip("needs /root/reference (container tier)")
    path = os.path.join(GOLDEN, "dropin_dicts.pkl.gz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/dropin_dicts.pkl.gz not recorded yet")
    import importlib
    load_reference()
    rep = importlib.import_module("barc4dip.report.markdown")
    with gzip.open(path, "rb") as fh:
        dicts = pickle.load(fh)
    g = _golden()[name]
    for complete in (False, True):
        text = rep.logbook_report(dicts[name], complete=complete, notes=False)
        assert sc.markdown_skeleton(text) == g["markdown"][f"complete={complete}"]


def test_reference_has_no_stack_report_kind():
    """The reference's report registry knows 'speckles' and 'sharpness' only: the stack dicts are consumed by its
    plotting layer, not by logbook_report (recorded so that nobody looks for a missing formatter here)."""
    g = _golden()
    assert "Unsupported report kind" in g["speckle_stack_stats"]["markdown_error"]
    assert "Unsupported report kind" in g["sharpness_stack_stats"]["markdown_error"]
