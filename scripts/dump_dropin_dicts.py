#!/usr/bin/env python
"""GPU box: the drop-in's aggregator dicts on the inputs of oracle/schema.py -> gpurun_out/dropin_dicts.pkl.gz
(maps zeroed: the report does not read them). Copied to tests/golden/ and consumed by the container-tier test that
runs the reference's own logbook_report (report/markdown.py:37) on them."""
import gzip, os, pickle, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import barc4dip_b200 as dip
from oracle import schema as sc

out = {name: sc.strip_big_arrays(call()) for name, call in sc.schema_calls(dip.metrics).items()}
os.makedirs("gpurun_out", exist_ok=True)
with gzip.open("gpurun_out/dropin_dicts.pkl.gz", "wb") as fh:
    pickle.dump(out, fh, protocol=4)
print({k: len(sc.schema_tree(v)) for k, v in out.items()})
