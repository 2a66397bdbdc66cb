#!/usr/bin/env python
"""One-line digest of a bench.py JSON line: ms/step, roofline fraction, stage times."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(path, "ms/step", round(d["ms_per_step"], 2), "value", round(d["value"]), "frac", round(r.get("frac", 0), 3),
          r.get("kernel"), r.get("stage_ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
