#!/usr/bin/env python
"""Print value / e2e / ms per step of a bench.py JSON line read from stdin (sweep helper)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print(tag, "value %.0f (%.1f ms)  e2e %.0f (%.1f ms)  groups %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
                                                                 d["config"]["context_groups"]), flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
