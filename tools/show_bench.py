#!/usr/bin/env python3
"""Prints the interesting fields of a bench.py JSON line (development helper)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.2f %s  ms/step %.2f | e2e %.2f ms/step %.2f" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)))
for k in ("stage_ms", "roofline", "roofline_hbm", "proof", "cpu_baseline", "clocks"):
    if k in d:
        print(k, json.dumps(d[k]))
