#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into one row per kernel.
Usage: python scripts/launch_summary.py gpurun_out/x_launches.csv "<command that was profiled>" > profiles/x_launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit, i_grid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[i_name].split("(")[0].replace("void ", "")
    v = float(r[i_val].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r[i_unit], 1)
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print("# ncu launch list, steady state: per-launch times are cold-cache and serialised -> compare SHARES, not absolutes")
if len(sys.argv) > 2:
    print("# command: " + sys.argv[2])
print("kernel,launches,total_us,avg_us,max_us,share")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.1f,%.1f,%.4f" % (n, a[0], a[1], a[1] / a[0], a[2], a[1] / tot))
