#!/usr/bin/env python3
"""Median per-stage kernel time from an ncu gpu__time_duration launch list (csv)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split("(")[0].replace("void ", "")].append(float(r[-1]) / 1e3)
tot = 0.0
for k, v in agg.items():
    v2 = sorted(v)
    med = v2[len(v2) // 2]
    if k.startswith("wf_") and "init" not in k:
        tot += med
    print("%-28s n %4d  median %8.1f us  p90 %8.1f us" % (k, len(v), med, v2[int(len(v2) * 0.9)]))
print("round (sum of medians) %.1f us" % tot)
