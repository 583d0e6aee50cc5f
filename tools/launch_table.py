#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name -> markdown table on stdout.
Usage: python tools/launch_table.py gpurun_out/launches_<tag>.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    k = r[iK].split("(")[0][:70]
    v = float(r[iV].replace(",", ""))
    v = {"us": v / 1e3, "ns": v / 1e6, "s": v * 1e3, "ms": v}.get(r[iU], v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | ms / launch | share |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / tot:.1f}% |")
print(f"\ntotal {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches")
