#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the launch count, total and mean
duration and the share of the summed GPU time.  usage: python tools/launch_summary.py gpurun_out/r02_launches.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
agg = OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    key = (name[:70], r[ix["Grid Size"]], r[ix["Block Size"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
for (name, g, b), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:70s} grid={g:>14s} block={b:>14s} n={n:4d} total_us={us:10.1f} mean_us={us/n:9.1f} share={100*us/tot:5.1f}%")
