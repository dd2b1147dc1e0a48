#!/usr/bin/env python
"""Aggregate an ncu --metrics gpu__time_duration.sum --csv launch list by kernel name."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"].replace("<unnamed>::", "")[:70]
    agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")) / 1000.0)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'mean us':>9s} {'sum us':>10s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:70s} {len(v):5d} {sum(v)/len(v):9.1f} {sum(v):10.1f} {100*sum(v)/tot:5.1f}%")
