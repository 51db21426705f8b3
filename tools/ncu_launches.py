"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/ncu_launches.py launches.csv"""
import collections
import csv
import sys

lines = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} launches {v[0]:4d}  total {v[1]:10.3f} ms  mean {v[1] / v[0]:8.3f} ms  share {100 * v[1] / tot:5.1f}%")
