"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel and per-grid totals."""
import collections
import csv
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare shares, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3:10.1f} us {v[0]:5d} launches {100 * v[1] / tot:5.1f}%  {k}")
for pat in sys.argv[2:]:
    g = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if pat in r["Kernel Name"]:
            g[r["Grid Size"]][0] += 1
            g[r["Grid Size"]][1] += float(r["Metric Value"]) / 1e3
    print(f"-- {pat} by grid")
    for k, v in sorted(g.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:>16} x{v[0]:3d} {v[1]:8.1f} us  ({v[1] / v[0]:6.1f} us each)")
