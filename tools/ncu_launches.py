#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).
    python tools/ncu_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"# launch list of `{sys.argv[1]}` ({len(rows)} launches, {tot / 1e6:.3f} ms under ncu: cold caches, serialised)\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:100]}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")
