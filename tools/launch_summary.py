"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, total, average, share.
usage: launch_summary.py launches.csv [skip_first_n_launches]"""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [ln for ln in open(path) if ln.startswith('"')]
rows = list(csv.DictReader(lines))[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("fishk::", "").replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us of kernel time (ncu: cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} n={v[0]:4d} total={v[1] / 1e3:9.1f} us  avg={v[1] / v[0] / 1e3:8.1f} us  share={v[1] / tot:6.1%}")
