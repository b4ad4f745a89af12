"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (dev tool)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = list(csv.reader(open(path)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = name.replace("climb::<unnamed>::", "").replace("void ", "")[:80]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
print(f"{'us':>10} {'share':>6} {'n':>4}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:4d}  {k}")
