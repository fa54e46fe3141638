"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python profiles/summarize_launches.py launches.csv [--list]
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Grid Size")
agg, total, n = collections.OrderedDict(), 0.0, 0
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)[:80]
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    if "--list" in sys.argv:
        print(f"{v:9.1f} us  {name:60s} grid {r[gi]}")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    total += v
    n += 1
print(f"launches {n}  total {total:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.1f} us {100 * t / total:5.1f}% {c:4d}  {k}")
