"""Aggregate an ncu launch list (gpu__time_duration.sum csv): per-kernel totals and the slowest launches.
usage: python tools/agg_launches.py launches.csv [min_us]"""
import collections
import csv
import re
import sys

f = sys.argv[1]
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 400.0
rows = list(csv.reader(open(f)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.defaultdict(lambda: [0, 0.0])
slow = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    n = re.sub(r"\(.*", "", r[ki]).replace("void hm::", "").replace("<unnamed>::", "")
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[n][0] += 1
    agg[n][1] += v
    if v / 1e3 >= min_us:
        slow.append((int(r[0]), n, r[gi], v / 1e3))
tot = sum(v[1] for v in agg.values())
print("total %.3f ms over %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-50s %5d %10.3f ms %5.1f%%" % (n[:50], c, t / 1e6, 100 * t / tot))
print("--- launches >= %.0f us" % min_us)
for i, n, g, v in slow:
    print(i, n, g, "%.0f us" % v)
