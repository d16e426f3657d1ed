#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`): launches, total time and share per kernel."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = r[ki].split("(")[0][:42]
    tot[name] += us; cnt[name] += 1
s = sum(tot.values())
print("%-44s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-44s %8d %12.1f %6.1f%%" % (k, cnt[k], tot[k], 100 * tot[k] / s))
