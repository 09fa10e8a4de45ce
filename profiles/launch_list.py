#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: mean duration per kernel (us), launches."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("rod::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
    acc.setdefault(name, []).append(v)
for name, vs in acc.items():
    if name.startswith("at::"):
        continue
    print("%-52s n=%3d  mean %7.1f us  min %7.1f  max %7.1f" % (name[:52], len(vs), sum(vs) / len(vs), min(vs), max(vs)))
