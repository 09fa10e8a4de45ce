#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top SASS lines by executed instructions.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python ncu_source_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:90]
        H = rows[i + 1]
        si, ie, ws = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
        body = []
        j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            if len(r) > max(ie, ws):
                try:
                    body.append((int(r[ie]), int(r[ws]), r[si]))
                except ValueError:
                    pass
            j += 1
        tot = sum(b[0] for b in body)
        samp = sum(b[1] for b in body) or 1
        print("== %s\n   SASS lines %d, warp-instructions executed %d" % (name, len(body), tot))
        for n, s, src in sorted(body, key=lambda x: -x[0])[:top]:
            print("%10d %5.1f%%  stall-samples %5.1f%%  %s" % (n, 100.0 * n / max(tot, 1), 100.0 * s / samp, src[:110]))
        i = j
        break   # first launch only
    i += 1
