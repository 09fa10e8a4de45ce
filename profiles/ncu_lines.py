#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: ncu_lines.py <ncu --page source --csv output> <lib.so> <kernel substring> [top N]
Disassembles the kernel from the shared library with `nvdisasm -g` (line info from -lineinfo) and joins
it with the ncu SASS page by instruction order."""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

src_csv, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if kname in out:
        lines = out.splitlines()
        break
assert lines, "kernel not found"
# walk the function: remember current source line for every instruction
cur, in_fn, insn_line = None, False, []
for ln in lines:
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        in_fn = kname in ln
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        insn_line.append(cur)
rows = list(csv.reader(open(src_csv)))
H = rows[1]
ie, ws = H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        body.append((int(r[ie]), int(r[ws])))
    except (ValueError, IndexError):
        pass
n = min(len(body), len(insn_line))
agg = {}
for (ins, st), loc in zip(body[:n], insn_line[:n]):
    a = agg.setdefault(loc, [0, 0])
    a[0] += ins; a[1] += st
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print("SASS instructions: ncu %d, nvdisasm %d; warp-instructions %d" % (len(body), len(insn_line), ti))
srcs = {}
for loc, (ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    if loc:
        for root in (os.path.dirname(os.path.abspath(lib)) + "/csrc", "."):
            pth = os.path.join(root, loc[0])
            if os.path.isfile(pth):
                srcs.setdefault(pth, open(pth).read().splitlines())
                if loc[1] - 1 < len(srcs[pth]):
                    text = srcs[pth][loc[1] - 1].strip()[:90]
                break
    print("%-24s inst %5.1f%%  stall %5.1f%%  %s" % ("%s:%d" % loc if loc else "?", 100.0 * ins / ti, 100.0 * st / ts, text))
