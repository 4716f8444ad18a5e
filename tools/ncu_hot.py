#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report (first kernel, or -k index).
    python tools/ncu_hot.py report.ncu-rep [kernel_index] [top_n]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
kernels, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
k = kernels[which]
hdr = k["hdr"]
i_src, i_s = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
rows = k["rows"]
total = sum(int(r[i_s] or 0) for r in rows)
print(k["name"][:110], " total samples", total, " instructions", len(rows))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][i_s] or 0))[:top]
for i in sorted(order):
    r = rows[i]
    prev = rows[i - 1][i_src].strip()[:50] if i else ""
    print(f"{i:5d} {int(r[i_s]):7d} {100.0 * int(r[i_s]) / total:5.1f}%  {r[i_src].strip()[:70]:70s} | prev: {prev}")
