#!/usr/bin/env python
"""Per-kernel averages of an ncu --set full report as one markdown table (the head of
profiles/r2_ncu_fused_kernels.md); the per-launch blocks come from tools/ncu_summary.py.

    python tools/ncu_table.py prof.ncu-rep
"""
import collections
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "avg us", 1e-3), ("dram__bytes_read.sum", "DRAM read MB", None),
        ("dram__bytes_write.sum", "DRAM write MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("petb200::<unnamed>::", "").replace("void ", "")
        vals = []
        for key, _, scale in COLS:
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]]
            vals.append(v * UNIT.get(u, 1.0) if (scale is None or key.startswith("gpu__time")) else v)
        a = acc.setdefault(name, [0, [0.0] * len(COLS)])
        a[0] += 1
        a[1] = [x + y for x, y in zip(a[1], vals)]
    print("| kernel | launches | " + " | ".join(c[1] for c in COLS) + " | DRAM GB/s |")
    print("|---|---:|" + "---:|" * (len(COLS) + 1))
    for name, (n, tot) in acc.items():
        avg = [t / n for t in tot]
        gbs = (avg[1] + avg[2]) / avg[0] * 1e3 if avg[0] else 0.0
        print(f"| `{name}` | {n} | " + " | ".join(f"{v:.1f}" if i else f"{v:.1f}" for i, v in enumerate(avg)) + f" | {gbs:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1])
