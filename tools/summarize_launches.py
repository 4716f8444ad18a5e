#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("petb200::<unnamed>::", "")
        name = re.sub(r"^void ", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# ncu launch list summary: {path}\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if v[1] / total < 0.0005:
            continue
        print(f"| `{k[:80]}` | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / total * 100:.1f}% | {v[1] / v[0]:.1f} |")
    print(f"\ntotal {total / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
