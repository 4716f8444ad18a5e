#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
    python tools/summarize_launches.py gpurun_out/launches.csv --json profiles/traffic.json

The JSON form (DRAM bytes per launch of every kernel) is what bench.py quotes as
`roofline.traffic`.
"""
import collections
import csv
import json
import re
import sys

SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, json_out=None):
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    tot = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("petb200::<unnamed>::", "")
        name = re.sub(r"^void ", "", name)
        v = float(row["Metric Value"].replace(",", "")) * SCALE.get(row["Metric Unit"], 1.0)
        tot[name][row["Metric Name"]] += v
        ids[name].add(row["ID"])
    if json_out:
        table = {}
        for k, v in tot.items():
            if "dram__bytes_read.sum" not in v:
                continue
            n = len(ids[k])
            table[k] = {"launches": n, "avg_us": v["gpu__time_duration.sum"] / n,
                        "bytes_per_launch": (v["dram__bytes_read.sum"] + v.get("dram__bytes_write.sum", 0.0)) / n}
        with open(json_out, "w") as fh:
            json.dump({"source": path, "bytes_per_launch": table}, fh, indent=1, sort_keys=True)
    total = sum(v["gpu__time_duration.sum"] for v in tot.values())
    has_dram = any("dram__bytes_read.sum" in v for v in tot.values())
    print(f"# ncu launch list summary: {path}\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n")
    cols = "| kernel | launches | total ms | share | avg us |" + (" DRAM read MB | DRAM write MB | DRAM GB/s |" if has_dram else "")
    print(cols)
    print("|---|---:|---:|---:|---:|" + ("---:|---:|---:|" if has_dram else ""))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = v["gpu__time_duration.sum"]
        if t / total < 0.0005:
            continue
        n = len(ids[k])
        line = f"| `{k[:80]}` | {n} | {t / 1e3:.3f} | {t / total * 100:.1f}% | {t / n:.1f} |"
        if has_dram:
            rd, wr = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0)
            line += f" {rd / 1e6:.0f} | {wr / 1e6:.0f} | {(rd + wr) / (t * 1e-6) / 1e9:.0f} |"
        print(line)
    print(f"\ntotal {total / 1e3:.2f} ms over {sum(len(i) for i in ids.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--json" else None)
