#!/usr/bin/env python
"""Top warp-stall locations of the kernels in an ncu report (needs --import-source on / -lineinfo).

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [n_lines]
"""
import csv
import io
import subprocess
import sys


def main(path, n=22):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = raw.split('"Kernel Name",')
    for blk in blocks[1:]:
        lines = blk.splitlines()
        print("##", lines[0][:110])
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        hdr, data = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
        idx = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[idx["# Samples"]] or 0) for r in data) or 1
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in reasons}
        print("stall mix: " + ", ".join(f"{h[6:]} {v / tot * 100:.0f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
            s = int(r[idx["# Samples"]])
            main_reason = max(reasons, key=lambda h: int(r[idx[h]] or 0))
            print(f"{s / tot * 100:5.1f}%  {main_reason[6:]:14s} {r[idx['Source']][:100]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
