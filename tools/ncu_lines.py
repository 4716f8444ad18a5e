#!/usr/bin/env python
"""Warp-stall samples of an ncu report aggregated per CUDA source line (needs -lineinfo and
--import-source on).

    python tools/ncu_lines.py prof.ncu-rep [n_lines] [file.cu:lo-hi]

With a file range, prints every line of that range with its share of the kernel's samples (to read the
time split of one warp role, e.g. the single-thread MMA issuer)."""
import collections
import csv
import subprocess
import sys


def main(path, n=28, span=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur_file = hdr = kernel = None
    agg = collections.defaultdict(lambda: [0, ""])
    order = []
    for ln in raw.splitlines():
        if ln.startswith('"File Path"'):
            cur_file = next(csv.reader([ln]))[1].split("/")[-1]
        elif ln.startswith('"Function Name"'):
            k = next(csv.reader([ln]))[1]
            if k != kernel:
                kernel = k
                order.append(len(order))
        elif ln.startswith('"Line No"'):
            hdr = next(csv.reader([ln]))
        elif hdr is not None:
            r = next(csv.reader([ln]))
            if len(r) == len(hdr) and r[0] != "":
                key = (order[-1], kernel, cur_file, int(r[0]))
                agg[key][0] += int(r[hdr.index("# Samples")] or 0)
                agg[key][1] = r[1]
    tot = collections.Counter()
    for (o, k, f, l), (s, _) in agg.items():
        tot[(o, k)] += s
    for (o, k), t in sorted(tot.items()):
        print(f"## {k[:100]}  ({t} samples)")
        items = [(s, f, l, src) for (oo, kk, f, l), (s, src) in agg.items() if oo == o]
        if span:
            fname, rng = span.split(":")
            lo, hi = (int(x) for x in rng.split("-"))
            for s, f, l, src in sorted(items, key=lambda x: x[2]):
                if f == fname and lo <= l <= hi and s:
                    print(f"{s / t * 100:5.2f}% {f}:{l:4d}  {src.strip()[:110]}")
        else:
            for s, f, l, src in sorted(items, reverse=True)[:n]:
                print(f"{s / t * 100:5.1f}% {f}:{l:4d}  {src.strip()[:110]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28, sys.argv[3] if len(sys.argv) > 3 else None)
