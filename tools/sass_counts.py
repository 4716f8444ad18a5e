#!/usr/bin/env python
"""Instruction-mnemonic counts per kernel from `cuobjdump -sass libpetb200.so`: the proof that the
hot kernels use tcgen05 (UTCHMMA / UTCBAR), tensor memory (LDTM / STTM), TMA (UTMALDG / UBLKCP) and
legacy warp MMA (HMMA) where DESIGN.md says they do.

    python tools/sass_counts.py > profiles/r2_sass_counts.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "metatrain_b200", "csrc", "libpetb200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "UTCATOMSWS", "HMMA",
         "LDGSTS", "SYNCS", "ELECT", "FFMA", "MUFU", "LDG", "STG", "ATOMG", "RED"]


def strip_params(name):
    """Drop the trailing (parameter list) of a demangled function name."""
    if not name.endswith(")"):
        return name
    depth = 0
    for k in range(len(name) - 1, -1, -1):
        depth += name[k] == ")"
        depth -= name[k] == "("
        if depth == 0:
            return name[:k]
    return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = list(counts)
    try:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except (OSError, subprocess.CalledProcessError):
        pass
    print("# SASS instruction counts per kernel (`cuobjdump -sass metatrain_b200/csrc/libpetb200.so`, sm_100a)\n")
    print("UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (tensor memory), "
          "UTMALDG / UTMASTG / UTMAREDG = cp.async.bulk.tensor (TMA tile load / store / reduce-add), UTMAPF = TMA L2 prefetch, UBLKCP = cp.async.bulk, "
          "HMMA = warp-level mma.sync, LDGSTS = cp.async, SYNCS = mbarrier ops.\n")
    cols = [w for w in WATCH if any(c[w] for c in counts.values())]
    print("| kernel | total | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for fn, c in sorted(counts.items(), key=lambda kv: -kv[1]["_total"]):
        name = demangle.get(fn, fn)
        name = re.sub(r"petb200::\(anonymous namespace\)::", "", name)
        name = strip_params(name).replace("void ", "").replace("petb200::<unnamed>::", "")
        name = name.replace("(int)", "").replace("(bool)", "")
        if not any(c[w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "HMMA")) and c["_total"] < 1500:
            continue
        print(f"| `{name[:70]}` | {c['_total']} | " + " | ".join(str(c[w]) for w in cols) + " |")
        tot.update(c)
    lib_tot = collections.Counter()
    for c in counts.values():
        lib_tot.update(c)
    print("\nWhole library: " + ", ".join(f"{w} {lib_tot[w]}" for w in cols) + f"; {len(counts)} kernels.")


if __name__ == "__main__":
    sys.exit(main())
