#!/usr/bin/env python
"""Evaluate the elongated 1x1x8 water box (3 072 atoms, several tiles per persistent CTA) with fused paths
switched off one at a time; meant to be run under compute-sanitizer, whose slowdown changes the timing of the
producer / consumer roles inside the persistent kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import apply_lora, load_long_box, seed_all  # noqa: E402
from metatrain_b200 import B200PETBackend, engine, evaluate  # noqa: E402


def make_backend(g, precision):
    seed_all(0)
    be = B200PETBackend(g["hypers"], g["atomic_types"], precision=precision)
    be.add_output(g["target"], {g["target"] + "___0": g["out_shape"]})
    apply_lora(be, g)
    return be.to("cuda:0").eval()


case = sys.argv[1] if len(sys.argv) > 1 else "water_long_1x1x8"
g = load_long_box(case)
batch = {k: v.to("cuda:0") for k, v in g["batch"].items()}


def run(tag, combine=True):
    be = make_backend(g, precision="bf16x3")
    be.emit_nef = False
    if not combine:
        evaluate(be, **batch, target=g["target"])
        for C in be._pw.combine:
            C["img"] = None
    errs = []
    for _ in range(3):
        out = evaluate(be, **batch, target=g["target"])
        f = out["dE_dpos"].cpu().numpy()
        errs.append(float(np.abs(f - g["ref64_dE_dpos"]).max()))
    print(f"{tag:40s} " + " ".join(f"{e:.2e}" for e in errs), flush=True)


run("default")
engine.USE_STAGE_SCHEDULE = False
run("python schedule")
engine.USE_FUSED_CHAINS = False
run("python schedule, no chain kernels")
run("python schedule, no chain, no combine", combine=False)
engine.USE_FUSED_CHAINS = True
run("python schedule, chain, no combine", combine=False)
