#!/usr/bin/env python
"""Latency of one energy+forces step on small systems (host-overhead bound): Si-64 and water-384."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import golden_inputs, load_golden, seed_all  # noqa: E402
from metatrain_b200 import B200PETBackend, evaluate  # noqa: E402

for case in ("si_64", "water_384"):
    g = load_golden(case)
    seed_all(0)
    be = B200PETBackend(g["hypers"], g["atomic_types"], precision="bf16x3")
    be.add_output(g["target"], {g["target"] + "___0": [1]})
    be = be.to("cuda:0").eval()
    be.emit_nef = False
    inp = golden_inputs(g, "cuda:0")
    for _ in range(5):
        evaluate(be, **inp, target=g["target"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        out = evaluate(be, **inp, target=g["target"])
        out["dE_dpos"].cpu()
    dt = (time.perf_counter() - t0) / n
    na = inp["positions"].shape[0]
    print(f"{case}: {na} atoms, {dt * 1e3:.2f} ms per step (host-synchronous) = {na / dt:,.0f} atom-steps/s")
    # the same step through md.GraphedEvaluator: device Verlet list + one CUDA graph replay per step
    from metatrain_b200 import GraphedEvaluator  # noqa: E402
    md = GraphedEvaluator(be, inp["species"], inp["cells"][0], periodic=True, skin=0.3, target=g["target"])
    pos = inp["positions"].clone()
    for _ in range(3):
        md(pos)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        pos = pos + 1e-3 * torch.randn_like(pos)
        md(pos)["dE_dpos"].cpu()
    dt = (time.perf_counter() - t0) / n
    print(f"{case}: CUDA-graph MD step {dt * 1e3:.2f} ms = {na / dt:,.0f} atom-steps/s "
          f"({md.n_captures} captures, {md.n_replays} replays, list builds {md.verlet.n_builds})")
