#!/usr/bin/env python
"""cProfile of the host side of one energy+forces step (10k-atom box)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import DEFAULT_HYPERS, seed_all  # noqa: E402
from metatrain_b200 import B200PETBackend, evaluate  # noqa: E402
from metatrain_b200.systems import make_batch, replicate, water_384  # noqa: E402

dev = torch.device("cuda:0")
seed_all(0)
be = B200PETBackend(dict(DEFAULT_HYPERS), [1, 8], precision="bf16x3")
be.add_output("energy", {"energy___0": [1]})
be = be.to(dev).eval()
be.emit_nef = False
batch = {k: v.to(dev) for k, v in make_batch([replicate(water_384(), (3, 3, 3))], 4.5).items()}
for _ in range(3):
    evaluate(be, **batch, target="energy")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    evaluate(be, **batch, target="energy")
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
