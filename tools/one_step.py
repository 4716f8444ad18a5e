#!/usr/bin/env python
"""One energy+forces step of the benchmark workload inside a cudaProfilerStart/Stop range (for
`ncu --profile-from-start off ...`), after two warm-up steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import DEFAULT_HYPERS, seed_all  # noqa: E402
from metatrain_b200 import B200PETBackend, evaluate  # noqa: E402
from metatrain_b200.systems import make_batch, replicate, water_384  # noqa: E402

dev = "cuda:0"
seed_all(0)
be = B200PETBackend(dict(DEFAULT_HYPERS), [1, 8], precision="bf16x3")
be.add_output("energy", {"energy___0": [1]})
be = be.to(dev).eval()
be.emit_nef = False
batch = {k: v.to(dev) for k, v in make_batch([replicate(water_384(), (3, 3, 3))], 4.5).items()}
for _ in range(2):
    evaluate(be, **batch, target="energy")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = evaluate(be, **batch, target="energy")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("energy", float(out["energies"]))
