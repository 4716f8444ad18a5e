#!/usr/bin/env python
"""How much of a step is GPU idle time?  Sum of per-call device durations (CUDA events around every
C-ABI call) vs the device time of the whole step, plus the host time spent issuing one step."""
import contextlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import DEFAULT_HYPERS, seed_all  # noqa: E402
from metatrain_b200 import B200PETBackend, evaluate, lib  # noqa: E402
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

# whole step, device time and host issue time
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    evaluate(be, **batch, target="energy")
e1.record()
host_ms = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 5

records = []


@contextlib.contextmanager
def hook(name, args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    yield
    b.record()
    records.append((name, a, b))


lib.profile_hook = hook
evaluate(be, **batch, target="energy")
lib.profile_hook = None
torch.cuda.synchronize()
tot = {}
for name, a, b in records:
    t, n = tot.get(name, (0.0, 0))
    tot[name] = (t + a.elapsed_time(b), n + 1)
kernel_ms = sum(t for t, _ in tot.values())
print(f"step (device events, 5 steps back to back): {step_ms:.2f} ms; host time to issue one step: {host_ms:.2f} ms")
print(f"sum of C-ABI call durations in one step: {kernel_ms:.2f} ms over {len(records)} calls  -> "
      f"{step_ms - kernel_ms:.2f} ms per step outside libpetb200 kernels")
for name, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"  {name:22s} {n:4d} calls {t:8.3f} ms")
