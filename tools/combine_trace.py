#!/usr/bin/env python
"""Timeline of CTA 0 of petb200_combine_fwd (clock64 stamps, cycles relative to the first event).
Build the library with -DPETB200_COMBINE_TRACE first:
    make -C metatrain_b200/csrc clean && make -C metatrain_b200/csrc -j8 NVCCEXTRA=-DPETB200_COMBINE_TRACE
Roles: 0 = GEMM1 issuer, 1/2 = epilogue groups, 3 = row producer (warp 0), 4 = GEMM2 issuer, 5 = store warp 0."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metatrain_b200 import lib  # noqa: E402
from metatrain_b200.lib import call, ptr  # noqa: E402

dev = "cuda:0"
E, d = 392040, 128
torch.manual_seed(0)
t = torch.randn(E, d, device=dev)
m = torch.randn(E, d, device=dev)
perm = torch.randperm(E, device=dev)
rev = torch.empty(E, dtype=torch.int32, device=dev)
half = E // 2
rev[perm[:half]] = perm[half:2 * half].int()
rev[perm[half:2 * half]] = perm[:half].int()
wa = torch.randn(256, 256, device=dev) / 16
wb = torch.randn(128, 256, device=dev) / 16
s_vec, b_fold, b_b = wa.sum(1).contiguous(), torch.randn(256, device=dev) * 0.1, torch.randn(128, device=dev) * 0.1
h = lib.load()
img = [torch.empty(h.petb200_combine_image_bytes(d, b), device=dev, dtype=torch.uint8) for b in (0, 1)]
call("combine_pack", ptr(wa), ptr(wb), d, ptr(img[0]), ptr(img[1]))
p1, st = torch.empty(E, 256, device=dev), torch.empty(E, 2, device=dev)


def run():
    call("combine_fwd", ptr(t), d, ptr(rev), ptr(img[0]), ptr(s_vec), ptr(b_fold), ptr(b_b), E, d, ptr(m), d,
         ptr(p1), ptr(st))


for _ in range(2):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    run()
b.record()
torch.cuda.synchronize()
print(f"combine_fwd: {a.elapsed_time(b) / 5 * 1e3:.1f} us per launch")
buf = torch.zeros(6 * 4 * 16 * 8, device=dev, dtype=torch.int64)
fn = h.petb200_debug_combine_trace
fn.argtypes = [ctypes.c_void_p]
if fn(buf.data_ptr()) != 0:
    sys.exit("library was not built with -DPETB200_COMBINE_TRACE")
run()
torch.cuda.synchronize()
fn(None)
tr = buf.cpu().view(6, 4, 16, 8)
t0 = int(tr[tr > 0].min())
names = {0: ["G1 wait acc1_empty", "acc1_empty ok", "W stage 0 ready", "W stage 1 ready", "G1 issued"],
         1: ["wait acc1_full", "acc1_full ok", "loaded", "computed", "a2_empty ok", "A2 stored"],
         3: ["own copies landed", "x_empty ok", "own parked", "rev landed", "rev parked"],
         4: ["G2 wait a2_full", "a2_full ok", "acc2_empty ok", "G2 issued"],
         5: ["wait acc2_full", "acc2_full ok", "store done"]}
names[2] = names[1]
names[5] = ["store: wait acc2_full", "store: acc2_full ok", "store done"]
events = []
for role in range(6):
    for tile in range(4):
        for c in range(16):
            for ev in range(8):
                v = int(tr[role, tile, c, ev])
                if v:
                    events.append((v - t0, role, tile + 2, c, names[role][ev]))
for ts, role, tile, c, nm in sorted(events):
    if tile in (3, 4):
        print(f"{ts:8d}  role {role}  tile {tile}  chunk {c:2d}  {nm}")
