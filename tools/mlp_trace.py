#!/usr/bin/env python
"""Timeline of CTA 0 of the fused feed-forward forward kernel (clock64 stamps, cycles relative to
the first event).  Roles: 0 = MMA warp, 1/2 = epilogue groups, 3 = activation producer warp 0."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metatrain_b200 import lib  # noqa: E402
from metatrain_b200.lib import call, ptr  # noqa: E402

dev = "cuda:0"
M, d, F = 392040, 128, 256
torch.manual_seed(0)
x = torch.randn(M, d, device=dev)
w_in, b_in = torch.randn(2 * F, d, device=dev) * d ** -0.5, torch.randn(2 * F, device=dev) * 0.1
w_out, b_out = torch.randn(d, F, device=dev) * F ** -0.5, torch.randn(d, device=dev) * 0.1
h = lib.load()
img = [torch.empty(h.petb200_mlp_image_bytes(F, b), device=dev, dtype=torch.uint8) for b in (0, 1)]
call("mlp_pack", ptr(w_in), ptr(w_out), d, F, ptr(img[0]), ptr(img[1]))
y = torch.empty(M, d, device=dev)
for _ in range(2):
    call("mlp_fwd", ptr(x), d, ptr(img[0]), ptr(b_in), ptr(b_out), M, d, F, ptr(y), d)
torch.cuda.synchronize()
buf = torch.zeros(4 * 4 * 16 * 16, device=dev, dtype=torch.int64)
fn = h.petb200_debug_mlp_trace
fn.argtypes = [ctypes.c_void_p]
fn(buf.data_ptr())
if len(sys.argv) > 1 and sys.argv[1] == "bwd":
    dy = torch.randn(M, d, device=dev)
    call("mlp_bwd", ptr(x), d, ptr(dy), d, ptr(img[1]), ptr(b_in), M, d, F, ptr(y), d)
else:
    call("mlp_fwd", ptr(x), d, ptr(img[0]), ptr(b_in), ptr(b_out), M, d, F, ptr(y), d)
torch.cuda.synchronize()
fn(None)
t = buf.cpu().view(4, 4, 16, 16)
t0 = int(t[t > 0].min())
names = {0: ["G1 wait acc1_empty", "acc1_empty ok", "W k1 ready", "G2 wait a2_full", "a2_full ok", "W k0 ready",
             "k0 MMAs issued / ug issued", "k1 MMAs issued / G1 done", "G2 MMAs issued"],
         1: ["wait acc1_full", "acc1_full ok", "loaded", "computed", "A2 stored"],
         2: ["wait acc1_full", "acc1_full ok", "loaded", "computed", "A2 stored"],
         3: ["wait x_empty", "x_empty ok", "copies landed", "converted"]}
events = []
for role in range(4):
    for tile in range(4):
        for c in range(16):
            for ev in range(16):
                v = int(t[role, tile, c, ev])
                if v:
                    if role in (1, 2) and c == 15:
                        nm = ["wait acc2_full", "acc2_full ok", "final epilogue done"][ev]
                    else:
                        nm = names[role][ev]
                    events.append((v - t0, role, tile, c, nm))
for ts, role, tile, c, nm in sorted(events):
    if tile in (1, 2):
        print(f"{ts:8d}  role {role}  tile {tile}  chunk {c:2d}  {nm}")
