#!/usr/bin/env python
"""Time petb200_edge_head_fwd / _bwd and petb200_compress_fwd / _bwd alone on 392 040 edge rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metatrain_b200 import lib  # noqa: E402
from metatrain_b200.lib import call, ptr  # noqa: E402

dev = "cuda:0"
E, d, N = 392040, 128, 10368
torch.manual_seed(0)
m = torch.randn(E, d, device=dev)
w1, w2 = torch.randn(d, d, device=dev) / 11, torch.randn(d, d, device=dev) / 11
b1, b2, w_e = torch.randn(d, device=dev) * 0.1, torch.randn(d, device=dev) * 0.1, torch.randn(d, device=dev) / 11
fc = torch.rand(E, device=dev)
ctr = torch.sort(torch.randint(0, N, (E,))).values.int().to(dev)
d_atomic = torch.randn(N, 1, device=dev)
h = lib.load()
img = [torch.empty(h.petb200_chain_image_bytes(d), device=dev, dtype=torch.uint8) for _ in range(2)]
call("chain_pack", ptr(w1), ptr(w2), d, ptr(img[0]), ptr(img[1]))
tiles = -(-E // 128)
e1p, e2p, pe = torch.empty(tiles * 128, d, device=dev), torch.empty(E, d, device=dev), torch.empty(E, device=dev)
d_m, d_fc = torch.empty(E, d, device=dev), torch.zeros(E, device=dev)
vec, dist = torch.randn(E, 3, device=dev), torch.rand(E, device=dev) + 0.5
geo_w, table = torch.randn(d, 4, device=dev) * 0.3, torch.randn(2, d, device=dev) * 0.3
z = torch.randint(0, 2, (E,), device=dev, dtype=torch.int32)
t_out, d_vec, d_dist = torch.empty(E, d, device=dev), torch.zeros(E, 3, device=dev), torch.zeros(E, device=dev)
runs = {
    "edge_head_fwd": lambda: call("edge_head_fwd", ptr(m), d, ptr(img[0]), ptr(b1), ptr(b2), ptr(w_e), 0.1, E, d,
                                  ptr(e1p), ptr(e2p), ptr(pe)),
    "edge_head_bwd": lambda: call("edge_head_bwd", ptr(d_atomic), ptr(ctr), ptr(fc), ptr(e1p), ptr(e2p), ptr(pe),
                                  ptr(img[1]), ptr(w_e), E, d, ptr(d_m), d, ptr(d_fc)),
    "compress_fwd": lambda: call("compress_fwd", ptr(m), d, ptr(img[0]), ptr(b1), ptr(geo_w), ptr(table), ptr(z),
                                 ptr(vec), ptr(dist), ptr(b2), E, d, ptr(e1p), ptr(t_out), d),
    "compress_bwd": lambda: call("compress_bwd", ptr(m), d, ptr(e1p), ptr(img[1]), ptr(geo_w), E, d, ptr(d_m), d, 1,
                                 ptr(d_vec), ptr(d_dist)),
}
for name, fn in runs.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:16s} {a.elapsed_time(b) / 10 * 1e3:7.1f} us")
