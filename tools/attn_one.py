#!/usr/bin/env python
"""Run attention fwd/bwd on a synthetic water-like topology (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metatrain_b200.lib import call, ptr
dev = "cuda:0"
N = 10368
g = torch.Generator().manual_seed(0)
counts = torch.randint(26, 49, (N,), generator=g)
row_ptr = torch.zeros(N + 1, dtype=torch.int32); row_ptr[1:] = torch.cumsum(counts, 0)
E = int(row_ptr[-1]); row_ptr = row_ptr.to(dev)
qkv = torch.randn(E + N, 384, device=dev); fc = torch.rand(E, device=dev)
out = torch.empty(E + N, 128, device=dev); lse = torch.empty(E + N, 8, device=dev)
dsum = torch.empty(E + N, 8, device=dev); go = torch.randn(E + N, 128, device=dev); dqkv = torch.empty_like(qkv); dfc = torch.zeros(E, device=dev)
for _ in range(3):
    call("attention_fwd", ptr(qkv), ptr(row_ptr), ptr(fc), N, E, 8, 16, 0.25, 48, ptr(out), ptr(lse))
    call("attention_bwd", ptr(qkv), ptr(out), ptr(lse), ptr(go), ptr(row_ptr), ptr(fc), N, E, 8, 16, 0.25, 48, ptr(dqkv), ptr(dfc), ptr(dsum))
torch.cuda.synchronize()
