#!/usr/bin/env python
"""Run + time attention fwd/bwd on a synthetic water-like topology (also used for ncu captures).
usage: attn_one.py [precision: 0 = fp32 CUDA-core, 1 = bf16x3 tensor-core] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metatrain_b200.lib import call, ptr  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
N = 10368
g = torch.Generator().manual_seed(0)
counts = torch.randint(26, 49, (N,), generator=g)
row_ptr = torch.zeros(N + 1, dtype=torch.int32)
row_ptr[1:] = torch.cumsum(counts, 0)
E = int(row_ptr[-1])
row_ptr = row_ptr.to(dev)
qkv = torch.randn(E + N, 384, device=dev)
fc = torch.rand(E, device=dev)
out = torch.empty(E + N, 128, device=dev)
lse = torch.empty(E + N, 8, device=dev)
dsum = torch.empty(E + N, 8, device=dev)
go = torch.randn(E + N, 128, device=dev)
dqkv = torch.empty_like(qkv)
dfc = torch.zeros(E, device=dev)


def fwd():
    call("attention_fwd", ptr(qkv), ptr(row_ptr), ptr(fc), N, E, 8, 16, 0.25, 48, prec, ptr(out), ptr(lse))


def bwd():
    call("attention_bwd", ptr(qkv), ptr(out), ptr(lse), ptr(go), ptr(row_ptr), ptr(fc), N, E, 8, 16, 0.25,
         48, prec, ptr(dqkv), ptr(dfc), ptr(dsum))


for name, fn in (("fwd", fwd), ("bwd", bwd)):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tok = E + N
    gb = tok * (2048 if name == "fwd" else 4096 + 64) / 1e9
    print(f"attention {name} prec={prec}: {ms * 1e3:8.1f} us   {gb / ms * 1e3:7.0f} GB/s algorithmic  ({tok} tokens)")
