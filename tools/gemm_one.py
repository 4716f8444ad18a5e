#!/usr/bin/env python
"""Run one GEMM shape a few times (for ncu captures): gemm_one.py N K epi [M] [prec]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metatrain_b200 import engine
from metatrain_b200.lib import *
N, K, epi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
M = int(sys.argv[4]) if len(sys.argv) > 4 else 392040
prec = int(sys.argv[5]) if len(sys.argv) > 5 else PREC_BF16X3
dev = "cuda:0"
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1
n_out = N // 2 if epi == EPI_SWIGLU else (2 * N if epi == EPI_SWIGLU_BWD else N)
out = torch.empty(M, n_out, device=dev)
kw = {}
if epi == EPI_NONE: kw["residual"] = torch.randn(M, N, device=dev)
if epi in (EPI_SILU, EPI_SWIGLU): kw["aux_out"] = torch.empty(M, N, device=dev)
if epi == EPI_SWIGLU_BWD: kw["aux_in"] = torch.randn(M, 2 * N, device=dev)
if epi == EPI_MUL_DSILU: kw["aux_in"] = torch.randn(M, N, device=dev)
for _ in range(4):
    engine.gemm(a, w, out, epilogue=epi, precision=prec, **kw)
torch.cuda.synchronize()
