#!/usr/bin/env python
"""Micro-benchmark of petb200_gemm on the shapes of the PET step (GPU only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metatrain_b200 import engine
from metatrain_b200.lib import *

M = int(sys.argv[1]) if len(sys.argv) > 1 else 392040
dev = "cuda:0"
shapes = [(128, 128, EPI_NONE, "128x128 plain"), (128, 128, EPI_SILU, "128x128 silu+aux"),
          (384, 128, EPI_NONE, "qkv 384x128"), (512, 128, EPI_SWIGLU, "mlp-in swiglu 512x128"),
          (128, 256, EPI_NONE, "mlp-out 128x256"), (256, 256, EPI_SILU, "combine-a 256x256"),
          (256, 128, EPI_SWIGLU_BWD, "swiglu-bwd 256x128"), (128, 512, EPI_NONE, "mlp-in^T 128x512"),
          (128, 384, EPI_NONE, "qkv^T 128x384")]
for prec, pname in ((PREC_FP32, "fp32"), (PREC_BF16X3, "bf16x3"), (PREC_BF16, "bf16")):
    for N, K, epi, name in shapes:
        a = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.1
        n_out = N // 2 if epi == EPI_SWIGLU else (2 * N if epi == EPI_SWIGLU_BWD else N)
        out = torch.empty(M, n_out, device=dev)
        aux_cols = {EPI_SILU: N, EPI_SWIGLU: N, EPI_SWIGLU_BWD: 2 * N}.get(epi)
        aux = torch.randn(M, aux_cols, device=dev) if aux_cols else None
        kw = {}
        if epi in (EPI_SILU, EPI_SWIGLU):
            kw["aux_out"] = aux
        if epi == EPI_SWIGLU_BWD:
            kw["aux_in"] = aux
        res = torch.randn(M, N, device=dev) if epi == EPI_NONE else None
        if res is not None:
            kw["residual"] = res
        for _ in range(3):
            engine.gemm(a, w, out, epilogue=epi, precision=prec, **kw)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            engine.gemm(a, w, out, epilogue=epi, precision=prec, **kw)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        byt = 4.0 * M * (K + n_out + (aux_cols or 0) + (N if res is not None else 0))
        print(f"{pname:7s} {name:26s} {ms*1e3:8.1f} us  {2.0*M*N*K/ms*1e-9:7.1f} TFLOP/s  {byt/ms*1e-6:7.0f} GB/s (algorithmic)")
