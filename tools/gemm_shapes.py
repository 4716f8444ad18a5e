#!/usr/bin/env python
"""Time the dense contractions of one PET step in isolation (CUDA events, warm L2 for the weights):
node-path shapes (M = atoms) next to the edge-path shapes (M = edges)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metatrain_b200 import engine  # noqa: E402
from metatrain_b200.lib import (EPI_MUL_DSILU, EPI_NONE, EPI_SILU, EPI_SWIGLU, EPI_SWIGLU_BWD,  # noqa: E402
                                PREC_BF16X3)

dev = "cuda:0"
N_ATOMS, N_EDGES = 10368, 392040
SHAPES = [  # label, M, N(out rows of W), K, epilogue, residual
    ("node w_con      256->128", N_ATOMS, 128, 256, EPI_NONE, False),
    ("node w_o        128->128", N_ATOMS, 128, 128, EPI_NONE, False),
    ("node w_exp      128->256", N_ATOMS, 256, 128, EPI_NONE, True),
    ("node wc_in  256->2x512 swiglu", N_ATOMS, 1024, 256, EPI_SWIGLU, False),
    ("node wc_out     512->256", N_ATOMS, 256, 512, EPI_NONE, True),
    ("node wc_out_t 256->512 swiglu_bwd", N_ATOMS, 512, 256, EPI_SWIGLU_BWD, False),
    ("node wc_in_t   1024->256", N_ATOMS, 256, 1024, EPI_NONE, False),
    ("node w_exp_t    256->128", N_ATOMS, 128, 256, EPI_NONE, False),
    ("node w_con_t    128->256", N_ATOMS, 256, 128, EPI_NONE, True),
    ("edge w_o        128->128", N_EDGES, 128, 128, EPI_NONE, True),
    ("edge w_2        128->128", N_EDGES, 128, 128, EPI_NONE, False),
    ("edge comb w_a   256->256 silu", N_EDGES, 256, 256, EPI_SILU, False),
    ("edge comb w_b   256->128", N_EDGES, 128, 256, EPI_NONE, True),
    ("edge head       128->128 silu", N_EDGES, 128, 128, EPI_SILU, False),
    ("edge qkv_t      384->128", N_EDGES, 128, 384, EPI_NONE, False),
    ("edge dsilu      128->128", N_EDGES, 128, 128, EPI_MUL_DSILU, False),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
print(f"{'shape':38s} {'us':>8s} {'GB/s':>8s} {'TFLOP/s':>8s}")
for label, M, N, K, epi, res in SHAPES:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    n_out = N // 2 if epi == EPI_SWIGLU else (2 * N if epi == EPI_SWIGLU_BWD else N)
    out = torch.empty(M, n_out, device=dev)
    kw = {}
    nbytes = 4 * M * (K + n_out)
    if res:
        kw["residual"] = torch.randn(M, n_out, device=dev)
        nbytes += 4 * M * n_out
    if epi in (EPI_SILU, EPI_SWIGLU):
        kw["aux_out"] = torch.empty(M, N, device=dev)
        nbytes += 4 * M * N
    if epi == EPI_SWIGLU_BWD:
        kw["aux_in"] = torch.randn(M, 2 * N, device=dev)
        nbytes += 4 * M * 2 * N
    if epi == EPI_MUL_DSILU:
        kw["aux_in"] = torch.randn(M, N, device=dev)
        nbytes += 4 * M * N
    for _ in range(3):
        engine.gemm(a, w, out, epilogue=epi, precision=PREC_BF16X3, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        engine.gemm(a, w, out, epilogue=epi, precision=PREC_BF16X3, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{label:38s} {us:8.1f} {nbytes / us * 1e-3:8.0f} {2.0 * M * N * K / us * 1e-6:8.1f}")
