#!/usr/bin/env python
"""Time the fused feed-forward kernels against the unfused GEMM sequence (E = 392 040 rows)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metatrain_b200 import engine, lib  # noqa: E402
from metatrain_b200.lib import EPI_SWIGLU, EPI_SWIGLU_BWD, PREC_BF16X3, call, ptr  # noqa: E402

dev = "cuda:0"
M, d, F = int(sys.argv[1]) if len(sys.argv) > 1 else 392040, 128, 256
torch.manual_seed(0)
x, dy = torch.randn(M, d, device=dev), torch.randn(M, d, device=dev)
w_in, b_in = torch.randn(2 * F, d, device=dev) * d ** -0.5, torch.randn(2 * F, device=dev) * 0.1
w_out, b_out = torch.randn(d, F, device=dev) * F ** -0.5, torch.randn(d, device=dev) * 0.1
w_in_t, w_out_t = w_in.T.contiguous(), w_out.T.contiguous()
h = lib.load()
img = [torch.empty(h.petb200_mlp_image_bytes(F, b), device=dev, dtype=torch.uint8) for b in (0, 1)]
call("mlp_pack", ptr(w_in), ptr(w_out), d, F, ptr(img[0]), ptr(img[1]))
y, dx = torch.empty(M, d, device=dev), torch.empty(M, d, device=dev)
rstd, ug, s = torch.empty(M, device=dev), torch.empty(M, 2 * F, device=dev), torch.empty(M, F, device=dev)
d_ug, d_xh = torch.empty(M, 2 * F, device=dev), torch.empty(M, d, device=dev)


def fused_fwd():
    call("mlp_fwd", ptr(x), d, ptr(img[0]), ptr(b_in), ptr(b_out), M, d, F, ptr(y), d)


def fused_bwd():
    call("mlp_bwd", ptr(x), d, ptr(dy), d, ptr(img[1]), ptr(b_in), M, d, F, ptr(dx), d)


def unfused_fwd():
    call("rms_rstd", ptr(x), M, d, ptr(rstd))
    engine.gemm(x, w_in, s, bias=b_in, row_scale=rstd, epilogue=EPI_SWIGLU, aux_out=ug, precision=PREC_BF16X3)
    engine.gemm(s, w_out, y, bias=b_out, residual=x, precision=PREC_BF16X3)


def unfused_bwd():
    engine.gemm(dy, w_out_t, d_ug, epilogue=EPI_SWIGLU_BWD, aux_in=ug, precision=PREC_BF16X3)
    engine.gemm(d_ug, w_in_t, d_xh, precision=PREC_BF16X3)
    call("rms_bwd", ptr(d_xh), ptr(x), ptr(rstd), ptr(dy), M, d, ptr(dx))


for name, fn, kb in (("fused fwd", fused_fwd, 1.0), ("fused bwd", fused_bwd, 1.5),
                     ("unfused fwd", unfused_fwd, 1.0), ("unfused bwd", unfused_bwd, 1.5)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = 2.0 * M * d * 3 * F * (1 if "fwd" in name else 5.0 / 3.0)
    print(f"{name:12s} {ms * 1e3:8.1f} us   {M * kb * 1024 / ms / 1e6:7.0f} GB/s (algorithmic minimum bytes)"
          f"   {flops / ms / 1e9:7.1f} TFLOP/s")
