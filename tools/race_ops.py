#!/usr/bin/env python
"""Bit-reproducibility of the persistent kernels under a changed timing.

    python tools/race_ops.py save  out.pt      # plain run: store every op's outputs
    compute-sanitizer python tools/race_ops.py check out.pt   # slowed-down run: must be bit-identical

Every op runs on enough rows that a persistent CTA walks several tiles (the case the small op tests do not
reach).  All kernels are deterministic, so any difference is a synchronisation bug (or a tool artefact)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from metatrain_b200 import engine, lib  # noqa: E402
from metatrain_b200.lib import (EPI_NONE, EPI_SILU, EPI_SWIGLU, PREC_BF16X3, call, ptr)  # noqa: E402

dev = "cuda:0"
mode, path = sys.argv[1], sys.argv[2]
M = int(sys.argv[3]) if len(sys.argv) > 3 else 120000
h = lib.load()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


results = {}


def gemm_case(name, N, K, epi, residual=False):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    n_out = N // 2 if epi == EPI_SWIGLU else N
    out = torch.zeros(M, n_out, device=dev)
    kw = {}
    if epi in (EPI_SILU, EPI_SWIGLU):
        kw["aux_out"] = torch.zeros(M, N, device=dev)
    if residual:
        kw["residual"] = rnd(M, N, seed=3)
    engine.gemm(a, w, out, bias=rnd(N, seed=4), epilogue=epi, precision=PREC_BF16X3, **kw)
    results[name] = out
    if "aux_out" in kw:
        results[name + "/aux"] = kw["aux_out"]


def norm_linear_case():
    d, n_out = 128, 384
    x, w, b = rnd(M, d, seed=1), rnd(n_out, d, seed=3, scale=d ** -0.5), rnd(n_out, seed=4, scale=0.1)
    img = torch.empty(h.petb200_norm_linear_image_bytes(n_out), device=dev, dtype=torch.uint8)
    call("norm_linear_pack", ptr(w), d, n_out, ptr(img))
    out, rstd = torch.zeros(M, n_out, device=dev), torch.zeros(M, device=dev)
    call("norm_linear", ptr(x), d, ptr(img), ptr(b), M, d, n_out, ptr(out), n_out, ptr(rstd))
    results["norm_linear"], results["norm_linear/rstd"] = out, rstd


def mlp_case():
    d, F = 128, 256
    x, dy = rnd(M, d, seed=1), rnd(M, d, seed=2)
    w_in, b_in = rnd(2 * F, d, seed=4, scale=d ** -0.5), rnd(2 * F, seed=5, scale=0.1)
    w_out, b_out = rnd(d, F, seed=6, scale=F ** -0.5), rnd(d, seed=7, scale=0.1)
    img = [torch.empty(h.petb200_mlp_image_bytes(F, b), device=dev, dtype=torch.uint8) for b in (0, 1)]
    call("mlp_pack", ptr(w_in), ptr(w_out), d, F, ptr(img[0]), ptr(img[1]))
    y, dx = torch.zeros(M, d, device=dev), torch.zeros(M, d, device=dev)
    call("mlp_fwd", ptr(x), d, ptr(img[0]), ptr(b_in), ptr(b_out), M, d, F, ptr(y), d)
    call("mlp_bwd", ptr(x), d, ptr(dy), d, ptr(img[1]), ptr(b_in), M, d, F, ptr(dx), d)
    results["mlp_fwd"], results["mlp_bwd"] = y, dx


def attention_case():
    nh, hd, d = 8, 16, 128
    n_atoms = M // 40
    g = torch.Generator().manual_seed(5)
    counts = torch.randint(30, 48, (n_atoms,), generator=g)
    row_ptr = torch.zeros(n_atoms + 1, dtype=torch.int32)
    row_ptr[1:] = torch.cumsum(counts, 0).int()
    E, mx = int(row_ptr[-1]), int(counts.max())
    row_ptr = row_ptr.to(dev)
    qkv, fc, go = rnd(E + n_atoms, 3 * d, seed=1), torch.rand(E, generator=g).to(dev), rnd(E + n_atoms, d, seed=9)
    out, lse = torch.zeros(E + n_atoms, d, device=dev), torch.zeros(E + n_atoms, nh, device=dev)
    call("attention_fwd", ptr(qkv), ptr(row_ptr), ptr(fc), n_atoms, E, nh, hd, 0.25, mx, PREC_BF16X3, ptr(out), ptr(lse))
    d_qkv, d_fc = torch.zeros_like(qkv), torch.zeros(E, device=dev)
    dsum = torch.zeros((E + n_atoms) * nh * 2, device=dev)
    call("attention_bwd", ptr(qkv), ptr(out), ptr(lse), ptr(go), ptr(row_ptr), ptr(fc), n_atoms, E, nh, hd, 0.25, mx,
         PREC_BF16X3, ptr(d_qkv), ptr(d_fc), ptr(dsum))
    results.update({"attention_fwd": out, "attention_fwd/lse": lse, "attention_bwd": d_qkv, "attention_bwd/d_fc": d_fc})


cases = [
    ("gemm 128x128 +residual (stationary)", lambda: gemm_case("gemm_none_res", 128, 128, EPI_NONE, True)),
    ("gemm 128x384 (streaming)", lambda: gemm_case("gemm_k384", 128, 384, EPI_NONE, True)),
    ("gemm 256x256 silu+aux", lambda: gemm_case("gemm_silu", 256, 256, EPI_SILU)),
    ("gemm 512x128 swiglu", lambda: gemm_case("gemm_swiglu", 512, 128, EPI_SWIGLU)),
    ("norm_linear", norm_linear_case),
    ("mlp fwd / bwd", mlp_case),
    ("attention fwd / bwd", attention_case),
]
for label, fn in cases:
    fn()
    torch.cuda.synchronize()
    print("ran", label, flush=True)
results = {k: v.cpu() for k, v in results.items()}
if mode == "save":
    torch.save(results, path)
else:
    ref = torch.load(path)
    for k, v in results.items():
        same = torch.equal(v, ref[k])
        diff = (v - ref[k]).abs().max().item() if not same else 0.0
        bad = int((v != ref[k]).any(dim=-1).sum()) if (not same and v.dim() > 1) else int((v != ref[k]).sum())
        print(f"{k:24s} {'bit-identical' if same else f'DIFFERS: max abs {diff:.3e}, {bad} rows'}")
