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

# PETB200_LIB=build_tmp/chaos/libpetb200.so (`make chaos` in csrc/) selects the stress build (lib.py reads it)
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


def combine_case():
    d, E = 128, M
    t_all, m, g = rnd(E, d, seed=1), rnd(E, d, seed=2), rnd(E, d, seed=3)
    w_a, w_b = rnd(2 * d, 2 * d, seed=6, scale=(2 * d) ** -0.5), rnd(d, 2 * d, seed=8, scale=(2 * d) ** -0.5)
    s_vec, b_fold, b_b = w_a.sum(1).contiguous(), rnd(2 * d, seed=7, scale=0.1), rnd(d, seed=9, scale=0.1)
    gen = torch.Generator().manual_seed(11)
    perm = torch.randperm(E, generator=gen)
    rev = torch.empty(E, dtype=torch.int32)
    half = E // 2
    rev[perm[:half]] = perm[half:2 * half].int()
    rev[perm[half:2 * half]] = perm[:half].int()
    if E % 2:
        rev[perm[-1]] = perm[-1].int()
    rev = rev.to(dev)
    imgs = [torch.empty(h.petb200_combine_image_bytes(d, b), device=dev, dtype=torch.uint8) for b in (0, 1)]
    call("combine_pack", ptr(w_a), ptr(w_b), d, ptr(imgs[0]), ptr(imgs[1]))
    tiles = -(-E // 128)
    p1, stats = torch.zeros(tiles * 128, 2 * d, device=dev), torch.zeros(E, 2, device=dev)
    call("combine_fwd", ptr(t_all), d, ptr(rev), ptr(imgs[0]), ptr(s_vec), ptr(b_fold), ptr(b_b), E, d, ptr(m), d,
         ptr(p1), ptr(stats))
    d_cat = torch.zeros(E, 2 * d, device=dev)
    call("combine_bwd", ptr(g), d, ptr(p1), ptr(t_all), d, ptr(rev), ptr(stats), ptr(imgs[1]), ptr(s_vec),
         ptr(b_fold), E, d, ptr(d_cat))
    out = torch.zeros(E, d, device=dev)
    call("combine_scatter_bwd", ptr(d_cat), ptr(g), ptr(rev), E, d, ptr(out))
    results.update({"combine_fwd/m": m, "combine_fwd/p1": p1, "combine_fwd/stats": stats, "combine_bwd": d_cat,
                    "combine_scatter_bwd": out})


def chain_case():
    d, E, N = 128, M, M // 40
    m = rnd(E, d, seed=1)
    w1, w2 = rnd(d, d, seed=2, scale=d ** -0.5), rnd(d, d, seed=3, scale=d ** -0.5)
    b1, b2, w_e = rnd(d, seed=4, scale=0.1), rnd(d, seed=5, scale=0.1), rnd(d, seed=6, scale=d ** -0.5)
    gen = torch.Generator().manual_seed(7)
    fc = torch.rand(E, generator=gen).to(dev)
    ctr = torch.sort(torch.randint(0, N, (E,), generator=gen)).values.int().to(dev)
    d_atomic = rnd(N, 1, seed=8)
    img = [torch.empty(h.petb200_chain_image_bytes(d), device=dev, dtype=torch.uint8) for _ in range(2)]
    call("chain_pack", ptr(w1), ptr(w2), d, ptr(img[0]), ptr(img[1]))
    tiles = -(-E // 128)
    e1p, e2p, pe = torch.zeros(tiles * 128, d, device=dev), torch.zeros(E, d, device=dev), torch.zeros(E, device=dev)
    call("edge_head_fwd", ptr(m), d, ptr(img[0]), ptr(b1), ptr(b2), ptr(w_e), 0.1, E, d, ptr(e1p), ptr(e2p), ptr(pe))
    d_m, d_fc = torch.zeros(E, d, device=dev), torch.zeros(E, device=dev)
    call("edge_head_bwd", ptr(d_atomic), ptr(ctr), ptr(fc), ptr(e1p), ptr(e2p), ptr(pe), ptr(img[1]), ptr(w_e), E, d,
         ptr(d_m), d, ptr(d_fc))
    results.update({"edge_head_fwd/e1": e1p, "edge_head_fwd/e2": e2p, "edge_head_fwd/pe": pe,
                    "edge_head_bwd/d_m": d_m, "edge_head_bwd/d_fc": d_fc})
    vec, dist = rnd(E, 3, seed=9), torch.rand(E, generator=gen).to(dev) + 0.5
    geo_w, table = rnd(d, 4, seed=10, scale=0.3), rnd(2, d, seed=11, scale=0.3)
    z = torch.randint(0, 2, (E,), generator=gen).int().to(dev)
    c1, t_out = torch.zeros(tiles * 128, d, device=dev), torch.zeros(E, d, device=dev)
    call("compress_fwd", ptr(m), d, ptr(img[0]), ptr(b1), ptr(geo_w), ptr(table), ptr(z), ptr(vec), ptr(dist), ptr(b2),
         E, d, ptr(c1), ptr(t_out), d)
    d_m2, d_vec, d_dist = torch.zeros(E, d, device=dev), torch.zeros(E, 3, device=dev), torch.zeros(E, device=dev)
    call("compress_bwd", ptr(m), d, ptr(c1), ptr(img[1]), ptr(geo_w), E, d, ptr(d_m2), d, 1, ptr(d_vec), ptr(d_dist))
    results.update({"compress_fwd/c1": c1, "compress_fwd/t": t_out, "compress_bwd/d_m": d_m2,
                    "compress_bwd/d_vec": d_vec, "compress_bwd/d_dist": d_dist})


cases = [
    ("gemm 128x128 +residual (stationary)", lambda: gemm_case("gemm_none_res", 128, 128, EPI_NONE, True)),
    ("gemm 128x384 (streaming)", lambda: gemm_case("gemm_k384", 128, 384, EPI_NONE, True)),
    ("gemm 256x256 silu+aux", lambda: gemm_case("gemm_silu", 256, 256, EPI_SILU)),
    ("gemm 512x128 swiglu", lambda: gemm_case("gemm_swiglu", 512, 128, EPI_SWIGLU)),
    ("norm_linear", norm_linear_case),
    ("mlp fwd / bwd", mlp_case),
    ("attention fwd / bwd", attention_case),
    ("combine fwd / bwd / scatter", combine_case),
    ("edge head and token builder chains", chain_case),
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
