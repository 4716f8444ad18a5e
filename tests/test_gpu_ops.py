"""GPU unit tests: every C-ABI op of libpetb200 against a plain PyTorch fp32/fp64 reference
of the same op (called through the C ABI via ctypes, on ``cuda:0``)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from metatrain_b200 import engine, lib  # noqa: E402
from metatrain_b200.lib import (EPI_MUL_DSILU, EPI_NONE, EPI_SILU, EPI_SWIGLU, EPI_SWIGLU_BWD,  # noqa: E402
                                PREC_BF16, PREC_BF16X3, PREC_FP32, call, ptr)

DEV = "cuda:0"
# (precision id, absolute tolerance scale relative to fp32) for the GEMM tests
PRECISIONS = [pytest.param(PREC_FP32, 1.0, id="fp32"), pytest.param(PREC_BF16X3, 3.0, id="bf16x3"),
              pytest.param(PREC_BF16, 1500.0, id="bf16")]


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def ragged_rows(n_atoms, max_count, seed=0, with_empty=True):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(0 if with_empty else 1, max_count + 1, (n_atoms,), generator=g)
    if with_empty and n_atoms > 2:
        counts[1] = 0
        counts[2] = max_count
    row_ptr = torch.zeros(n_atoms + 1, dtype=torch.int32)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return row_ptr.to(DEV), int(row_ptr[-1]), int(counts.max())


def assert_close(a, b, atol, rtol, what=""):
    a, b = a.double().cpu(), b.double().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: max err {err.max():.3e} (|ref| max {b.abs().max():.3e}), {int(bad.sum())} bad"


# ---------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("prec,tol", PRECISIONS)
@pytest.mark.parametrize("M,N,K", [(1000, 128, 128), (77, 384, 128), (513, 128, 384),
                                   (300, 256, 256), (129, 1024, 256), (40, 256, 512), (0, 128, 128),
                                   (40000, 128, 128), (19000, 512, 1024)])
def test_gemm_plain_bias_rowscale_residual(M, N, K, prec, tol):
    a, w, b = rnd(M, K), rnd(N, K, scale=0.1), rnd(N)
    rs, res = rnd(M).abs() + 0.5, rnd(M, N)
    out = torch.empty(M, N, device=DEV)
    engine.gemm(a, w, out, bias=b, row_scale=rs, residual=res, precision=prec)
    ref = rs[:, None].double() * (a.double() @ w.double().T) + b.double() + res.double()
    assert_close(out, ref, 1e-4 * tol, 1e-5 * tol, "gemm")
    # accumulate
    out2 = out.clone()
    engine.gemm(a, w, out2, accumulate=True, precision=prec)
    assert_close(out2, ref + a.double() @ w.double().T, 2e-4 * tol, 1e-5 * tol, "gemm accumulate")


def test_gemm_bf16x3_error_level():
    """The 2-term split must sit ~2^-16 relative, far below single-pass bf16 (~2^-9)."""
    M, N, K = 4096, 128, 256
    a, w = rnd(M, K), rnd(N, K, scale=0.1)
    ref = a.double() @ w.double().T
    errs = {}
    for name, prec in (("fp32", PREC_FP32), ("bf16x3", PREC_BF16X3), ("bf16", PREC_BF16)):
        out = torch.empty(M, N, device=DEV)
        engine.gemm(a, w, out, precision=prec)
        errs[name] = float((out.double() - ref).abs().max() / ref.abs().max())
    print(errs)
    assert errs["fp32"] < 2e-6 and errs["bf16x3"] < 3e-5 and 1e-4 < errs["bf16"] < 2e-2


def test_gemm_strided_views():
    M, N, K = 333, 128, 128
    big_a, big_w = rnd(M, 3 * K), rnd(2 * N, 2 * K, scale=0.1)
    big_out = torch.zeros(M + 50, N, device=DEV)
    a, w = big_a[:, K:2 * K], big_w[N:, K:]
    engine.gemm(a, w, big_out[50:])
    assert_close(big_out[50:], a.double() @ w.double().T, 1e-4, 1e-5, "gemm views")
    assert float(big_out[:50].abs().max()) == 0.0
    # tensor-core path: strided A / C, row-slice view of a dense weight
    w_rows = rnd(3 * N, K, scale=0.1)
    big_out.zero_()
    engine.gemm(a, w_rows[N:2 * N], big_out[50:], precision=PREC_BF16X3)
    assert_close(big_out[50:], a.double() @ w_rows[N:2 * N].double().T, 3e-4, 3e-5, "gemm tc views")
    assert float(big_out[:50].abs().max()) == 0.0


@pytest.mark.parametrize("prec,tol", PRECISIONS)
@pytest.mark.parametrize("M,N,K", [(700, 128, 256), (65, 256, 128)])
def test_gemm_silu_and_dsilu(M, N, K, prec, tol):
    a, w, b = rnd(M, K), rnd(N, K, scale=0.1), rnd(N)
    out, pre = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    engine.gemm(a, w, out, bias=b, epilogue=EPI_SILU, aux_out=pre, precision=prec)
    ref_pre = a.double() @ w.double().T + b.double()
    assert_close(pre, ref_pre, 1e-4 * tol, 1e-5 * tol, "silu pre")
    assert_close(out, F.silu(ref_pre), 1e-4 * tol, 1e-5 * tol, "silu out")
    # dgrad through the SiLU: d_pre = (g @ W2) * silu'(pre)
    g, w2 = rnd(M, K, seed=3), rnd(N, K, seed=4, scale=0.1)
    d = torch.empty(M, N, device=DEV)
    engine.gemm(g, w2, d, epilogue=EPI_MUL_DSILU, aux_in=pre, precision=prec)
    p = pre.double().clone().requires_grad_(True)
    (F.silu(p) * (g.double() @ w2.double().T)).sum().backward()
    assert_close(d, p.grad, 1e-4 * tol, 1e-5 * tol, "mul_dsilu")


@pytest.mark.parametrize("prec,tol", PRECISIONS)
@pytest.mark.parametrize("M,Fdim,K", [(500, 256, 128), (90, 512, 256)])
def test_gemm_swiglu_fwd_bwd(M, Fdim, K, prec, tol):
    a, w, b = rnd(M, K), rnd(2 * Fdim, K, scale=0.1), rnd(2 * Fdim)
    rs = rnd(M).abs() + 0.5
    out, ug = torch.empty(M, Fdim, device=DEV), torch.empty(M, 2 * Fdim, device=DEV)
    engine.gemm(a, w, out, bias=b, row_scale=rs, epilogue=EPI_SWIGLU, aux_out=ug, precision=prec)
    ref_ug = rs[:, None].double() * (a.double() @ w.double().T) + b.double()
    u, g = ref_ug.chunk(2, dim=-1)
    assert_close(ug, ref_ug, 1e-4 * tol, 1e-5 * tol, "swiglu preact")
    assert_close(out, u * torch.sigmoid(g), 1e-4 * tol, 1e-5 * tol, "swiglu out")
    # backward: d_ug from d_s = go @ Wout   (Wout^T passed as the [F, d] operand)
    go, wout_t = rnd(M, 128, seed=5), rnd(Fdim, 128, seed=6, scale=0.1)
    d_ug = torch.empty(M, 2 * Fdim, device=DEV)
    engine.gemm(go, wout_t, d_ug, epilogue=EPI_SWIGLU_BWD, aux_in=ug, precision=prec)
    x = ug.double().clone().requires_grad_(True)
    uu, gg = x.chunk(2, dim=-1)
    ((uu * torch.sigmoid(gg)) * (go.double() @ wout_t.double().T)).sum().backward()
    assert_close(d_ug, x.grad, 1e-4 * tol, 1e-5 * tol, "swiglu bwd")


def test_gemm_rejects_bad_shapes():
    a, w, out = rnd(16, 128), rnd(100, 128), torch.empty(16, 100, device=DEV)
    with pytest.raises(RuntimeError, match="multiple of 128"):
        engine.gemm(a, w, out)


# ----------------------------------------------------------------------- attention
def _attention_reference(qkv, row_ptr, fc, n_atoms, n_edges, nh, scale):
    """Dense per-atom attention in fp64 with autograd (transformer.py:86-152 semantics)."""
    d = qkv.shape[1] // 3
    hd = d // nh
    out = torch.zeros(n_edges + n_atoms, d, dtype=torch.float64)
    rp = row_ptr.cpu().tolist()
    for i in range(n_atoms):
        rows = [n_edges + i] + list(range(rp[i], rp[i + 1]))
        x = qkv[rows]
        q, k, v = (x[:, j * d:(j + 1) * d].reshape(len(rows), nh, hd).transpose(0, 1) for j in range(3))
        w = torch.cat([torch.ones(1, dtype=torch.float64), fc[rp[i]:rp[i + 1]]])
        bias = torch.log(w.clamp_min(1e-15))
        a = torch.softmax(q @ k.transpose(-1, -2) * scale + bias[None, None, :], dim=-1)
        out[rows] = (a @ v).transpose(0, 1).reshape(len(rows), d)
    return out


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_BF16X3])
@pytest.mark.parametrize("n_atoms,max_count", [(37, 48), (5, 70), (3, 1), (41, 63), (23, 15), (19, 31)])
def test_attention_fwd_bwd(n_atoms, max_count, prec):
    """PREC_FP32: packed-fp32 CUDA-core kernels; PREC_BF16X3: mma.sync tensor-core kernels
    (rows of <= 63 neighbours; the (5, 70) case exercises the documented fallback)."""
    tol = 1.0 if prec == PREC_FP32 else 4.0
    nh, hd = 8, 16
    d = nh * hd
    row_ptr, E, mx = ragged_rows(n_atoms, max_count, seed=n_atoms)
    qkv = rnd(E + n_atoms, 3 * d)
    g = torch.Generator().manual_seed(1)
    fc = torch.rand(E, generator=g).to(DEV)
    if E > 3:
        fc[0] = 0.0       # clamped key
        fc[1] = 1e-20
        fc[2] = 1.0
    scale = 0.25
    out = torch.empty(E + n_atoms, d, device=DEV)
    lse = torch.empty(E + n_atoms, nh, device=DEV)
    call("attention_fwd", ptr(qkv), ptr(row_ptr), ptr(fc), n_atoms, E, nh, hd, scale, mx, prec,
         ptr(out), ptr(lse))
    x = qkv.double().cpu().requires_grad_(True)
    f = fc.double().cpu().requires_grad_(True)
    ref = _attention_reference(x, row_ptr, f, n_atoms, E, nh, scale)
    assert_close(out, ref.detach(), 2e-5 * tol, 1e-5 * tol, "attention out")
    go = rnd(E + n_atoms, d, seed=9)
    ref.backward(go.double().cpu())
    d_qkv = torch.empty_like(qkv)
    d_fc = torch.zeros(E, device=DEV)
    dsum = torch.empty(E + n_atoms, nh, device=DEV)
    call("attention_bwd", ptr(qkv), ptr(out), ptr(lse), ptr(go), ptr(row_ptr), ptr(fc), n_atoms, E,
         nh, hd, scale, mx, prec, ptr(d_qkv), ptr(d_fc), ptr(dsum))
    assert_close(d_qkv, x.grad, 5e-5 * tol, 1e-4 * tol, "attention d_qkv")
    ref_dfc = f.grad.clone()
    assert_close(d_fc, ref_dfc, 5e-4 * tol, 1e-4 * tol, "attention d_fc")


# ------------------------------------------------------------------------ row-wise
@pytest.mark.parametrize("d", [128, 256])
def test_rms_rstd_and_bwd(d):
    M = 1001
    x, gamma = rnd(M, d), rnd(d).abs() + 0.5
    rstd = torch.empty(M, device=DEV)
    call("rms_rstd", ptr(x), M, d, ptr(rstd))
    ref_rstd = torch.rsqrt((x.double() ** 2).mean(-1) + torch.finfo(torch.float32).eps)
    assert_close(rstd, ref_rstd, 1e-6, 1e-5, "rstd")
    # y = rms_norm(x) * gamma ; dy given ; d_xhat = dy * gamma
    dy, base = rnd(M, d, seed=1), rnd(M, d, seed=2)
    xx = x.double().clone().requires_grad_(True)
    y = F.rms_norm(xx, (d,), gamma.double(), torch.finfo(torch.float32).eps)
    y.backward(dy.double())
    out = torch.empty(M, d, device=DEV)
    d_xhat = (dy * gamma).contiguous()
    call("rms_bwd", ptr(d_xhat), ptr(x), ptr(rstd), ptr(base), M, d, ptr(out))
    assert_close(out, xx.grad + base.double(), 2e-5, 1e-5, "rms_bwd")
    call("rms_bwd", ptr(d_xhat), ptr(x), ptr(rstd), None, M, d, ptr(out))
    assert_close(out, xx.grad, 2e-5, 1e-5, "rms_bwd no base")


def _random_involution(E, seed=0):
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(E, generator=g)
    rev = torch.arange(E)
    for k in range(0, E - 1, 2):
        a, b = int(perm[k]), int(perm[k + 1])
        rev[a], rev[b] = b, a
    return rev.to(torch.int32).to(DEV)


def test_combine_ln_fwd_bwd_scatter():
    E, d = 999, 128
    t, gamma, beta = rnd(E, d), rnd(2 * d).abs() + 0.5, rnd(2 * d)
    rev = _random_involution(E)
    cc, mean, rstd = torch.empty(E, 2 * d, device=DEV), torch.empty(E, device=DEV), torch.empty(E, device=DEV)
    call("combine_ln_fwd", ptr(t), ptr(rev), ptr(gamma), ptr(beta), E, d, ptr(cc), ptr(mean), ptr(rstd))
    tt = t.double().clone().requires_grad_(True)
    cat = torch.cat([tt, tt[rev.long()]], dim=-1)
    ref = F.layer_norm(cat, (2 * d,), gamma.double(), beta.double(), 1e-5)
    assert_close(cc, ref.detach(), 2e-5, 1e-5, "combine_ln_fwd")
    g, base = rnd(E, 2 * d, seed=3), rnd(E, d, seed=4)
    ref.backward(g.double())
    d_cat = torch.empty(E, 2 * d, device=DEV)
    call("combine_ln_bwd", ptr(g), ptr(t), ptr(rev), ptr(gamma), ptr(mean), ptr(rstd), E, d, ptr(d_cat))
    out = torch.empty(E, d, device=DEV)
    call("combine_scatter_bwd", ptr(d_cat), ptr(base), ptr(rev), E, d, ptr(out))
    assert_close(out, tt.grad + base.double(), 5e-5, 1e-5, "combine bwd")


@pytest.mark.parametrize("E,ghosts", [(999, 0), (128, 0), (5, 0), (40000, 0), (3000, 77), (0, 0)])
def test_combine_fused_fwd_bwd(E, ghosts):
    """petb200_combine_fwd / _bwd (one tcgen05 kernel each: reversed-message gather, LayerNorm, both
    Linears, residual) against fp64 autograd of backend.py:559-575.  ``ghosts`` rows behind the E own
    rows play the halo rows of atom-sharded runs (some reversed edges point at them)."""
    d = 128
    t_all = rnd(E + ghosts, d, seed=1)
    if E > 20:
        t_all[:5] *= 20.0
        t_all[5:10] += 3.0           # rows with a large mean
    m0, g = rnd(E, d, seed=2), rnd(E, d, seed=3)
    gamma, beta = rnd(2 * d, seed=4).abs() + 0.5, rnd(2 * d, seed=5, scale=0.3)
    w_a, b_a = rnd(2 * d, 2 * d, seed=6, scale=(2 * d) ** -0.5), rnd(2 * d, seed=7, scale=0.1)
    w_b, b_b = rnd(d, 2 * d, seed=8, scale=(2 * d) ** -0.5), rnd(d, seed=9, scale=0.1)
    rev = _random_involution(E) if E > 0 else torch.zeros(0, dtype=torch.int32, device=DEV)
    if ghosts:
        rev = rev.clone()
        rev[:ghosts] = torch.arange(E, E + ghosts, dtype=torch.int32, device=DEV)
    wa_fold = (w_a.double() * gamma.double()[None, :]).float().contiguous()
    s_vec = wa_fold.double().sum(1).float().contiguous()
    b_fold = (w_a.double() @ beta.double() + b_a.double()).float().contiguous()
    handle = lib.load()
    imgs = [torch.empty(handle.petb200_combine_image_bytes(d, b), device=DEV, dtype=torch.uint8) for b in (0, 1)]
    call("combine_pack", ptr(wa_fold), ptr(w_b), d, ptr(imgs[0]), ptr(imgs[1]))
    m = m0.clone()
    tiles = -(-E // 128)
    p1, stats = torch.empty(tiles * 128, 2 * d, device=DEV), torch.empty(E, 2, device=DEV)
    call("combine_fwd", ptr(t_all), d, ptr(rev), ptr(imgs[0]), ptr(s_vec), ptr(b_fold), ptr(b_b), E, d, ptr(m), d,
         ptr(p1), ptr(stats))
    if E == 0:
        return
    tt = t_all.double().cpu().requires_grad_(True)
    rv = rev.long().cpu()
    cat = torch.cat([tt[:E], tt[rv]], dim=-1)
    pre = F.layer_norm(cat, (2 * d,), gamma.double().cpu(), beta.double().cpu(), 1e-5) @ w_a.double().cpu().T + b_a.double().cpu()
    out = m0.double().cpu() + tt[:E] + F.silu(pre) @ w_b.double().cpu().T + b_b.double().cpu()
    # private layout of the saved pre-activations: [tile][chunk of 32 units][unit][edge of the tile]
    p_rows = p1.view(tiles, 8, 32, 128).permute(0, 3, 1, 2).reshape(tiles * 128, 2 * d)[:E]
    assert_close(p_rows, pre.detach(), 2e-4, 2e-5, "combine_fwd pre-activations")
    assert_close(m, out.detach(), 2e-4, 2e-5, "combine_fwd output")
    assert_close(stats[:, 0], cat.detach().mean(1), 1e-5, 1e-5, "combine_fwd mean")
    assert_close(stats[:, 1], (cat.detach().var(1, unbiased=False) + 1e-5).rsqrt(), 1e-5, 2e-5, "combine_fwd rstd")
    # backward: gradient w.r.t. cat (before the scatter of the reversed half)
    cat.retain_grad()
    (out - tt[:E]).backward(g.double().cpu())
    d_cat = torch.empty(E, 2 * d, device=DEV)
    call("combine_bwd", ptr(g), d, ptr(p1), ptr(t_all), d, ptr(rev), ptr(stats), ptr(imgs[1]), ptr(s_vec),
         ptr(b_fold), E, d, ptr(d_cat))
    assert_close(d_cat, cat.grad, 3e-4, 3e-5, "combine_bwd")


def _chain_images(w1, w2):
    handle = lib.load()
    imgs = [torch.empty(handle.petb200_chain_image_bytes(128), device=DEV, dtype=torch.uint8) for _ in range(2)]
    call("chain_pack", ptr(w1), ptr(w2), 128, ptr(imgs[0]), ptr(imgs[1]))
    return imgs


def _private_rows(p1, E, width):
    """[tile][chunk of 32 units][unit][edge of the tile] -> [E, width]."""
    tiles = p1.shape[0] // 128
    return p1.view(tiles, width // 32, 32, 128).permute(0, 3, 1, 2).reshape(tiles * 128, width)[:E]


@pytest.mark.parametrize("E", [999, 128, 3, 40000, 0])
def test_edge_head_fused_fwd_bwd(E):
    """petb200_edge_head_fwd / _bwd against fp64 autograd of backend.py:171-217, 762-772: edge head (two
    Linears with SiLU), last layer, cutoff-weighted sum; the backward starts from d_atomic."""
    d, N = 128, max(E // 30, 1)
    m = rnd(E, d, seed=1)
    w1, b1 = rnd(d, d, seed=2, scale=d ** -0.5), rnd(d, seed=3, scale=0.1)
    w2, b2 = rnd(d, d, seed=4, scale=d ** -0.5), rnd(d, seed=5, scale=0.1)
    w_e, b_e = rnd(1, d, seed=6, scale=d ** -0.5), 0.37
    fc = torch.rand(E, device=DEV)
    ctr = torch.sort(torch.randint(0, N, (E,), generator=torch.Generator().manual_seed(7))).values.to(torch.int32).to(DEV)
    d_atomic = rnd(N, 1, seed=8)
    imgs = _chain_images(w1, w2)
    tiles = -(-E // 128)
    e1p, e2p, pe = torch.empty(tiles * 128, d, device=DEV), torch.empty(E, d, device=DEV), torch.empty(E, device=DEV)
    call("edge_head_fwd", ptr(m), d, ptr(imgs[0]), ptr(b1), ptr(b2), ptr(w_e), b_e, E, d, ptr(e1p), ptr(e2p), ptr(pe))
    if E == 0:
        return
    mm = m.double().cpu().requires_grad_(True)
    fcc = fc.double().cpu().requires_grad_(True)
    p1 = mm @ w1.double().cpu().T + b1.double().cpu()
    p2 = F.silu(p1) @ w2.double().cpu().T + b2.double().cpu()
    pred = F.silu(p2) @ w_e.double().cpu().T + b_e                     # [E, 1]
    assert_close(_private_rows(e1p, E, d), p1.detach(), 2e-4, 2e-5, "edge head: first pre-activation")
    assert_close(e2p, p2.detach(), 2e-4, 2e-5, "edge head: second pre-activation")
    assert_close(pe, pred.detach()[:, 0], 2e-4, 2e-5, "edge head: edge predictions")
    loss = (d_atomic.double().cpu()[ctr.long().cpu()] * fcc[:, None] * pred).sum()
    loss.backward()
    d_m, d_fc = torch.empty(E, d, device=DEV), torch.zeros(E, device=DEV)
    call("edge_head_bwd", ptr(d_atomic), ptr(ctr), ptr(fc), ptr(e1p), ptr(e2p), ptr(pe), ptr(imgs[1]), ptr(w_e), E, d,
         ptr(d_m), d, ptr(d_fc))
    assert_close(d_m, mm.grad, 3e-4, 3e-5, "edge head: d_m")
    assert_close(d_fc, fcc.grad, 3e-4, 3e-5, "edge head: d_fc")


@pytest.mark.parametrize("E,with_table,need_dm", [(999, True, True), (130, False, True), (40000, True, False), (0, True, True)])
def test_compress_fused_fwd_bwd(E, with_table, need_dm):
    """petb200_compress_fwd / _bwd (token builder, transformer.py:500-521 with the concatenation folded)
    against fp64 autograd."""
    d, S = 128, 5
    m = rnd(E, d, seed=1)
    vec, dist = rnd(E, 3, seed=2), rnd(E, seed=3).abs() + 0.5
    w1m, b_fold = rnd(d, d, seed=4, scale=d ** -0.5), rnd(d, seed=5, scale=0.1)
    geo_w = rnd(d, 4, seed=6, scale=0.3)
    table = rnd(S, d, seed=7, scale=0.3) if with_table else None
    z = torch.randint(0, S, (E,), generator=torch.Generator().manual_seed(8)).to(torch.int32).to(DEV)
    w2, b2 = rnd(d, d, seed=9, scale=d ** -0.5), rnd(d, seed=10, scale=0.1)
    imgs = _chain_images(w1m, w2)
    tiles = -(-E // 128)
    c1, t_out = torch.empty(tiles * 128, d, device=DEV), torch.empty(E, d, device=DEV)
    call("compress_fwd", ptr(m), d, ptr(imgs[0]), ptr(b_fold), ptr(geo_w), ptr(table), ptr(z), ptr(vec), ptr(dist),
         ptr(b2), E, d, ptr(c1), ptr(t_out), d)
    if E == 0:
        return
    mm = m.double().cpu().requires_grad_(True)
    geo = torch.cat([vec, dist[:, None]], dim=1).double().cpu().requires_grad_(True)
    pre = mm @ w1m.double().cpu().T + geo @ geo_w.double().cpu().T + b_fold.double().cpu()
    if with_table:
        pre = pre + table.double().cpu()[z.long().cpu()]
    tt = F.silu(pre) @ w2.double().cpu().T + b2.double().cpu()
    assert_close(_private_rows(c1, E, d), pre.detach(), 2e-4, 2e-5, "compress: pre-activation")
    assert_close(t_out, tt.detach(), 2e-4, 2e-5, "compress: tokens")
    d_t = rnd(E, d, seed=11)
    tt.backward(d_t.double().cpu())
    base_m, base_v, base_d = rnd(E, d, seed=12), rnd(E, 3, seed=13), rnd(E, seed=14)
    d_m, d_vec, d_dist = base_m.clone(), base_v.clone(), base_d.clone()
    call("compress_bwd", ptr(d_t), d, ptr(c1), ptr(imgs[1]), ptr(geo_w), E, d, ptr(d_m) if need_dm else None, d, 1,
         ptr(d_vec), ptr(d_dist))
    if need_dm:
        assert_close(d_m, base_m.double().cpu() + mm.grad, 3e-4, 3e-5, "compress: d_m (accumulated)")
    assert_close(d_vec, base_v.double().cpu() + geo.grad[:, :3], 3e-4, 3e-5, "compress: d_vec")
    assert_close(d_dist, base_d.double().cpu() + geo.grad[:, 3], 3e-4, 3e-5, "compress: d_dist")


def test_embedding_transpose_compress_geom():
    E, d = 777, 128
    table = rnd(5, d)
    idx = torch.randint(0, 5, (E,), generator=torch.Generator().manual_seed(0)).to(torch.int32).to(DEV)
    out = torch.empty(E, d, device=DEV)
    call("embedding", ptr(table), ptr(idx), E, d, ptr(out), d)
    assert torch.equal(out, table[idx.long()])
    # transpose + column scale
    w, cs = rnd(384, 128), rnd(128)
    wt, ws = torch.empty(128, 384, device=DEV), torch.empty(384, 128, device=DEV)
    call("transpose_scale", ptr(w), 384, 128, ptr(cs), ptr(wt), ptr(ws))
    assert torch.equal(ws, w * cs[None, :]) and torch.equal(wt, (w * cs[None, :]).T.contiguous())
    w2 = rnd(100, 36)
    w2t = torch.empty(36, 100, device=DEV)
    call("transpose_scale", ptr(w2), 100, 36, None, ptr(w2t), None)
    assert torch.equal(w2t, w2.T.contiguous())
    # compress_input
    vec, dist, wg, bg, msg = rnd(E, 3), rnd(E).abs(), rnd(d, 4), rnd(d), rnd(E, d, seed=2)
    for nbr in (None, table):
        width = 3 * d if nbr is not None else 2 * d
        cat = torch.empty(E, width, device=DEV)
        call("compress_input", ptr(vec), ptr(dist), ptr(wg), ptr(bg), ptr(nbr), ptr(idx), ptr(msg), E, d, ptr(cat))
        geo = torch.cat([vec, dist[:, None]], 1).double() @ wg.double().T + bg.double()
        parts = [geo] + ([table[idx.long()].double()] if nbr is not None else []) + [msg.double()]
        assert_close(cat, torch.cat(parts, 1), 1e-5, 1e-5, "compress_input")
    # geometry embedder backward on a strided view
    big = rnd(E, 3 * d, seed=7)
    d_vec, d_dist = torch.ones(E, 3, device=DEV), torch.ones(E, device=DEV)
    call("geom_embed_bwd", ptr(big[:, d:2 * d]), 3 * d, ptr(wg), E, d, 1, ptr(d_vec), ptr(d_dist))
    r4 = big[:, d:2 * d].double() @ wg.double()
    assert_close(d_vec, r4[:, :3] + 1, 1e-4, 1e-5, "geom bwd vec")
    assert_close(d_dist, r4[:, 3] + 1, 1e-4, 1e-5, "geom bwd dist")


def test_readout_and_sum():
    n_atoms, d, P = 41, 128, 3
    row_ptr, E, _ = ragged_rows(n_atoms, 30, seed=5)
    ctr = torch.repeat_interleave(torch.arange(n_atoms), (row_ptr[1:] - row_ptr[:-1]).long().cpu()).to(torch.int32).to(DEV)
    nf, ef = rnd(n_atoms, d), rnd(E, d)
    wn, bn, we, be = rnd(P, d, scale=0.1), rnd(P), rnd(P, d, seed=1, scale=0.1), rnd(P, seed=1)
    fc = torch.rand(E, generator=torch.Generator().manual_seed(2)).to(DEV)
    atomic, pe = torch.empty(n_atoms, P, device=DEV), torch.empty(E, P, device=DEV)
    call("readout_fwd", ptr(nf), ptr(ef), ptr(wn), ptr(bn), ptr(we), ptr(be), ptr(fc), ptr(row_ptr),
         n_atoms, E, d, P, ptr(atomic), ptr(pe))
    nfd, efd, fcd = (t.double().clone().requires_grad_(True) for t in (nf, ef, fc))
    ref_pe = efd @ we.double().T + be.double()
    ref = (nfd @ wn.double().T + bn.double()).index_add(0, ctr.long(), ref_pe * fcd[:, None])
    assert_close(pe, ref_pe.detach(), 1e-5, 1e-5, "edge pred")
    assert_close(atomic, ref.detach(), 1e-4, 1e-5, "atomic")
    g = rnd(n_atoms, P, seed=8)
    ref.backward(g.double())
    dn, de, dfc = torch.empty(n_atoms, d, device=DEV), torch.empty(E, d, device=DEV), torch.zeros(E, device=DEV)
    call("readout_bwd", ptr(g), ptr(pe), ptr(wn), ptr(we), ptr(fc), ptr(ctr), None, None, n_atoms, E, d, P,
         ptr(dn), ptr(de), ptr(dfc))
    assert_close(dn, nfd.grad, 1e-5, 1e-5, "d node feat")
    assert_close(de, efd.grad, 1e-5, 1e-5, "d edge feat")
    assert_close(dfc, fcd.grad, 1e-4, 1e-5, "d fc")
    # with the SiLU pre-activation fused
    npre, epre = rnd(n_atoms, d, seed=11), rnd(E, d, seed=12)
    call("readout_bwd", ptr(g), ptr(pe), ptr(wn), ptr(we), ptr(fc), ptr(ctr), ptr(npre), ptr(epre),
         n_atoms, E, d, P, ptr(dn), ptr(de), None)
    ds = lambda x: torch.sigmoid(x) * (1 + x * (1 - torch.sigmoid(x)))  # noqa: E731
    assert_close(dn, nfd.grad * ds(npre.double()), 1e-5, 1e-5, "d node pre")
    assert_close(de, efd.grad * ds(epre.double()), 1e-5, 1e-5, "d edge pre")
    # per-structure sums
    struct_ptr = torch.tensor([0, 10, 10, 41], dtype=torch.int32, device=DEV)
    energies = torch.empty(3, P, device=DEV)
    call("sum_over_atoms", ptr(atomic), ptr(struct_ptr), 3, P, ptr(energies))
    ref_e = torch.stack([atomic[:10].double().sum(0), torch.zeros(P, dtype=torch.float64, device=DEV),
                         atomic[10:].double().sum(0)])
    assert_close(energies, ref_e, 1e-4, 1e-5, "sum over atoms")


def test_csr_nef_roundtrip():
    n_atoms = 23
    row_ptr, E, mx = ragged_rows(n_atoms, 9, seed=3)
    counts = (row_ptr[1:] - row_ptr[:-1]).long()
    ctr = torch.repeat_interleave(torch.arange(n_atoms, device=DEV), counts).to(torch.int32)
    x = rnd(E, 3)
    nef = torch.empty(n_atoms, mx, 3, device=DEV)
    call("csr_to_nef", ptr(x), ptr(row_ptr), n_atoms, E, mx, 3, ptr(nef))
    mask = torch.arange(mx, device=DEV)[None, :] < counts[:, None]
    assert torch.equal(nef[mask], x) and float(nef[~mask].abs().sum()) == 0.0
    back = torch.empty(E, 3, device=DEV)
    call("nef_to_csr", ptr(nef), ptr(row_ptr), ptr(ctr), n_atoms, E, mx, 3, ptr(back))
    assert torch.equal(back, x)


# ------------------------------------------------------------- fused feed-forward block
def _mlp_reference(x, w_in, b_in, w_out, b_out, gamma):
    """transformer.py:229-232 (PreLN) with FeedForward :21-50 in fp64."""
    xh = F.rms_norm(x, (x.shape[1],), gamma, torch.finfo(torch.float32).eps)
    v, g = (xh @ w_in.T + b_in).chunk(2, dim=-1)
    return x + (v * torch.sigmoid(g)) @ w_out.T + b_out


@pytest.mark.parametrize("M,d_ff", [(1000, 256), (128, 64), (77, 512), (40000, 256)])
def test_mlp_fused_fwd_bwd(M, d_ff):
    """petb200_mlp_fwd / petb200_mlp_bwd (one tcgen05 kernel each) against fp64 autograd."""
    d = 128
    x, dy = rnd(M, d, seed=1), rnd(M, d, seed=2)
    x[: min(M, 5)] *= 30.0   # rows with a very different norm
    gamma = rnd(d, seed=3).abs() + 0.5
    w_in, b_in = rnd(2 * d_ff, d, seed=4, scale=d ** -0.5), rnd(2 * d_ff, seed=5, scale=0.1)
    w_out, b_out = rnd(d, d_ff, seed=6, scale=d_ff ** -0.5), rnd(d, seed=7, scale=0.1)
    w_in_folded = (w_in * gamma[None, :]).contiguous()
    handle = lib.load()
    img_f = torch.empty(handle.petb200_mlp_image_bytes(d_ff, 0), device=DEV, dtype=torch.uint8)
    img_b = torch.empty(handle.petb200_mlp_image_bytes(d_ff, 1), device=DEV, dtype=torch.uint8)
    call("mlp_pack", ptr(w_in_folded), ptr(w_out), d, d_ff, ptr(img_f), ptr(img_b))
    y = torch.empty(M, d, device=DEV)
    call("mlp_fwd", ptr(x), d, ptr(img_f), ptr(b_in), ptr(b_out), M, d, d_ff, ptr(y), d)
    xx = x.double().cpu().requires_grad_(True)
    ref = _mlp_reference(xx, w_in.double().cpu(), b_in.double().cpu(), w_out.double().cpu(),
                         b_out.double().cpu(), gamma.double().cpu())
    assert_close(y, ref.detach(), 1e-4, 2e-5, "fused mlp forward")
    ref.backward(dy.double().cpu())
    dx = torch.empty(M, d, device=DEV)
    call("mlp_bwd", ptr(x), d, ptr(dy), d, ptr(img_b), ptr(b_in), M, d, d_ff, ptr(dx), d)
    assert_close(dx, xx.grad, 1e-4, 2e-5, "fused mlp backward")


def test_mlp_fused_rejects_unsupported_width():
    x = rnd(8, 128)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        call("mlp_fwd", ptr(x), 128, ptr(x), ptr(x), ptr(x), 8, 128, 96, ptr(x), 128)


@pytest.mark.parametrize("M,K", [(1000, 384), (77, 128), (30000, 384)])
def test_gemm_rms_bwd_epilogue(M, K):
    """EPI_RMS_BWD: dgrad through Linear(RMSNorm(x)) in one kernel == gemm followed by rms_bwd."""
    d = 128
    g_out, w_t = rnd(M, K, seed=1), rnd(d, K, seed=2, scale=K ** -0.5)
    x, base = rnd(M, d, seed=3), rnd(M, d, seed=4)
    x[: min(M, 3)] *= 25.0
    rstd = torch.empty(M, device=DEV)
    call("rms_rstd", ptr(x), M, d, ptr(rstd))
    out = torch.empty(M, d, device=DEV)
    engine.gemm(g_out, w_t, out, epilogue=lib.EPI_RMS_BWD, aux_in=x, row_scale=rstd, residual=base,
                precision=PREC_BF16X3)
    dxh = g_out.double() @ w_t.double().T
    xh = x.double() * rstd.double()[:, None]
    ref = base.double() + rstd.double()[:, None] * (dxh - xh * (dxh * xh).mean(-1, keepdim=True))
    assert_close(out, ref, 2e-4, 3e-5, "gemm + rms_bwd epilogue")
    engine.gemm(g_out, w_t, out, epilogue=lib.EPI_RMS_BWD, aux_in=x, row_scale=rstd, precision=PREC_BF16X3)
    assert_close(out, ref - base.double(), 2e-4, 3e-5, "gemm + rms_bwd epilogue, no base")
    with pytest.raises(RuntimeError, match="tensor-core precisions only"):
        engine.gemm(g_out, w_t, out, epilogue=lib.EPI_RMS_BWD, aux_in=x, row_scale=rstd, precision=PREC_FP32)


@pytest.mark.parametrize("with_table", [False, True])
def test_compress_gemm_folds_the_concatenation(with_table):
    """petb200_compress_gemm == Linear(cat[W_geo.(r,d)+b_geo | NbrEmb[z] | m]) followed by SiLU
    (transformer.py:500-521) with the concatenation folded into an epilogue term."""
    E, d, S = 3001, 128, 3
    vec, dist, m = rnd(E, 3, seed=1), rnd(E, seed=2).abs() + 0.5, rnd(E, d, seed=3)
    z = torch.randint(0, S, (E,), generator=torch.Generator().manual_seed(4)).to(torch.int32).to(DEV)
    w_geo, b_geo = rnd(d, 4, seed=5, scale=0.5), rnd(d, seed=6, scale=0.1)
    nbr = rnd(S, d, seed=7) if with_table else None
    width = 3 * d if with_table else 2 * d
    w1, b1 = rnd(d, width, seed=8, scale=width ** -0.5), rnd(d, seed=9, scale=0.1)
    geo = torch.cat([vec, dist[:, None]], 1).double() @ w_geo.double().T + b_geo.double()
    parts = [geo] + ([nbr.double()[z.long()]] if with_table else []) + [m.double()]
    pre_ref = torch.cat(parts, 1) @ w1.double().T + b1.double()
    w64 = w1.double()
    geo_fold = (w64[:, :d] @ w_geo.double()).float().contiguous()
    b_fold = (b1.double() + w64[:, :d] @ b_geo.double()).float().contiguous()
    nbr_fold = (nbr.double() @ w64[:, d:2 * d].T).float().contiguous() if with_table else None
    w1m = engine.split_weight(w1[:, -d:].contiguous())
    pre, out = torch.empty(E, d, device=DEV), torch.empty(E, d, device=DEV)
    call("compress_gemm", ptr(m), d, ptr(w1m), ptr(b_fold), ptr(geo_fold), ptr(nbr_fold), ptr(z), ptr(vec),
         ptr(dist), E, d, ptr(pre), ptr(out), PREC_BF16X3)
    assert_close(pre, pre_ref, 1e-4, 3e-5, "compress_gemm pre-activation")
    assert_close(out, F.silu(pre_ref), 1e-4, 3e-5, "compress_gemm output")


@pytest.mark.parametrize("M,n_out", [(1000, 384), (128, 64), (50001, 384), (77, 1024)])
def test_norm_linear(M, n_out):
    """petb200_norm_linear: rmsnorm(x) @ W^T + b (QKV projection) and the rstd it hands to the backward."""
    d = 128
    x = rnd(M, d, seed=1)
    x[: min(M, 4)] *= 40.0
    gamma = rnd(d, seed=2).abs() + 0.5
    w, b = rnd(n_out, d, seed=3, scale=d ** -0.5), rnd(n_out, seed=4, scale=0.1)
    w_folded = (w * gamma[None, :]).contiguous()
    img = torch.empty(lib.load().petb200_norm_linear_image_bytes(n_out), device=DEV, dtype=torch.uint8)
    call("norm_linear_pack", ptr(w_folded), d, n_out, ptr(img))
    out, rstd = torch.empty(M, n_out, device=DEV), torch.empty(M, device=DEV)
    call("norm_linear", ptr(x), d, ptr(img), ptr(b), M, d, n_out, ptr(out), n_out, ptr(rstd))
    ref_rstd = torch.rsqrt((x.double() ** 2).mean(-1) + torch.finfo(torch.float32).eps)
    ref = F.rms_norm(x.double(), (d,), gamma.double(), torch.finfo(torch.float32).eps) @ w.double().T + b.double()
    assert_close(rstd, ref_rstd, 1e-6, 1e-5, "norm_linear rstd")
    assert_close(out, ref, 1e-4, 3e-5, "norm_linear output")
