"""GPU unit tests: every clipk kernel (through the C ABI) against a plain PyTorch fp32 restatement of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from easynlp_b200 import _lib as L  # noqa: E402
from easynlp_b200 import ops  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV) * scale


def assert_close(a, b, rtol, atol, name=""):
    a = a.float(); b = b.float()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), f"{name}: {int(bad.sum())}/{bad.numel()} bad, max err {err.max().item():.4g}, ref max {b.abs().max().item():.4g}"


# --------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [
    (128, 128, 64, 0, 0), (300, 384, 192, 0, 0), (1576, 2304, 768, 0, 0), (616, 768, 3072, 0, 0),
    (1576, 768, 3072, 0, 1), (200, 128, 512, 0, 1), (768, 768, 1576, 1, 1), (2304, 768, 616, 1, 1), (128, 512, 8, 1, 1),
    (8, 512, 768, 0, 1), (256, 128, 136, 1, 0),
])
def test_gemm_variants(M, N, K, a_mn, b_mn):
    A = rnd(M, K, seed=1).bfloat16(); B = rnd(N, K, seed=2).bfloat16()
    ref = A.float() @ B.float().t()
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, b, out, a_mn_major=a_mn, b_mn_major=b_mn)
    assert_close(out, ref, 1e-3, 1e-3 * math.sqrt(K), "gemm")


@pytest.mark.parametrize("M,N,K", [(394, 96, 128), (100, 288, 64), (1000, 768, 192)])
def test_gemm_bf16_store_path_slices_and_edges(M, N, K):
    """bf16 results leave through the TMA store unit: ragged M (clipped boxes), N below / across the tile width, and an output
    that is a column slice of a wider buffer (ldo > N) whose surroundings must stay untouched."""
    A = rnd(M, K, seed=21).bfloat16(); W = rnd(N, K, seed=22, scale=0.2).bfloat16(); bias = rnd(N, seed=23)
    acc = A.float() @ W.float().t() + bias
    wide = torch.full((M + 3, N + 128), 7.0, device=DEV, dtype=torch.bfloat16)
    out = wide[1:M + 1, 64:64 + N]
    ops.gemm(A, W, out, bias=bias)
    assert_close(out, acc, 1e-2, 2e-2, "slice")
    chk = wide.clone(); chk[1:M + 1, 64:64 + N] = 7.0
    assert torch.all(chk == 7.0)
    # GELU pair + multiply-by-saved-derivative with the fused column sum, same slices
    g = torch.full_like(wide, 5.0); a = torch.full_like(wide, 3.0)
    gs, as_ = g[1:M + 1, 64:64 + N], a[1:M + 1, 64:64 + N]
    ops.gemm(A, W, gs, bias=bias, mode=L.EPI_QUICK_GELU, out2=as_)
    sg = torch.sigmoid(1.702 * acc)
    assert_close(as_, acc * sg, 1e-2, 1e-2, "act")
    assert_close(gs, sg * (1 + 1.702 * acc * (1 - sg)), 1e-2, 1e-2, "act'")
    g2 = g.clone(); g2[1:M + 1, 64:64 + N] = 5.0
    assert torch.all(g2 == 5.0)
    o = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); cs = torch.zeros(N, device=DEV)
    ops.gemm(A, W, o, mode=L.EPI_MUL_AUX, aux=gs, colsum=cs)
    want = (A.float() @ W.float().t()) * gs.float()
    assert_close(o, want, 2e-2, 3e-2, "mul_aux")
    assert_close(cs, want.sum(0), 1e-3, 5e-2, "fused colsum")


def test_gemm_splitk_accumulates():
    M, N, K = 768, 3072, 1576
    A = rnd(M, K, seed=3).bfloat16(); B = rnd(N, K, seed=4).bfloat16()
    base = rnd(M, N, seed=5)
    out = base.clone()
    ops.gemm(A.t().contiguous(), B.t().contiguous(), out, a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD, splits=7)
    assert_close(out, base + A.float() @ B.float().t(), 1e-3, 1e-3 * math.sqrt(K), "splitk")


def test_gemm_epilogues():
    M, N, K = 394, 512, 256
    A = rnd(M, K, seed=6).bfloat16(); W = rnd(N, K, seed=7, scale=0.1).bfloat16()
    bias = rnd(N, seed=8); res = rnd(M, N, seed=9)
    acc = A.float() @ W.float().t()
    # linear + bias + residual -> f32 (+ bf16 shadow)
    out = torch.empty(M, N, device=DEV); out2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, out, bias=bias, residual=res, out2=out2)
    assert_close(out, acc + bias + res, 1e-3, 2e-2, "linear")
    assert_close(out2, acc + bias + res, 1e-2, 2e-2, "linear-bf16")
    # bf16 out with alpha
    ob = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, ob, bias=bias, alpha=0.5)
    assert_close(ob, 0.5 * acc + bias, 1e-2, 2e-2, "alpha")
    # GELUs: out2 = act(z), out = act'(z) (saved for backward instead of z)
    zf = (acc + bias)
    for mode, fn in ((L.EPI_QUICK_GELU, lambda z: z * torch.sigmoid(1.702 * z)), (L.EPI_ERF_GELU, torch.nn.functional.gelu)):
        g = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); a = torch.empty_like(g)
        ops.gemm(A, W, g, bias=bias, mode=mode, out2=a)
        zz = zf.clone().requires_grad_(True)
        y = fn(zz)
        y.backward(torch.ones_like(y))
        assert_close(a, y, 1e-2, 1e-2, "act")
        assert_close(g, zz.grad, 1e-2, 1e-2, "act'")
        # backward epilogue: out = acc * aux
        o = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); cs = torch.ones(N, device=DEV)
        ops.gemm(A, W, o, mode=L.EPI_MUL_AUX, aux=g, colsum=cs)
        assert_close(o, acc * g.float(), 2e-2, 3e-2, "mul_aux")
        assert_close(cs, 1 + (acc * g.float()).sum(0), 1e-3, 5e-2, "fused colsum")


# --------------------------------------------------------------------------------------------- attention
def attn_ref(qkv, mask, B, L, H):
    d = H * 64
    q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + mask[:, None, None, :]
    p = s.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * L, d)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,L,H,masked", [(2, 197, 12, False), (3, 77, 12, True), (2, 17, 2, False), (2, 16, 2, True), (1, 256, 2, False), (2, 128, 1, True)])
def test_attention_fwd_bwd(B, L, H, masked):
    d = H * 64
    qkv = rnd(B * L, 3 * d, seed=11, scale=1.5).bfloat16()
    mask = None
    if masked:
        lens = torch.randint(max(1, L // 4), L + 1, (B,), device=DEV)
        mask = ((torch.arange(L, device=DEV)[None, :] >= lens[:, None]).float() * -10000.0).contiguous()
    ctx = torch.zeros(B * L, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, L, device=DEV)
    ops.attention_fwd(qkv, mask, ctx, lse, B, L, H)
    qf = qkv.float().requires_grad_(True)
    o_ref, lse_ref = attn_ref(qf, mask, B, L, H)
    assert_close(ctx, o_ref, 2e-2, 2e-2, "ctx")
    assert_close(lse, lse_ref, 1e-3, 1e-2, "lse")
    dctx = rnd(B * L, d, seed=12).bfloat16()
    o_ref.backward(dctx.float())
    dqkv = torch.zeros(B * L, 3 * d, device=DEV, dtype=torch.bfloat16)
    dbias = torch.full((3 * d,), 0.5, device=DEV)                        # accumulated into (+=), like every gradient buffer
    ops.attention_bwd(qkv, mask, ctx, lse, dctx, dqkv, B, L, H, dqkv_colsum=dbias)
    ref = qf.grad
    scale = ref.abs().max().item()
    assert_close(dqkv, ref, 3e-2, 2e-2 * scale, "dqkv")
    cref = ref.sum(0)
    assert_close(dbias - 0.5, cref, 1e-2, 5e-3 * cref.abs().max().item() + 1e-3 * scale * math.sqrt(B * L), "fused QKV bias gradient")


@pytest.mark.parametrize("B,L,H", [(2, 257, 3), (1, 272, 2), (3, 260, 1)])
def test_attention_fwd_more_than_256_tokens(B, L, H):
    """forward only (ViT-L/14: 257 tokens): one 272-column score buffer, two K / V boxes, three query tiles"""
    d = H * 64
    qkv = rnd(B * L, 3 * d, seed=13, scale=1.5).bfloat16()
    ctx = torch.zeros(B * L, d, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, H, L, device=DEV)
    ops.attention_fwd(qkv, None, ctx, lse, B, L, H)
    o_ref, lse_ref = attn_ref(qkv.float(), None, B, L, H)
    assert_close(ctx, o_ref, 2e-2, 2e-2, "ctx")
    assert_close(lse, lse_ref, 1e-3, 1e-2, "lse")


# --------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,d,eps", [(394, 768, 1e-5), (77, 768, 1e-12), (33, 128, 1e-5), (20, 1024, 1e-5)])
def test_layernorm_fwd_bwd(rows, d, eps):
    x = rnd(rows, d, seed=21, scale=2.0) + 0.5
    g = 1 + 0.1 * rnd(d, seed=22); b = 0.1 * rnd(d, seed=23)
    yb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16); yf = torch.empty(rows, d, device=DEV)
    mean = torch.empty(rows, device=DEV); rstd = torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, g, b, eps, yb, yf, mean, rstd)
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gr, br, eps)
    assert_close(yf, ref, 1e-4, 1e-4, "ln y")
    assert_close(yb, ref, 1e-2, 1e-2, "ln y bf16")
    dy = rnd(rows, d, seed=24); dy_add = rnd(rows, d, seed=25); dx_add = rnd(rows, d, seed=26)
    ref.backward(dy + dy_add)
    dxf = torch.empty(rows, d, device=DEV); dxb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16)
    dg = torch.zeros(d, device=DEV); db = torch.zeros(d, device=DEV); dbias = torch.zeros(d, device=DEV)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dy_add=dy_add, dx_add=dx_add, dx_f32=dxf, dx_bf16=dxb, dgamma=dg, dbeta=db, dbias=dbias)
    assert_close(dxf, xr.grad + dx_add, 1e-3, 1e-3, "ln dx")
    assert_close(dxb, xr.grad + dx_add, 1e-2, 2e-2, "ln dx bf16")
    assert_close(dg, gr.grad, 1e-3, 1e-3 * math.sqrt(rows), "ln dgamma")
    assert_close(db, br.grad, 1e-3, 1e-3 * math.sqrt(rows), "ln dbeta")
    assert_close(dbias, (xr.grad + dx_add).sum(0), 1e-3, 1e-3 * math.sqrt(rows), "ln dbias")
    # fused residual add: LN(x + add) with the sum written out
    add = rnd(rows, d, seed=28).bfloat16(); xo = torch.empty(rows, d, device=DEV)
    ops.layernorm_fwd(x, g, b, eps, yb, yf, mean, rstd, add=add, x_out=xo)
    assert_close(xo, x + add.float(), 1e-6, 1e-6, "ln add x_out")
    assert_close(yf, torch.nn.functional.layer_norm(x + add.float(), (d,), g, b, eps), 1e-4, 1e-4, "ln add y")
    # bf16 dy, strided x rows (CLS pooling)
    L_ = 5
    xs = rnd(rows * L_, d, seed=27)
    ops.layernorm_fwd(xs, g, b, eps, yb, None, mean, rstd, rows=rows, ldx=L_ * d)
    assert_close(yb, torch.nn.functional.layer_norm(xs.view(rows, L_, d)[:, 0], (d,), g, b, eps), 1e-2, 1e-2, "ln strided")


def test_colsum():
    x = rnd(1000, 768, seed=31)
    out = torch.ones(768, device=DEV)
    ops.colsum(x, out, 1000, 768)
    assert_close(out, 1 + x.sum(0), 1e-4, 1e-3, "colsum f32")
    xb = x.bfloat16()
    out.zero_()
    ops.colsum(xb, out, 1000, 768)
    assert_close(out, xb.float().sum(0), 1e-4, 1e-3, "colsum bf16")


# --------------------------------------------------------------------------------------------- embeddings
def test_vit_patch_pipeline():
    B, R, P, W = 3, 64, 16, 128
    g = R // P; Lv = g * g + 1
    pix = rnd(B, 3, R, R, seed=41)
    patches = torch.empty(B * g * g, 3 * P * P, device=DEV, dtype=torch.bfloat16)
    ops.im2col_patches(pix, patches, B, R, P)
    ref = torch.nn.functional.unfold(pix, P, stride=P).transpose(1, 2).reshape(B * g * g, 3 * P * P)
    assert_close(patches, ref, 1e-2, 1e-2, "im2col")
    patch_out = rnd(B * g * g, W, seed=42); cls = rnd(W, seed=43); pos = rnd(Lv, W, seed=44)
    x0 = torch.empty(B * Lv, W, device=DEV)
    ops.vit_assemble(patch_out, cls, pos, x0, B, Lv, W)
    refx = torch.cat([cls.expand(B, 1, W), patch_out.view(B, g * g, W)], 1) + pos
    assert_close(x0.view(B, Lv, W), refx, 1e-6, 1e-6, "assemble")
    dp = torch.empty(B * g * g, W, device=DEV, dtype=torch.bfloat16)
    ops.vit_assemble_bwd(x0, dp, B, Lv, W)
    assert_close(dp.view(B, g * g, W), x0.view(B, Lv, W)[:, 1:], 1e-2, 1e-2, "assemble bwd")


def test_bert_embed():
    B, Lt, H, V = 4, 16, 128, 512
    ids = torch.randint(0, V, (B, Lt), device=DEV)
    ids[:, -3:] = 0
    word = rnd(V, H, seed=51); pos = rnd(64, H, seed=52); typ = rnd(2, H, seed=53)
    e = torch.empty(B * Lt, H, device=DEV)
    ops.bert_embed(ids.view(-1), word, pos, typ, e, B * Lt, Lt, H, V)
    assert_close(e.view(B, Lt, H), word[ids] + pos[:Lt] + typ[0], 1e-6, 1e-6, "embed")
    de = rnd(B * Lt, H, seed=54)
    dword = torch.zeros(V, H, device=DEV)
    ops.bert_embed_bwd(ids.view(-1), de, dword, B * Lt, H, V)
    ref = torch.zeros(V, H, device=DEV).index_add_(0, ids.view(-1), de)
    ref[0] = 0
    assert_close(dword, ref, 1e-5, 1e-5, "embed bwd")


def test_l2norm():
    x = rnd(37, 512, seed=61)
    y = torch.empty_like(x); n = torch.empty(37, device=DEV)
    ops.l2norm_fwd(x, y, n, 37, 512)
    xr = x.clone().requires_grad_(True)
    ref = xr / xr.norm(dim=-1, keepdim=True)
    assert_close(y, ref, 1e-5, 1e-6, "l2norm")
    dy = rnd(37, 512, seed=62)
    ref.backward(dy)
    dx = torch.empty_like(x); dxb = torch.empty(37, 512, device=DEV, dtype=torch.bfloat16)
    ops.l2norm_bwd(dy, y, n, dx, dxb, 37, 512)
    assert_close(dx, xr.grad, 1e-4, 1e-6, "l2norm bwd")


# --------------------------------------------------------------------------------------------- loss
@pytest.mark.parametrize("nq,nk,E,off", [(8, 8, 512, 0), (50, 200, 64 * 2, 100), (256, 256, 512, 0), (33, 70, 768, 7)])
def test_ce_strip(nq, nk, E, off):
    Q = torch.nn.functional.normalize(rnd(nq, E, seed=71), dim=-1)
    K = torch.nn.functional.normalize(rnd(nk, E, seed=72), dim=-1)
    ls = torch.tensor(math.log(1 / 0.07), device=DEV)
    lse = torch.empty(nq, device=DEV); rows = torch.empty(nq, device=DEV)
    S = torch.empty(nq, nk, device=DEV); St = torch.empty(nk, nq, device=DEV)
    ops.ce_strip_fwd(Q, K, ls, off, lse, rows, S_out=S, lds=nk)
    ops.ce_strip_fwd(Q, K, ls, off, lse, rows, S_out=St, lds=nq, transpose_out=True)
    Qr = Q.clone().requires_grad_(True); Kr = K.clone().requires_grad_(True); lr = ls.clone().requires_grad_(True)
    Sr = (Qr @ Kr.t()) * lr.exp()
    lab = off + torch.arange(nq, device=DEV)
    per = torch.nn.functional.cross_entropy(Sr, lab, reduction="none")
    assert_close(S, Sr, 1e-5, 1e-5, "S")
    assert_close(St.t(), Sr, 1e-5, 1e-5, "S^T")
    assert_close(rows, per, 1e-4, 1e-5, "per-row CE")
    assert_close(lse, torch.logsumexp(Sr, -1), 1e-5, 1e-5, "lse")
    coef = 0.37
    (per.sum() * coef).backward()
    dQ = torch.ones(nq, E, device=DEV); dK = torch.zeros(nk, E, device=DEV); dls = torch.zeros(1, device=DEV)
    ops.ce_strip_bwd(Q, K, ls, lse, off, coef, True, dQ, True, dls)
    ops.ce_strip_bwd(K, Q, ls, lse, off, coef, False, dK, False)
    assert_close(dQ - 1, Qr.grad, 1e-3, 1e-5, "dQ")
    assert_close(dK, Kr.grad, 1e-3, 1e-5, "dK")
    assert_close(dls, lr.grad.view(1), 1e-3, 1e-4, "dscale")
    tot = torch.zeros(1, device=DEV)
    ops.reduce_sum(rows, nq, 0.5, tot)
    assert_close(tot, per.sum().view(1) * 0.5, 1e-5, 1e-5, "reduce")


@pytest.mark.parametrize("nq,nk,E,off,scale", [(8, 8, 512, 0, 1 / 0.07), (50, 200, 128, 100, 1 / 0.07), (256, 256, 512, 0, 100.0), (33, 70, 768, 7, 30.0),
                                               (256, 2048, 512, 512, 100.0)])
def test_ce_tensor_core_path(nq, nk, E, off, scale):
    """what the training step runs: dots from ONE K = 3E GEMM over bf16 hi/lo splits (fp32-level logits), row kernels for
    lse / loss / dS, gradient GEMMs in bf16 -- against the fp32 torch loss"""
    Q = torch.nn.functional.normalize(rnd(nq, E, seed=71), dim=-1)
    K = torch.nn.functional.normalize(rnd(nk, E, seed=72), dim=-1)
    ls = torch.tensor(math.log(scale), device=DEV)
    nkp = (nk + 7) // 8 * 8
    Qs = torch.empty(nq, 3 * E, device=DEV, dtype=torch.bfloat16); Ks = torch.zeros(nkp, 3 * E, device=DEV, dtype=torch.bfloat16)
    ops.split_bf16x3(Q, Qs, 0); ops.split_bf16x3(K, Ks, 1)
    S = torch.empty(nq, nkp, device=DEV)
    ops.gemm(Qs, Ks, S)
    assert_close(S[:, :nk], Q @ K.t(), 0.0, 3e-6, "hi/lo split dots")          # bf16 x 3 ~ 2^-17 relative on |q||k| = 1
    lse = torch.empty(nq, device=DEV); rows = torch.empty(nq, device=DEV)
    ops.ce_rows_fwd(S, ls, off, lse, rows, nq, nk)
    Qr = Q.clone().requires_grad_(True); Kr = K.clone().requires_grad_(True); lr = ls.clone().requires_grad_(True)
    Sr = (Qr @ Kr.t()) * lr.exp()
    lab = off + torch.arange(nq, device=DEV)
    per = torch.nn.functional.cross_entropy(Sr, lab, reduction="none")
    assert_close(S[:, :nk], Sr, 1e-5, 3e-6 * scale, "scaled logits")
    assert_close(lse, torch.logsumexp(Sr, -1), 1e-5, 3e-6 * scale, "lse")
    assert_close(rows, per, 1e-4, 1e-5 * scale, "per-row CE")
    coef = 0.37
    (per.sum() * coef).backward()
    dS = torch.full((nq, nkp), 9.0, device=DEV, dtype=torch.bfloat16); dls = torch.zeros(1, device=DEV)
    ops.ce_rows_bwd(S, ls, lse, off, coef, dS, nq, nk, dscale_log=dls)
    assert torch.all(dS[:, nk:] == 0)
    dQ = torch.empty(nq, E, device=DEV); dK = torch.empty(nk, E, device=DEV)
    ops.gemm(dS, Ks[:, :E], dQ, b_mn_major=1)
    ops.gemm(dS[:, :nk], Qs[:, :E], dK, a_mn_major=1, b_mn_major=1)
    gq = Qr.grad.abs().max().item(); gk = Kr.grad.abs().max().item()
    assert_close(dQ, Qr.grad, 2e-2, 1e-2 * gq, "dQ (bf16 operands)")
    assert_close(dK, Kr.grad, 2e-2, 1e-2 * gk, "dK (bf16 operands)")
    assert_close(dls, lr.grad.view(1), 2e-3, 1e-3 * abs(lr.grad.item()) + 1e-4, "dscale")


@pytest.mark.parametrize("nq,nk,E,off", [(24, 24, 128, 0), (100, 1000, 512, 300), (513, 2049, 512, 0), (64, 40, 768, -8)])
def test_retrieval_rank_tensor_core(nq, nk, E, off):
    """rank-count GEMM epilogue (no N x N matrix) == brute-force fp32 ranks; ragged nq / nk, labels partly outside the gallery"""
    Q = torch.nn.functional.normalize(rnd(nq, E, seed=91), dim=-1)
    K = torch.nn.functional.normalize(rnd(nk, E, seed=92), dim=-1)
    lab = off + torch.arange(nq, device=DEV)
    ok = (lab >= 0) & (lab < nk)
    K[lab[ok]] = torch.nn.functional.normalize(K[lab[ok]] + 0.35 * Q[ok], dim=-1)        # matches rank high but not always first
    S = (Q.double() @ K.double().t())
    thr = torch.where(ok, S[torch.arange(nq, device=DEV), lab.clamp(0, nk - 1)], torch.full((nq,), float("inf"), device=DEV, dtype=torch.float64))
    want = ((S > thr[:, None]) & (torch.arange(nk, device=DEV)[None, :] != lab[:, None])).sum(1).int()
    margin = (S - thr[:, None]).abs()
    margin[torch.arange(nq, device=DEV)[ok], lab[ok]] = 1.0
    assert margin.min().item() > 1e-5                                       # no near-ties in this fixture: the ranks are well defined
    got = torch.full((nq,), -1, device=DEV, dtype=torch.int32)
    ops.retrieval_rank_tc(Q, K, got, label_offset=off)
    assert torch.equal(got, want), (got[:8], want[:8])
    ref = torch.empty(nq, device=DEV, dtype=torch.int32)
    if off >= 0 and off + nq <= nk:
        ops.retrieval_rank(Q, K, ref, label_offset=off)                     # the fp32 CUDA-core kernel agrees
        assert torch.equal(ref, want)


# --------------------------------------------------------------------------------------------- optimizer
def test_adamw_and_gradnorm():
    from oracle import clip_oracle as O
    n = 4096 * 3 + 8
    p = rnd(n, seed=81); g = rnd(n, seed=82, scale=0.01)
    m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    wb = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ws = torch.empty(1024, device=DEV, dtype=torch.float64); nc = torch.empty(2, device=DEV)
    pr, mr, vr = p.cpu().clone(), m.cpu().clone(), v.cpu().clone()
    for step in (1, 2, 3):
        ops.grad_norm(g, n, 1.0, ws, nc)
        gc = g.cpu().clone()
        tot = O.clip_grad_norm([gc], 1.0)
        assert_close(nc[0:1].cpu(), tot.view(1), 1e-5, 1e-6, "norm")
        ops.adamw_step(p, g, m, v, wb, n, 1e-3, 1e-4, step, clip_coef=nc[1:])
        O.adamw_step(pr, gc, mr, vr, step, 1e-3, 1e-4)
        assert_close(p.cpu(), pr, 1e-5, 1e-6, "adamw p")
        assert_close(wb.cpu(), pr, 1e-2, 1e-2, "adamw bf16 copy")
        g = g * 1.5 + 0.001
