"""GPU: the fused Philox dropout of the text tower (BertEmbeddings / attention probabilities / BertSelfOutput / BertOutput,
modeling_bert.py:128,238,267,345).  torch's RNG stream cannot be matched bit for bit, so the masks the kernels generate are dumped
with clipk_dropout_mask and fed to the oracle as explicit multipliers; statistics are checked separately."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from easynlp_b200 import ops  # noqa: E402
from easynlp_b200.engine import ClipEngine  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mask_of(rows, cols, p, seed, site, off=None):
    m = torch.empty(rows, cols, device=DEV)
    ops.dropout_mask(m, rows, cols, ops.make_dropout(p, seed, site, off))
    return m


def test_mask_statistics_and_determinism():
    p = 0.1
    m = mask_of(2048, 768, p, 1234, 7)
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - 1 / (1 - p)) < 1e-6
    frac = (m == 0).float().mean().item()
    n = m.numel()
    assert abs(frac - p) < 5 * math.sqrt(p * (1 - p) / n)                    # binomial 5 sigma
    assert abs(m.mean().item() - 1.0) < 5e-3                                  # inverted dropout keeps the expectation
    assert torch.equal(m, mask_of(2048, 768, p, 1234, 7))                     # same key -> same mask
    for other in (mask_of(2048, 768, p, 1235, 7), mask_of(2048, 768, p, 1234, 8),
                  mask_of(2048, 768, p, 1234, 7, torch.tensor([3], dtype=torch.int32, device=DEV))):
        assert (other != m).float().mean().item() > 0.1                       # seed / site / per-step offset decorrelate
    col_frac = (m == 0).float().mean(0)
    assert col_frac.max().item() < p + 6 * math.sqrt(p * (1 - p) / 2048)     # no column / row structure
    assert mask_of(64, 128, 0.0, 1, 1).eq(1).all()


def test_layernorm_dropout_modes():
    rows, d, eps, p = 77, 768, 1e-12, 0.25
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(rows, d, generator=g, device=DEV); add = torch.randn(rows, d, generator=g, device=DEV).bfloat16()
    gam = 1 + 0.1 * torch.randn(d, generator=g, device=DEV); bet = 0.1 * torch.randn(d, generator=g, device=DEV)
    dy = torch.randn(rows, d, generator=g, device=DEV)
    drop = ops.make_dropout(p, 99, 5)
    m = mask_of(rows, d, p, 99, 5)
    yb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16); yf = torch.empty(rows, d, device=DEV); xo = torch.empty(rows, d, device=DEV)
    mean = torch.empty(rows, device=DEV); rstd = torch.empty(rows, device=DEV)
    # mode 1: LN(x + dropout(add))
    ops.layernorm_fwd(x, gam, bet, eps, yb, yf, mean, rstd, add=add, x_out=xo, drop=drop, drop_mode=1)
    xr = x.clone().requires_grad_(True); ar = add.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr + ar * m, (d,), gam, bet, eps)
    assert torch.allclose(yf, ref, atol=2e-4) and torch.allclose(xo, x + add.float() * m, atol=1e-6)
    ref.backward(dy)
    dxf = torch.empty(rows, d, device=DEV); dxb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16); dbias = torch.zeros(d, device=DEV)
    ops.layernorm_bwd(dy, xo, gam, mean, rstd, dx_f32=dxf, dx_bf16=dxb, dbias=dbias, drop=drop, drop_mode=1)
    assert torch.allclose(dxf, xr.grad, atol=2e-3)                            # residual path: unmasked
    assert torch.allclose(dxb.float(), ar.grad, atol=3e-2, rtol=2e-2)         # branch path: masked
    assert torch.allclose(dbias, ar.grad.sum(0), atol=2e-2, rtol=1e-3)
    # mode 2: dropout(LN(x))
    ops.layernorm_fwd(x, gam, bet, eps, yb, yf, mean, rstd, drop=drop, drop_mode=2)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gam, bet, eps) * m
    assert torch.allclose(yf, ref, atol=2e-4)
    ref.backward(dy)
    ops.layernorm_bwd(dy, x, gam, mean, rstd, dx_f32=dxf, drop=drop, drop_mode=2)
    assert torch.allclose(dxf, xr.grad, atol=2e-3)


@pytest.mark.parametrize("B,L,H", [(3, 77, 12), (2, 16, 2), (2, 128, 1)])
def test_attention_dropout(B, L, H):
    d = H * 64; p = 0.2
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = (torch.randn(B * L, 3 * d, generator=g, device=DEV) * 1.5).bfloat16()
    lens = torch.randint(max(1, L // 4), L + 1, (B,), device=DEV)
    kmask = ((torch.arange(L, device=DEV)[None, :] >= lens[:, None]).float() * -10000.0).contiguous()
    drop = ops.make_dropout(p, 4321, 16)
    lk_pad = (L + 15) // 16 * 16
    pm = mask_of(B * H * L, lk_pad, p, 4321, 16).view(B, H, L, lk_pad)[..., :L]
    ctx = torch.zeros(B * L, d, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, H, L, device=DEV)
    ops.attention_fwd(qkv, kmask, ctx, lse, B, L, H, drop=drop)
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0 + kmask[:, None, None, :]
    o_ref = ((s.softmax(-1) * pm) @ v).permute(0, 2, 1, 3).reshape(B * L, d)
    err = (ctx.float() - o_ref).abs().max().item()
    assert err < 3e-2 * max(1.0, o_ref.abs().max().item()), err
    dctx = torch.randn(B * L, d, generator=g, device=DEV).bfloat16()
    o_ref.backward(dctx.float())
    dqkv = torch.zeros(B * L, 3 * d, device=DEV, dtype=torch.bfloat16)
    ops.attention_bwd(qkv, kmask, ctx, lse, dctx, dqkv, B, L, H, drop=drop)
    ref = qf.grad
    assert (dqkv.float() - ref).abs().max().item() < 3e-2 * ref.abs().max().item() + 3e-2 * 1.0


def test_model_with_dropout_matches_oracle_given_the_same_masks():
    z = np.load(os.path.join(GOLD, "tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    cfg = dict(cfg, text_hidden_dropout_prob=0.1, text_attention_probs_dropout_prob=0.1)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    eng = ClipEngine(cfg)
    eng.params.load_state_dict(sd)
    B, Lt = 10, 24
    pixels, ids = O.synthetic_batch(cfg, B, seq_len=Lt, seed=31)
    out = eng.forward(pixels.cuda(), ids.cuda(), save=True, train=True)
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    H = cfg["text_hidden_size"]; heads = cfg["text_num_attention_heads"]; M = B * Lt
    lk_pad = (Lt + 15) // 16 * 16
    off = eng._dev_pass          # dropout stream position of the forward pass just made (its backward saw the same value)
    drop = {"emb": mask_of(M, H, 0.1, eng.dropout_seed, 1, off).view(B, Lt, H).cpu()}
    for i in range(cfg["text_num_hidden_layers"]):
        drop[("attn", i)] = mask_of(B * heads * Lt, lk_pad, 0.1, eng.dropout_seed, 16 * (i + 1), off).view(B, heads, Lt, lk_pad)[..., :Lt].cpu()
        drop[("self_out", i)] = mask_of(M, H, 0.1, eng.dropout_seed, 16 * (i + 1) + 1, off).view(B, Lt, H).cpu()
        drop[("out", i)] = mask_of(M, H, 0.1, eng.dropout_seed, 16 * (i + 1) + 2, off).view(B, Lt, H).cpu()
    names = O.trainable_names(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(params)
    ref = O.clip_forward(full, cfg, pixels, ids, drop=drop)
    loss = O.clip_loss(ref["logits_per_text"])
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    # the eval-mode loss differs clearly from the dropout loss, so agreement below is evidence the masks were applied identically
    ref_eval = O.clip_loss(O.clip_forward(sd, cfg, pixels, ids)["logits_per_text"]).item()
    assert abs(loss.item() - ref_eval) > 5e-3
    # 1 %: 10-pair tiny model (bf16 noise is ~0.3 % here even without dropout); the eval-mode loss is 5 % away
    assert abs(out["loss"].item() - loss.item()) < 1e-2 * loss.item(), (out["loss"].item(), loss.item(), ref_eval)
    assert abs(out["loss"].item() - loss.item()) < 0.25 * abs(ref_eval - loss.item())
    assert (out["text_embeds"].cpu() - ref["text_embeds"]).abs().max().item() < 6e-3
    gnorm = math.sqrt(sum(float(g.double().norm()) ** 2 for g in grads if g is not None))
    for k, gr in zip(names, grads):
        if gr is None:
            continue
        err = (eng.params.g(k).cpu() - gr).norm().item()
        assert err < 0.15 * gr.norm().item() + 2e-4 * gnorm, (k, err, gr.norm().item())
    # eval / encode path: dropout off
    e1 = eng.encode(None, ids.cuda())["text_embeds"].clone()
    e2 = eng.encode(None, ids.cuda())["text_embeds"].clone()
    assert torch.equal(e1, e2) and (e1.cpu() - O.clip_forward(sd, cfg, None, ids)["text_embeds"]).abs().max().item() < 5e-3
