"""GPU parity of the whole CUDA path (ClipEngine over the C ABI) against
  (a) the committed golden fixtures produced by the UNMODIFIED reference (tests/golden/*.npz, oracle/make_golden.py), and
  (b) the CPU oracle on fresh seeded inputs.
Tolerances (bf16 tensor-core operands with fp32 accumulation vs the reference's fp32 PyTorch path) are written next to
each assertion; logits/loss are held to the north-star rtol 1e-3 (relative to the logit scale for logits)."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from easynlp_b200.engine import ClipEngine  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def max_err(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def oracle_grads(sd, cfg, pixels, ids, autocast=False):
    """fp32 oracle (== reference, see tests/golden) or the same graph under PyTorch's CPU bf16 autocast: the error of the
    latter against the former is the yardstick every bf16 tensor-core implementation (incl. the reference on a GPU with
    autocast(bfloat16)) is subject to."""
    names = O.trainable_names(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(params)
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = O.clip_forward(full, cfg, pixels, ids)
        out = {k: v.float() for k, v in out.items()}
    else:
        out = O.clip_forward(full, cfg, pixels, ids)
    loss = O.clip_loss(out["logits_per_text"])
    g = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    return loss.item(), {k: v.detach() for k, v in out.items()}, {k: v for k, v in zip(names, g) if v is not None}


def check_grads(eng, ref_grads, yard_grads, what):
    """per-tensor Frobenius error <= max(3 %, 1.5 x the PyTorch-bf16 yardstick) of the reference norm, plus a floor of
    1e-4 of the global gradient norm for tensors whose true gradient is ~0 (softmax is invariant to the key bias)."""
    gnorm = math.sqrt(sum(float(v.double().norm()) ** 2 for v in ref_grads.values()))
    worst = (0.0, None)
    for k, ref in ref_grads.items():
        got = eng.params.g(k).detach().float().cpu()
        err = (got - ref).norm().item()
        yard = (yard_grads[k].float() - ref).norm().item() if yard_grads is not None else 0.0
        tol = max(0.03 * ref.norm().item(), 1.5 * yard) + 1e-4 * gnorm
        if err / (ref.norm().item() + 1e-4 * gnorm) > worst[0]:
            worst = (err / (ref.norm().item() + 1e-4 * gnorm), k)
        assert err <= tol, f"{what}: grad {k}: err {err:.3e} > tol {tol:.3e} (|ref| {ref.norm().item():.3e}, bf16 yardstick {yard:.3e})"
    print(f"PARITY {what}: worst per-tensor relative gradient error {worst[0]:.3e} ({worst[1]})")


def load_tiny():
    z = np.load(os.path.join(GOLD, "tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    return z, cfg, sd


def test_tiny_forward_backward_step_vs_reference_golden():
    z, cfg, sd = load_tiny()
    eng = ClipEngine(cfg)
    eng.params.load_state_dict(sd)
    pixels = torch.from_numpy(z["pixels"]); ids = torch.from_numpy(z["ids"])
    out = eng.forward(pixels.cuda(), ids.cuda())
    torch.cuda.synchronize()
    scale = math.exp(float(sd["logit_scale"]))
    ref_img = torch.from_numpy(z["out.image_embeds"]); ref_txt = torch.from_numpy(z["out.text_embeds"]); ref_log = torch.from_numpy(z["out.logits_per_text"])
    e_img = max_err(out["image_embeds"], ref_img); e_txt = max_err(out["text_embeds"], ref_txt); e_log = max_err(out["logits_per_text"], ref_log)
    loss_ref = float(z["out.loss"]); loss = out["loss"].item()
    yl, yo, yg = oracle_grads(sd, cfg, pixels, ids, autocast=True)
    y_img = max_err(yo["image_embeds"], ref_img); y_txt = max_err(yo["text_embeds"], ref_txt); y_log = max_err(yo["logits_per_text"], ref_log)
    print(f"PARITY tiny fwd: embeds max err img {e_img:.2e} txt {e_txt:.2e}; logits max err {e_log:.2e} (scale {scale:.1f}); loss {loss:.6f} vs {loss_ref:.6f}")
    print(f"PARITY tiny fwd (PyTorch bf16 autocast yardstick): img {y_img:.2e} txt {y_txt:.2e} logits {y_log:.2e} loss {yl:.6f}")
    # 6 pairs through 2-layer towers with 3x-sharpened attention: bf16 noise is not averaged out -> yardstick-relative bounds
    assert e_img < 1.5 * y_img + 1e-4 and e_txt < 1.5 * y_txt + 1e-4
    assert e_log < 1.5 * y_log + 1e-3 and e_log < 5e-3 * scale       # and never more than 0.5 % of the logit scale
    assert abs(loss - loss_ref) < max(1e-3 * abs(loss_ref), 2.0 * abs(yl - loss_ref))   # 6-pair batch: within 2x of PyTorch-bf16's own loss error
    eng.zero_grad()
    eng.backward()
    torch.cuda.synchronize()
    ref_grads = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    check_grads(eng, ref_grads, yg, "tiny bwd vs reference golden")
    gn_ref = float(z["out.grad_norm"])
    eng.optimizer_step(lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    torch.cuda.synchronize()
    yn = math.sqrt(sum(float(v.double().norm()) ** 2 for v in yg.values()))
    # 5 %: this fixture sharpens attention 3x, where flash-style backward (D = rowsum(dO o O) from the bf16 O, as in
    # flash-attention) loses digits to cancellation in P o (dP - D); the ViT-B/16 test below holds 3 %
    gn_tol = max(5e-2, 1.5 * abs(yn - gn_ref) / gn_ref)
    print(f"PARITY tiny grad norm: {eng.norm_and_coef[0].item():.4f} vs {gn_ref:.4f} (PyTorch bf16: {yn:.4f})")
    assert abs(eng.norm_and_coef[0].item() - gn_ref) < gn_tol * gn_ref
    worst = 0.0
    for k in z.files:
        if not k.startswith("a."):
            continue
        name = k[2:]
        after_ref = torch.from_numpy(z[k]); before = sd[name]
        got = eng.params.p(name).cpu()
        # Adam's first step moves every element by ~lr regardless of gradient magnitude: compare the UPDATE
        du_ref = after_ref - before; du = got - before
        if du_ref.abs().max() == 0:
            assert du.abs().max() == 0, name      # pooler: no gradient -> untouched
            continue
        assert du.abs().max().item() <= 1.0011e-3 * (1 + before.abs().max().item()), name   # |update| <= lr (+ decay)
        # first Adam step = -lr * sign(g) wherever |g| >> eps: the update must agree wherever the reference gradient is
        # clearly non-zero (|g| > 1e-3 after clipping scale)
        g = torch.from_numpy(z["g." + name]) if ("g." + name) in z.files else None
        if g is not None:
            coef = min(1.0, 1.0 / (gn_ref + 1e-6))
            # |clipped g| >= 100 eps (update = -lr*sign(g) to within 1 %) and large within its tensor (bf16 noise cannot flip it)
            strong = ((g * coef).abs() > 1e-4) & (g.abs() > 0.25 * g.abs().max())
            if strong.any():
                err = (du - du_ref)[strong].abs().max().item()
                worst = max(worst, err / 1e-3)
                assert err < 1e-4, f"update {name}: max err {err:.3e} on well-conditioned elements"   # < 0.1 lr
    print(f"PARITY tiny step: worst update error on well-conditioned elements {worst:.4f} lr")
    # the bf16 shadow must track the master weights after the step
    for name in ("visual.proj", "bert.encoder.layer.0.intermediate.dense.weight"):
        assert max_err(eng.params.w(name), eng.params.p(name)) < 1e-2


def test_tiny_backward_vs_oracle_fresh_inputs():
    """Same model, new seeded batch (ragged lengths incl. a 2-token text): CUDA grads vs oracle autograd."""
    z, cfg, sd = load_tiny()
    eng = ClipEngine(cfg)
    eng.params.load_state_dict(sd)
    pixels, ids = O.synthetic_batch(cfg, 10, seq_len=24, seed=77)
    ids[3, 2:] = 0
    out = eng.forward(pixels.cuda(), ids.cuda())
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    l32, o32, g32 = oracle_grads(sd, cfg, pixels, ids)
    l16, o16, g16 = oracle_grads(sd, cfg, pixels, ids, autocast=True)
    assert abs(out["loss"].item() - l32) < max(1e-3 * l32, 2.0 * abs(l16 - l32))
    check_grads(eng, g32, g16, "tiny bwd vs oracle (fresh ragged batch)")


def test_b16_forward_vs_reference_golden():
    """ViT-B/16 + BERT-base, B=8, Lt=77 (BASELINE config 1 shape): weights regenerated from the seed."""
    z = np.load(os.path.join(GOLD, "b16_fwd.npz"))
    cfg = dict(O.vit_b16_bert_base_config(), text_attention_probs_dropout_prob=0.0, text_hidden_dropout_prob=0.0)
    sd = O.init_state_dict(cfg, seed=1234, scale_boost=2.0)
    assert abs(sum(v.double().sum().item() for v in sd.values()) - float(z["weights_checksum"])) < 1e-6 * abs(float(z["weights_checksum"])) + 1e-3
    pixels, ids = O.synthetic_batch(cfg, 8, seq_len=77, seed=1234)
    assert np.array_equal(ids.numpy(), z["ids"])
    eng = ClipEngine(cfg, with_optimizer_state=False)
    eng.params.load_state_dict(sd)
    out = eng.forward(pixels.cuda(), ids.cuda())
    torch.cuda.synchronize()
    scale = 1 / 0.07
    e_img = max_err(out["image_embeds"], torch.from_numpy(z["out.image_embeds"]))
    e_txt = max_err(out["text_embeds"], torch.from_numpy(z["out.text_embeds"]))
    e_log = max_err(out["logits_per_text"], torch.from_numpy(z["out.logits_per_text"]))
    loss_ref = float(z["out.loss"]); loss = out["loss"].item()
    print(f"PARITY b16 fwd: embeds max err img {e_img:.2e} txt {e_txt:.2e}; logits max err {e_log:.2e}; loss {loss:.6f} vs {loss_ref:.6f}")
    # Like-for-like yardstick: the SAME inputs through PyTorch's own bf16 autocast path (what the reference runs on a GPU
    # with autocast(bfloat16)), evaluated here with the CPU oracle.  Measured in the build container: image embeds max err
    # 4.2e-3, text 1.5e-3, logits 5.6e-2, loss rtol 6.6e-4 -- the fp32 reference cannot be matched more closely than that by
    # ANY bf16 tensor-core path, so the tolerances below are "no worse than 1.5x PyTorch-bf16", with the loss at rtol 1e-3.
    with torch.no_grad():
        ref32 = O.clip_forward(sd, cfg, pixels, ids)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ac = O.clip_forward(sd, cfg, pixels, ids)
    ac_img = max_err(ac["image_embeds"], ref32["image_embeds"]); ac_txt = max_err(ac["text_embeds"], ref32["text_embeds"])
    ac_log = max_err(ac["logits_per_text"], ref32["logits_per_text"])
    print(f"PARITY b16 fwd (PyTorch bf16 autocast yardstick): img {ac_img:.2e} txt {ac_txt:.2e} logits {ac_log:.2e}")
    assert e_img < 1.5 * ac_img + 1e-4 and e_txt < 1.5 * ac_txt + 1e-4
    assert e_log < 1.5 * ac_log + 1e-3
    assert e_log < 5e-3 * scale                           # and never more than 0.5 % of the logit scale (14.3)
    assert abs(loss - loss_ref) < 1e-3 * abs(loss_ref)    # loss rtol 1e-3 (north star)
    # backward: gradient slices pinned by the reference (+ the PyTorch-bf16 yardstick for the same slices)
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    gn = eng.params.grad.double().norm().item()
    gn_ref = float(z["out.grad_norm"])
    assert abs(gn - gn_ref) < 3e-2 * gn_ref
    _, _, yg = oracle_grads(sd, cfg, pixels, ids, autocast=True)
    worst = (0.0, None)
    for k in z.files:
        if not k.startswith("g."):
            continue
        name = k[2:]
        ref = torch.from_numpy(z[k])
        if name.endswith("]"):
            base, sl = name.rsplit("[:", 1)
            got = eng.params.g(base)[: int(sl[:-1])]; yard = yg[base][: int(sl[:-1])]
        else:
            got = eng.params.g(name); yard = yg[name]
        err = (got.detach().float().cpu() - ref).norm().item()
        yerr = (yard.float() - ref).norm().item()
        tol = max(0.03 * ref.norm().item(), 1.5 * yerr) + 1e-5 * gn_ref
        worst = max(worst, (err / (ref.norm().item() + 1e-5 * gn_ref), name))
        assert err <= tol, f"grad {name}: err {err:.3e} > tol {tol:.3e} (|ref| {ref.norm().item():.3e}, yardstick {yerr:.3e})"
    print(f"PARITY b16 bwd: grad norm {gn:.5f} vs {gn_ref:.5f}; worst pinned-slice relative error {worst[0]:.3e} ({worst[1]})")


def test_encode_matches_forward_and_recall_exact():
    z, cfg, sd = load_tiny()
    eng = ClipEngine(cfg, with_optimizer_state=False)
    eng.params.load_state_dict(sd)
    pixels, ids = O.synthetic_batch(cfg, 12, seq_len=16, seed=5)
    f = eng.forward(pixels.cuda(), ids.cuda(), save=False)
    a_img = f["image_embeds"].clone(); a_txt = f["text_embeds"].clone()
    e = eng.encode(pixels.cuda(), ids.cuda())
    assert torch.equal(e["image_embeds"], a_img) and torch.equal(e["text_embeds"], a_txt)
