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


def load_tiny():
    z = np.load(os.path.join(GOLD, "tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    return z, cfg, sd


def test_tiny_forward_backward_step_vs_reference_golden():
    z, cfg, sd = load_tiny()
    eng = ClipEngine(cfg)
    eng.params.load_state_dict(sd)
    pixels = torch.from_numpy(z["pixels"]).cuda(); ids = torch.from_numpy(z["ids"]).cuda()
    out = eng.forward(pixels, ids)
    torch.cuda.synchronize()
    scale = math.exp(float(sd["logit_scale"]))
    e_img = max_err(out["image_embeds"], torch.from_numpy(z["out.image_embeds"]))
    e_txt = max_err(out["text_embeds"], torch.from_numpy(z["out.text_embeds"]))
    e_log = max_err(out["logits_per_text"], torch.from_numpy(z["out.logits_per_text"]))
    loss_ref = float(z["out.loss"]); loss = out["loss"].item()
    print(f"PARITY tiny fwd: embeds max err img {e_img:.2e} txt {e_txt:.2e}; logits max err {e_log:.2e} (scale {scale:.1f}); loss {loss:.6f} vs {loss_ref:.6f}")
    assert e_img < 5e-3 and e_txt < 5e-3                 # unit-norm embeddings (|x| ~ 0.09), bf16 towers
    assert e_log < 5e-3 * scale                          # |dlogit| <= 0.5 % of the logit scale exp(logit_scale) = 14.3
    assert abs(loss - loss_ref) < 2e-3 * abs(loss_ref)   # rtol 2e-3 (tiny model, 6 pairs: bf16 noise is not averaged)
    eng.zero_grad()
    eng.backward()
    torch.cuda.synchronize()
    worst = 0.0
    for k in z.files:
        if not k.startswith("g."):
            continue
        name = k[2:]
        ref = torch.from_numpy(z[k])
        got = eng.params.g(name)
        r = rel_err(got, ref)
        worst = max(worst, r)
        assert r < 6e-2, f"grad {name}: rel err {r:.3e}"  # bf16 backward: a few % in Frobenius norm per tensor
    print(f"PARITY tiny bwd: worst per-tensor relative grad error {worst:.3e}")
    gn_ref = float(z["out.grad_norm"])
    eng.optimizer_step(lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    torch.cuda.synchronize()
    assert abs(eng.norm_and_coef[0].item() - gn_ref) < 3e-2 * gn_ref
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
        err = (du - du_ref).abs().max().item()
        worst = max(worst, err / 1e-3)
        assert err < 1.2e-3, f"update {name}: max err {err:.3e}"   # |update| <= lr = 1e-3; sign flips of ~0 grads allowed
    print(f"PARITY tiny step: worst update error {worst:.3f} lr")
    # the bf16 shadow must track the master weights after the step
    for name in ("visual.proj", "bert.encoder.layer.0.intermediate.dense.weight"):
        assert max_err(eng.params.w(name), eng.params.p(name)) < 1e-2


def test_tiny_backward_vs_oracle_fresh_inputs():
    """Same model, new seeded batch (ragged lengths incl. a length-1-ish text): CUDA grads vs oracle autograd."""
    z, cfg, sd = load_tiny()
    eng = ClipEngine(cfg)
    eng.params.load_state_dict(sd)
    pixels, ids = O.synthetic_batch(cfg, 10, seq_len=24, seed=77)
    ids[3, 2:] = 0
    out = eng.forward(pixels.cuda(), ids.cuda())
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    names = O.trainable_names(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(params)
    ref = O.clip_forward(full, cfg, pixels, ids)
    loss = O.clip_loss(ref["logits_per_text"])
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    assert abs(out["loss"].item() - loss.item()) < 2e-3 * loss.item()
    for k, g in zip(names, grads):
        if g is None:
            continue
        r = rel_err(eng.params.g(k), g)
        assert r < 6e-2, f"grad {k}: rel err {r:.3e}"


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
    # backward: gradient slices pinned by the reference
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    gn = eng.params.grad.double().norm().item()
    assert abs(gn - float(z["out.grad_norm"])) < 3e-2 * float(z["out.grad_norm"])
    for k in z.files:
        if not k.startswith("g."):
            continue
        name = k[2:]
        ref = torch.from_numpy(z[k])
        if name.endswith("]"):
            base, sl = name.rsplit("[:", 1)
            got = eng.params.g(base)[: int(sl[:-1])]
        else:
            got = eng.params.g(name)
        r = rel_err(got, ref)
        assert r < 6e-2, f"grad {name}: rel err {r:.3e}"


def test_encode_matches_forward_and_recall_exact():
    z, cfg, sd = load_tiny()
    eng = ClipEngine(cfg, with_optimizer_state=False)
    eng.params.load_state_dict(sd)
    pixels, ids = O.synthetic_batch(cfg, 12, seq_len=16, seed=5)
    f = eng.forward(pixels.cuda(), ids.cuda(), save=False)
    a_img = f["image_embeds"].clone(); a_txt = f["text_embeds"].clone()
    e = eng.encode(pixels.cuda(), ids.cuda())
    assert torch.equal(e["image_embeds"], a_img) and torch.equal(e["text_embeds"], a_txt)
