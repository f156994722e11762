"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference)
in the build container.  Test infrastructure only; never imported by the product.

Run:  python oracle/make_golden.py            (needs /root/reference; CPU only)

What it pins (all fp32, dropout 0):
  tiny_fwd_bwd.npz   tiny config (oracle.tiny_config, seed 7, B=6, Lt=16):
                     inputs, every weight, reference embeds / logits / loss, every
                     reference gradient, and the weights after ONE reference step
                     (clip_grad_norm_ 1.0 + easynlp.core.optimizers.AdamW, lr 1e-3, wd 1e-4).
  b16_fwd.npz        ViT-B/16 + BERT-base (seed 1234, B=8, Lt=77): reference embeds,
                     logits, loss, grad-norm and a few gradient slices.  Weights are NOT
                     stored (755 MB) -- they are regenerated from the seed by
                     oracle.init_state_dict (torch CPU generator, deterministic).
  recall.npz         scaled random embeddings + CLIPEvaluator-style hit counts computed
                     by the reference evaluator loop.

The reference cannot be imported through ``easynlp.appzoo`` (its __init__ eagerly imports
every app) so bare namespace modules are pre-seeded for ``easynlp.appzoo`` and
``easynlp.appzoo.clip`` -- the reference files themselves are untouched.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import clip_oracle as O  # noqa: E402


def import_reference():
    os.environ.setdefault("HOME", "/root")
    sys.path.insert(0, REF)
    for name, sub in (("easynlp.appzoo", "easynlp/appzoo"), ("easynlp.appzoo.clip", "easynlp/appzoo/clip")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    from easynlp.appzoo.clip.model import CLIPApp
    from easynlp.core.optimizers import AdamW
    return CLIPApp, AdamW


def write_checkpoint(dirname, cfg, sd):
    with open(os.path.join(dirname, "config.json"), "w") as f:
        json.dump(cfg, f)
    torch.save({"chinese_clip." + k: v for k, v in sd.items()}, os.path.join(dirname, "pytorch_model.bin"))
    with open(os.path.join(dirname, "vocab.txt"), "w") as f:
        f.write("[PAD]\n[UNK]\n[CLS]\n[SEP]\n")


def reference_app(CLIPApp, cfg, sd):
    with tempfile.TemporaryDirectory() as d:
        write_checkpoint(d, cfg, sd)
        app = CLIPApp(d)
    app.train()   # dropout probs are 0 in parity configs
    return app


def ref_forward_backward(app, pixels, ids):
    out = app({"pixel_values": pixels.clone(), "input_ids": ids.clone()})
    loss = app.compute_loss(out, [])["loss"]
    app.zero_grad()
    loss.backward()
    grads = {n.replace("chinese_clip.", ""): p.grad.detach().clone()
             for n, p in app.named_parameters() if p.grad is not None}
    return out, loss.detach(), grads


def check_close(name, a, b, rtol=2e-4, atol=2e-5):
    a = a.detach().float(); b = b.detach().float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    ok = torch.allclose(a, b, rtol=rtol, atol=atol + rtol * ref * 0.1)
    print(f"  {name:55s} max|d|={err:.3e}  max|ref|={ref:.3e}  {'ok' if ok else 'MISMATCH'}")
    if not ok:
        raise SystemExit(f"oracle != reference at {name}")


def tiny_case(CLIPApp, AdamW, outdir):
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg, seed=7, scale_boost=3.0)
    pixels, ids = O.synthetic_batch(cfg, 6, seq_len=16, seed=11)
    app = reference_app(CLIPApp, cfg, sd)
    out, loss, grads = ref_forward_backward(app, pixels, ids)

    # oracle vs reference
    print("tiny: oracle vs reference")
    sd_o = {k: v.clone() for k, v in sd.items()}
    st = {}
    res = O.train_step(sd_o, cfg, pixels, ids, st, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    check_close("loss", res["loss"], loss)
    check_close("logits_per_text", res["logits_per_text"], out["logits_per_text"])
    # grads before clipping: recompute reference clip to compare post-clip grads
    named = [(n, p) for n, p in app.named_parameters()]
    gn = torch.nn.utils.clip_grad_norm_([p for _, p in named], 1.0)
    check_close("grad_norm", res["grad_norm"], gn)
    for n, p in named:
        k = n.replace("chinese_clip.", "")
        if p.grad is None:
            assert k not in res["grads"], k
            continue
        check_close("grad " + k, res["grads"][k], p.grad)
    # reference optimizer step with the reference's own grouping rule (optimizers.py:519-523)
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [
        {"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 1e-4},
        {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
    ]
    opt = AdamW(groups, lr=1e-3, weight_decay=1e-4)
    opt.step()
    after = {n.replace("chinese_clip.", ""): p.detach().clone() for n, p in named}
    for k, v in after.items():
        check_close("post-step " + k, sd_o[k], v, rtol=1e-5, atol=5e-6)  # Adam: every update is ~lr = 1e-3; fp32 grad rounding moves it by < 0.5 %

    blob = {"cfg_json": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8),
            "pixels": pixels.numpy(), "ids": ids.numpy(),
            "out.image_embeds": out["image_embeds"].detach().numpy(),
            "out.text_embeds": out["text_embeds"].detach().numpy(),
            "out.logits_per_text": out["logits_per_text"].detach().numpy(),
            "out.loss": loss.numpy(), "out.grad_norm": gn.numpy()}
    for k, v in sd.items():
        blob["w." + k] = v.numpy()
    for k, v in grads.items():
        blob["g." + k] = v.numpy()          # UNclipped reference grads
    for k, v in after.items():
        blob["a." + k] = v.numpy()
    np.savez_compressed(os.path.join(outdir, "tiny_fwd_bwd.npz"), **blob)


def b16_case(CLIPApp, outdir):
    cfg = O.vit_b16_bert_base_config()
    cfg = dict(cfg, text_attention_probs_dropout_prob=0.0, text_hidden_dropout_prob=0.0)
    sd = O.init_state_dict(cfg, seed=1234, scale_boost=2.0)
    pixels, ids = O.synthetic_batch(cfg, 8, seq_len=77, seed=1234)
    app = reference_app(CLIPApp, cfg, sd)
    out, loss, grads = ref_forward_backward(app, pixels, ids)
    print("b16: oracle vs reference")
    taps = {}
    o = O.clip_forward(sd, cfg, pixels, ids, taps)
    check_close("image_embeds", o["image_embeds"], out["image_embeds"])
    check_close("text_embeds", o["text_embeds"], out["text_embeds"])
    check_close("logits_per_text", o["logits_per_text"], out["logits_per_text"])
    check_close("loss", O.clip_loss(o["logits_per_text"]), loss)
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    blob = {"out.image_embeds": out["image_embeds"].detach().numpy(),
            "out.text_embeds": out["text_embeds"].detach().numpy(),
            "out.logits_per_text": out["logits_per_text"].detach().numpy(),
            "out.loss": loss.numpy(), "out.grad_norm": gn.numpy(),
            "ids": ids.numpy(),
            "pixels_checksum": np.array(pixels.double().sum().item()),
            "weights_checksum": np.array(sum(v.double().sum().item() for v in sd.values()))}
    for k in ("visual.proj", "text_projection", "logit_scale", "visual.ln_post.weight",
              "visual.transformer.resblocks.0.attn.in_proj_bias",
              "visual.transformer.resblocks.11.mlp.c_fc.bias",
              "bert.encoder.layer.0.attention.self.query.bias",
              "bert.encoder.layer.11.output.LayerNorm.weight",
              "visual.class_embedding", "bert.embeddings.token_type_embeddings.weight"):
        blob["g." + k] = grads[k].numpy()
    blob["g.visual.conv1.weight[:8]"] = grads["visual.conv1.weight"][:8].numpy()
    blob["g.visual.transformer.resblocks.5.mlp.c_proj.weight[:4]"] = \
        grads["visual.transformer.resblocks.5.mlp.c_proj.weight"][:4].numpy()
    blob["g.bert.encoder.layer.6.intermediate.dense.weight[:4]"] = \
        grads["bert.encoder.layer.6.intermediate.dense.weight"][:4].numpy()
    # per-layer activation checksums from the oracle taps (oracle == reference asserted above)
    for k, v in taps.items():
        blob["tap." + k] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
    np.savez_compressed(os.path.join(outdir, "b16_fwd.npz"), **blob)


def recall_case(outdir):
    """Reference evaluator loop (appzoo/clip/evaluator.py:47-61) on scaled random embeddings."""
    g = torch.Generator().manual_seed(5)
    n, e = 300, 64
    img = torch.nn.functional.normalize(torch.randn(n, e, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(img + 0.9 * torch.randn(n, e, generator=g), dim=-1)
    agreement = txt @ img.t()
    r1 = r5 = r10 = 0
    for idx in range(n):
        _, ridx = torch.sort(agreement[idx].detach(), descending=True)
        if idx in ridx[:1]:
            r1 += 1
        if idx in ridx[:5]:
            r5 += 1
        if idx in ridx[:10]:
            r10 += 1
    hits = O.recall_at_k(txt, img)
    assert (hits[1], hits[5], hits[10]) == (r1, r5, r10)
    ranks = O.rank_of_match(txt, img)
    assert int((ranks < 1).sum()) == r1 and int((ranks < 5).sum()) == r5 and int((ranks < 10).sum()) == r10
    print(f"recall: r1={r1} r5={r5} r10={r10} of {n}")
    np.savez_compressed(os.path.join(outdir, "recall.npz"), image_embeds=img.numpy(), text_embeds=txt.numpy(),
                        hits=np.array([r1, r5, r10]), n=np.array(n))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    CLIPApp, AdamW = import_reference()
    tiny_case(CLIPApp, AdamW, outdir)
    recall_case(outdir)
    b16_case(CLIPApp, outdir)
    print("golden fixtures written to", outdir)


if __name__ == "__main__":
    main()
