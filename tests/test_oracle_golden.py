"""CPU: the oracle (oracle/clip_oracle.py) against the fixtures written by the UNMODIFIED reference (oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import clip_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tiny():
    z = np.load(os.path.join(GOLD, "tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    return z, cfg, sd


def test_tiny_forward_loss_grads_and_step_match_reference():
    z, cfg, sd = _tiny()
    assert cfg == O.tiny_config()
    pixels = torch.from_numpy(z["pixels"]); ids = torch.from_numpy(z["ids"])
    st = {}
    before = {k: v.clone() for k, v in sd.items()}
    res = O.train_step(sd, cfg, pixels, ids, st, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    assert torch.allclose(res["logits_per_text"], torch.from_numpy(z["out.logits_per_text"]), rtol=1e-4, atol=1e-4)
    assert abs(res["loss"].item() - float(z["out.loss"])) < 1e-5
    assert abs(res["grad_norm"].item() - float(z["out.grad_norm"])) < 1e-4 * float(z["out.grad_norm"])
    coef = min(1.0, 1.0 / (float(z["out.grad_norm"]) + 1e-6))
    n_g = 0
    for k in z.files:
        if k.startswith("g."):
            ref = torch.from_numpy(z[k]) * coef          # fixtures hold UNclipped grads; train_step clips in place
            assert torch.allclose(res["grads"][k[2:]], ref, rtol=2e-3, atol=2e-6 + 2e-4 * ref.abs().max().item()), k
            n_g += 1
    assert n_g == len(res["grads"])
    assert "bert.pooler.dense.weight" not in res["grads"]           # unused pooler: no gradient, no update
    for k in z.files:
        if k.startswith("a."):
            assert torch.allclose(sd[k[2:]], torch.from_numpy(z[k]), rtol=1e-5, atol=5e-6), k
    assert torch.equal(sd["bert.pooler.dense.weight"], before["bert.pooler.dense.weight"])


def test_tiny_init_is_seed_deterministic():
    z, cfg, sd = _tiny()
    again = O.init_state_dict(cfg, seed=7, scale_boost=3.0)
    for k, v in sd.items():
        assert torch.equal(again[k], v), k


def test_b16_forward_matches_reference():
    z = np.load(os.path.join(GOLD, "b16_fwd.npz"))
    cfg = dict(O.vit_b16_bert_base_config(), text_attention_probs_dropout_prob=0.0, text_hidden_dropout_prob=0.0)
    sd = O.init_state_dict(cfg, seed=1234, scale_boost=2.0)
    pixels, ids = O.synthetic_batch(cfg, 8, seq_len=77, seed=1234)
    assert np.array_equal(ids.numpy(), z["ids"])
    assert abs(pixels.double().sum().item() - float(z["pixels_checksum"])) < 1e-6
    taps = {}
    with torch.no_grad():
        out = O.clip_forward(sd, cfg, pixels, ids, taps)
    assert torch.allclose(out["image_embeds"], torch.from_numpy(z["out.image_embeds"]), atol=5e-6)
    assert torch.allclose(out["text_embeds"], torch.from_numpy(z["out.text_embeds"]), atol=5e-6)
    assert torch.allclose(out["logits_per_text"], torch.from_numpy(z["out.logits_per_text"]), atol=5e-5)
    assert abs(O.clip_loss(out["logits_per_text"]).item() - float(z["out.loss"])) < 1e-5
    for k, v in taps.items():
        ref = z["tap." + k]
        assert abs(v.double().abs().sum().item() - ref[1]) < 1e-4 * abs(ref[1]), k


def test_recall_matches_reference_evaluator_loop():
    z = np.load(os.path.join(GOLD, "recall.npz"))
    img = torch.from_numpy(z["image_embeds"]); txt = torch.from_numpy(z["text_embeds"])
    hits = O.recall_at_k(txt, img)
    assert [hits[1], hits[5], hits[10]] == z["hits"].tolist()
    ranks = O.rank_of_match(txt, img)
    assert [int((ranks < k).sum()) for k in (1, 5, 10)] == z["hits"].tolist()


def test_decay_grouping_and_schedule():
    # easynlp/core/optimizers.py:490,519-523 substring rule
    assert O.uses_weight_decay("chinese_clip.logit_scale")
    assert O.uses_weight_decay("chinese_clip.visual.ln_pre.weight")
    assert O.uses_weight_decay("chinese_clip.visual.class_embedding")
    assert not O.uses_weight_decay("chinese_clip.visual.ln_pre.bias")
    assert not O.uses_weight_decay("chinese_clip.visual.transformer.resblocks.0.attn.in_proj_bias")
    assert not O.uses_weight_decay("chinese_clip.bert.encoder.layer.0.output.LayerNorm.weight")
    assert O.uses_weight_decay("chinese_clip.bert.embeddings.word_embeddings.weight")
    # easynlp/core/optimizers.py:191-204
    assert O.warmup_linear_lambda(0, 10, 100) == 0.0
    assert O.warmup_linear_lambda(5, 10, 100) == 0.5
    assert O.warmup_linear_lambda(10, 10, 100) == 1.0
    assert abs(O.warmup_linear_lambda(55, 10, 100) - 0.5) < 1e-12
    assert O.warmup_linear_lambda(100, 10, 100) == 0.0


def test_hf_oracle_matches_reference_golden():
    """huggingface_clip branch (RobertaModel + frozen CLIPVisionModel + biased projections, appzoo/clip/model.py:73-104,128-144): the
    oracle restatement against the fixture written by the UNMODIFIED reference (oracle/make_golden_hf.py), forward and every gradient."""
    import json
    z = np.load(os.path.join(GOLD, "hf_tiny_fwd_bwd.npz"))
    raw = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    names = [k for k, v in sd.items() if v.is_floating_point()]
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(params)
    out = O.hf_clip_forward(full, raw, torch.from_numpy(z["pixels"]), torch.from_numpy(z["ids"]), torch.from_numpy(z["token_type_ids"]),
                            torch.from_numpy(z["attention_mask"]))
    for k in ("image_embeds", "text_embeds", "logits_per_text"):
        assert torch.allclose(out[k], torch.from_numpy(z["out." + k]), rtol=2e-4, atol=2e-5), k
    loss = O.clip_loss(out["logits_per_text"])
    assert abs(loss.item() - float(z["out.loss"])) < 1e-5
    grads = dict(zip(names, torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)))
    ref = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    assert len(ref) == 44
    for k, g in grads.items():
        if k.startswith("vision_encoder."):
            assert g is None and k not in ref, k                      # frozen by .detach()
        else:
            assert k in ref and torch.allclose(g, ref[k], rtol=2e-3, atol=2e-6), k
    # pad-aware position ids (modeling_roberta.py:1497-1510): padded tokens stay at pad_token_id, real ones count from pad + 1
    ids = torch.from_numpy(z["ids"])
    m = ids.ne(0).int()
    pos = torch.cumsum(m, 1) * m
    assert pos[3].tolist()[:3] == [1, 2, 0] and pos.max().item() == 16


def test_openclip_oracle_matches_reference_golden():
    """open_clip branch (OPEN_CLIP: ViT + causal text transformer + EOT pooling, modeling_openclip.py:255-383) against the fixture written
    by the UNMODIFIED reference (oracle/make_golden_openclip.py): forward and all 62 gradient tensors."""
    import json
    z = np.load(os.path.join(GOLD, "openclip_tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    names = list(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    out = O.openclip_forward(params, cfg, torch.from_numpy(z["pixels"]), torch.from_numpy(z["ids"]))
    for k in ("image_embeds", "text_embeds", "logits_per_text"):
        assert torch.allclose(out[k], torch.from_numpy(z["out." + k]), rtol=2e-4, atol=2e-5), k
    loss = O.clip_loss(out["logits_per_text"])
    assert abs(loss.item() - float(z["out.loss"])) < 1e-5
    grads = dict(zip(names, torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)))
    ref = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    assert len(ref) == 62
    for k, r in ref.items():
        assert torch.allclose(grads[k], r, rtol=2e-3, atol=2e-6), k


def test_wukong_oracle_matches_reference_golden():
    """wukong_clip sibling application (WukongModel: ViT + causal TextTransformer pooled at [SEP], eps 1e-7; modeling_wukong.py:234-413)
    against the fixture written by the UNMODIFIED reference (oracle/make_golden_wukong.py): features, loss and all 62 gradient tensors."""
    import json
    z = np.load(os.path.join(GOLD, "wukong_tiny.npz"))
    raw = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    names = list(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    out = O.wukong_forward(params, raw, torch.from_numpy(z["pixels"]), torch.from_numpy(z["ids"]))
    for k in ("image_features", "text_features"):
        assert torch.allclose(out[k], torch.from_numpy(z["out." + k]), rtol=2e-4, atol=2e-5), k
    loss = O.clip_loss(out["logits_per_text"])
    assert abs(loss.item() - float(z["out.loss"])) < 1e-5
    grads = dict(zip(names, torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)))
    ref = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    assert len(ref) == 62
    for k, r in ref.items():
        assert torch.allclose(grads[k], r, rtol=2e-3, atol=2e-6), k


def test_wukong_full_tokenizer_matches_reference_vectors():
    """FullTokenizer + the dataset's tokenize rule (appzoo/wukong_clip/bert_tokenizer.py:166-396, data.py:166-187): no never_split
    protection of special tokens, 200-character words, [CLS] ids[:30] [SEP] zero padded to 32"""
    import json
    from easynlp_b200.appzoo.wukong_clip.data import FullTokenizer, wukong_tokenize
    tok = FullTokenizer(os.path.join(GOLD, "tokenizer_vocab.txt"))
    cases = json.load(open(os.path.join(GOLD, "wukong_tokenizer.json"), encoding="utf-8"))["cases"]
    assert len(cases) == 10
    for c in cases:
        assert tok.tokenize(c["text"]) == c["tokens"], c["text"]
        assert wukong_tokenize(tok, c["text"])[0].tolist() == c["input_ids"], c["text"]
    assert wukong_tokenize(tok, [c["text"] for c in cases]).tolist() == [c["input_ids"] for c in cases]
