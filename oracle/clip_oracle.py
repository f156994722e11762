"""CPU oracle for the EasyNLP CLIP (chinese_clip) contrastive path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-PyTorch fp32 restatement of the
reference algorithm.  It is imported by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never
by the product path (``easynlp_b200``), which fails loudly when its CUDA
library is missing.

Parity pin: the reference ships no golden vectors for this path
(``tests/test_clip.py:65-66,141`` assert on constant strings), so the oracle is
pinned against the *reference itself* executed in the build container:
``oracle/make_golden.py`` imports ``/root/reference`` (CLIPApp, AdamW,
CLIPEvaluator), runs it on seeded inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against those fixtures.

Everything is functional over a flat ``state_dict`` whose keys are the
reference's checkpoint keys without the ``chinese_clip.``
prefix, so the same dict drives the oracle, the reference and the CUDA path.

Reference citations (paths relative to /root/reference):
  vit_forward        easynlp/modelzoo/models/clip/modeling_chineseclip.py:219-253
  residual block     easynlp/modelzoo/models/clip/modeling_chineseclip.py:170-205
  bert_forward       easynlp/modelzoo/models/bert/modeling_bert.py:72-129,132-268,320-346,792-920
  extended mask      easynlp/modelzoo/modeling_utils.py:382-440
  encode_text / norm easynlp/modelzoo/models/clip/modeling_chineseclip.py:343-365
  logits / loss      easynlp/appzoo/clip/model.py:148-164
  AdamW              easynlp/core/optimizers.py:405-464
  wd grouping/sched  easynlp/core/optimizers.py:472-539,191-204
  grad clip          easynlp/core/trainer.py:315-325
  recall@K           easynlp/appzoo/clip/evaluator.py:47-72
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config
def vit_b16_bert_base_config() -> dict:
    """BASELINE config 2: CLIP ViT-B/16 + BERT-base (keys = CHINESE_CLIP ctor args,
    modeling_chineseclip.py:256-276)."""
    return dict(
        model_type="chinese_clip", embed_dim=512,
        image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
        vocab_size=21128, text_attention_probs_dropout_prob=0.1, text_hidden_act="gelu",
        text_hidden_dropout_prob=0.1, text_hidden_size=768, text_initializer_range=0.02,
        text_intermediate_size=3072, text_max_position_embeddings=512,
        text_num_attention_heads=12, text_num_hidden_layers=12, text_type_vocab_size=2)


def vit_l14_bert_large_config() -> dict:
    """BASELINE config 4 shape: ViT-L/14 + 24-layer d=1024 text tower, E=768."""
    return dict(
        model_type="chinese_clip", embed_dim=768,
        image_resolution=224, vision_layers=24, vision_width=1024, vision_patch_size=14,
        vocab_size=21128, text_attention_probs_dropout_prob=0.1, text_hidden_act="gelu",
        text_hidden_dropout_prob=0.1, text_hidden_size=1024, text_initializer_range=0.02,
        text_intermediate_size=4096, text_max_position_embeddings=512,
        text_num_attention_heads=16, text_num_hidden_layers=24, text_type_vocab_size=2)


def tiny_config() -> dict:
    """Small shape used for committed golden fixtures (2 heads of 64 per tower)."""
    return dict(
        model_type="chinese_clip", embed_dim=128,
        image_resolution=64, vision_layers=2, vision_width=128, vision_patch_size=16,
        vocab_size=512, text_attention_probs_dropout_prob=0.0, text_hidden_act="gelu",
        text_hidden_dropout_prob=0.0, text_hidden_size=128, text_initializer_range=0.02,
        text_intermediate_size=512, text_max_position_embeddings=64,
        text_num_attention_heads=2, text_num_hidden_layers=2, text_type_vocab_size=2)


# --------------------------------------------------------------------------- init
def init_state_dict(cfg: dict, seed: int = 1234, scale_boost: float = 1.0) -> Dict[str, Tensor]:
    """Seeded random checkpoint with the reference's key names and init laws
    (modeling_chineseclip.py:219-233,316-341; BERT normal(0, initializer_range)).
    The draw order is fixed so the same seed gives the same weights everywhere
    (torch CPU generator).  LayerNorm gains/biases are perturbed so that gain and
    bias paths are exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd: Dict[str, Tensor] = {}
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]; E = cfg["embed_dim"]
    n_tok = (cfg["image_resolution"] // P) ** 2 + 1
    sc = W ** -0.5
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    sd["text_projection"] = randn(cfg["text_hidden_size"], E, std=cfg["text_hidden_size"] ** -0.5)
    sd["visual.class_embedding"] = randn(W, std=sc)
    sd["visual.positional_embedding"] = randn(n_tok, W, std=sc)
    sd["visual.proj"] = randn(W, E, std=sc)
    sd["visual.conv1.weight"] = randn(W, 3, P, P, std=(3 * P * P) ** -0.5)

    def ln(prefix, d):
        sd[prefix + ".weight"] = 1.0 + 0.1 * randn(d)
        sd[prefix + ".bias"] = 0.1 * randn(d)

    ln("visual.ln_pre", W)
    attn_std = W ** -0.5
    for i in range(cfg["vision_layers"]):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "attn.in_proj_weight"] = randn(3 * W, W, std=attn_std) * scale_boost
        sd[p + "attn.in_proj_bias"] = 0.02 * randn(3 * W)
        sd[p + "attn.out_proj.weight"] = randn(W, W, std=attn_std * (2 * cfg["vision_layers"]) ** -0.5)
        sd[p + "attn.out_proj.bias"] = 0.02 * randn(W)
        ln(p + "ln_1", W)
        sd[p + "mlp.c_fc.weight"] = randn(4 * W, W, std=(2 * W) ** -0.5)
        sd[p + "mlp.c_fc.bias"] = 0.02 * randn(4 * W)
        sd[p + "mlp.c_proj.weight"] = randn(W, 4 * W, std=attn_std * (2 * cfg["vision_layers"]) ** -0.5)
        sd[p + "mlp.c_proj.bias"] = 0.02 * randn(W)
        ln(p + "ln_2", W)
    ln("visual.ln_post", W)

    H = cfg["text_hidden_size"]; I = cfg["text_intermediate_size"]; r = cfg["text_initializer_range"]
    sd["bert.embeddings.position_ids"] = torch.arange(cfg["text_max_position_embeddings"]).unsqueeze(0)
    sd["bert.embeddings.word_embeddings.weight"] = randn(cfg["vocab_size"], H, std=r)
    sd["bert.embeddings.word_embeddings.weight"][0].zero_()          # padding_idx=0 (modeling_bert.py:77)
    sd["bert.embeddings.position_embeddings.weight"] = randn(cfg["text_max_position_embeddings"], H, std=r)
    sd["bert.embeddings.token_type_embeddings.weight"] = randn(cfg["text_type_vocab_size"], H, std=r)
    ln("bert.embeddings.LayerNorm", H)
    for i in range(cfg["text_num_hidden_layers"]):
        p = f"bert.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            sd[p + f"attention.self.{nm}.weight"] = randn(H, H, std=r) * scale_boost
            sd[p + f"attention.self.{nm}.bias"] = 0.02 * randn(H)
        sd[p + "attention.output.dense.weight"] = randn(H, H, std=r)
        sd[p + "attention.output.dense.bias"] = 0.02 * randn(H)
        ln(p + "attention.output.LayerNorm", H)
        sd[p + "intermediate.dense.weight"] = randn(I, H, std=r)
        sd[p + "intermediate.dense.bias"] = 0.02 * randn(I)
        sd[p + "output.dense.weight"] = randn(H, I, std=r)
        sd[p + "output.dense.bias"] = 0.02 * randn(H)
        ln(p + "output.LayerNorm", H)
    sd["bert.pooler.dense.weight"] = randn(H, H, std=r)
    sd["bert.pooler.dense.bias"] = torch.zeros(H)
    return sd


def synthetic_batch(cfg: dict, batch: int, seq_len: int = 77, seed: int = 1234) -> Tuple[Tensor, Tensor]:
    """SURVEY.md 8(d) synthetic inputs: pixels ~ N(0,1) fp32; ids: [CLS]=101 at
    position 0 (clamped to the vocab), uniform tokens for the first len_i ~ U{8..L}
    positions, 0-padding afterwards (exercises the ids != 0 mask)."""
    g = torch.Generator().manual_seed(seed)
    R = cfg["image_resolution"]
    pixels = torch.randn(batch, 3, R, R, generator=g)
    V = cfg["vocab_size"]
    lo = min(8, seq_len)
    lens = torch.randint(lo, seq_len + 1, (batch,), generator=g)
    ids = torch.randint(1, V, (batch, seq_len), generator=g)
    ids[:, 0] = min(101, V - 1)
    pos = torch.arange(seq_len).unsqueeze(0)
    ids = torch.where(pos < lens.unsqueeze(1), ids, torch.zeros_like(ids))
    return pixels, ids


# --------------------------------------------------------------------------- ViT
def quick_gelu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(1.702 * x)          # modeling_chineseclip.py:179-181


def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, eps).to(x.dtype)


def _mha(x: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, heads: int,
         add_mask: Optional[Tensor] = None, prob_mult: Optional[Tensor] = None) -> Tensor:
    """softmax(QK^T/sqrt(dh) + mask)V on [B, L, d] (nn.MultiheadAttention semantics,
    packed in_proj, modeling_chineseclip.py:188,198-200)."""
    B, L, d = x.shape
    dh = d // heads
    qkv = x @ w_in.t() + b_in
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(B, L, heads, dh).transpose(1, 2)
    k = k.view(B, L, heads, dh).transpose(1, 2)
    v = v.view(B, L, heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if add_mask is not None:
        s = s + add_mask
    p = s.softmax(dim=-1)
    if prob_mult is not None:           # explicit dropout multipliers (0 or 1/(1-p)) on the attention probabilities
        p = p * prob_mult
    o = (p @ v).transpose(1, 2).reshape(B, L, d)
    return o @ w_out.t() + b_out


def vit_forward(sd: Dict[str, Tensor], cfg: dict, pixels: Tensor, taps: Optional[dict] = None, pre: str = "visual.", eps: float = 1e-5) -> Tensor:
    """VisualTransformer.forward -> [B, embed_dim] (un-normalised).  `pre` / `eps`: Wukong keeps the same tower under
    `visual_encoder.` with LayerNorm eps 1e-7 (modeling_wukong.py:268-309)."""
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]
    heads = W // 64                                              # modeling_chineseclip.py:289
    x = F.conv2d(pixels, sd[pre + "conv1.weight"], stride=P)   # [B, W, g, g]
    B = x.shape[0]
    x = x.reshape(B, W, -1).permute(0, 2, 1)
    cls = sd[pre + "class_embedding"].to(x.dtype).expand(B, 1, W)
    x = torch.cat([cls, x], dim=1) + sd[pre + "positional_embedding"]
    x = _ln(x, sd[pre + "ln_pre.weight"], sd[pre + "ln_pre.bias"], eps)
    if taps is not None:
        taps["vit.ln_pre"] = x
    for i in range(cfg["vision_layers"]):
        p = f"{pre}transformer.resblocks.{i}."
        h = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        x = x + _mha(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"],
                     sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], heads)
        h = _ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps)
        h = quick_gelu(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        x = x + (h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"])
        if taps is not None:
            taps[f"vit.block{i}"] = x
    x = _ln(x[:, 0, :], sd[pre + "ln_post.weight"], sd[pre + "ln_post.bias"], eps)
    return x @ sd[pre + "proj"]


# --------------------------------------------------------------------------- BERT
def bert_forward(sd: Dict[str, Tensor], cfg: dict, ids: Tensor, taps: Optional[dict] = None,
                 drop: Optional[dict] = None) -> Tensor:
    """BertModel(text, attention_mask = ids != 0)[0] -> last_hidden_state [B, L, H].
    Dropout is the identity unless `drop` supplies explicit multiplier tensors (0 or 1/(1-p)) for the reference's nn.Dropout
    sites: drop["emb"] [B,L,H] (modeling_bert.py:128), drop[("attn", i)] [B,heads,L,L] (:238), drop[("self_out", i)] and
    drop[("out", i)] [B,L,H] (:267,345) -- used to test the fused Philox dropout of the CUDA path with ITS masks."""
    H = cfg["text_hidden_size"]; heads = cfg["text_num_attention_heads"]
    B, L = ids.shape
    eps = 1e-12                                                  # modeling_chineseclip.py:311
    x = (sd["bert.embeddings.word_embeddings.weight"][ids]
         + sd["bert.embeddings.token_type_embeddings.weight"][0]
         + sd["bert.embeddings.position_embeddings.weight"][:L])
    x = F.layer_norm(x, (H,), sd["bert.embeddings.LayerNorm.weight"], sd["bert.embeddings.LayerNorm.bias"], eps)
    drop = drop or {}
    if "emb" in drop:
        x = x * drop["emb"]
    if taps is not None:
        taps["bert.emb"] = x
    mask = (1.0 - ids.ne(0).to(x.dtype))[:, None, None, :] * -10000.0   # modeling_utils.py:438-439
    act = F.gelu if cfg["text_hidden_act"] == "gelu" else None
    assert act is not None, "only erf-GELU text towers are on the hot path"
    for i in range(cfg["text_num_hidden_layers"]):
        p = f"bert.encoder.layer.{i}."
        w_in = torch.cat([sd[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value")], 0)
        b_in = torch.cat([sd[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value")], 0)
        a = _mha(x, w_in, b_in, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"],
                 heads, mask, drop.get(("attn", i)))
        if ("self_out", i) in drop:
            a = a * drop[("self_out", i)]
        x = F.layer_norm(a + x, (H,), sd[p + "attention.output.LayerNorm.weight"],
                         sd[p + "attention.output.LayerNorm.bias"], eps)
        h = act(x @ sd[p + "intermediate.dense.weight"].t() + sd[p + "intermediate.dense.bias"])
        h = h @ sd[p + "output.dense.weight"].t() + sd[p + "output.dense.bias"]
        if ("out", i) in drop:
            h = h * drop[("out", i)]
        x = F.layer_norm(h + x, (H,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)
        if taps is not None:
            taps[f"bert.layer{i}"] = x
    return x


# --------------------------------------------------------------------------- CLIP head
def clip_forward(sd: Dict[str, Tensor], cfg: dict, pixels: Optional[Tensor], ids: Optional[Tensor],
                 taps: Optional[dict] = None, drop: Optional[dict] = None) -> dict:
    """CLIPApp.forward for model_type == chinese_clip (appzoo/clip/model.py:106-150)."""
    image_embeds = text_embeds = None
    if pixels is not None:
        f = vit_forward(sd, cfg, pixels, taps)
        image_embeds = f / f.norm(dim=-1, keepdim=True)
    if ids is not None:
        t = bert_forward(sd, cfg, ids, taps, drop)[:, 0, :] @ sd["text_projection"]
        text_embeds = t / t.norm(dim=-1, keepdim=True)
    out = {"image_embeds": image_embeds, "text_embeds": text_embeds}
    if image_embeds is not None and text_embeds is not None:
        lpt = (text_embeds @ image_embeds.t()) * sd["logit_scale"].exp()
        out["logits_per_text"] = lpt
        out["logits_per_image"] = lpt.T
    return out


def clip_loss(logits_per_text: Tensor) -> Tensor:
    """(CE(S, arange) + CE(S^T, arange)) / 2  (appzoo/clip/model.py:154-160)."""
    n = logits_per_text.shape[0]
    lab = torch.arange(n, device=logits_per_text.device)
    return (F.cross_entropy(logits_per_text, lab) + F.cross_entropy(logits_per_text.T, lab)) / 2.0


# --------------------------------------------------------------------------- optimizer
NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")      # optimizers.py:490


def uses_weight_decay(name: str) -> bool:
    """Substring rule of get_optimizer (optimizers.py:519-523): note ViT ln_*.weight,
    class/positional embeddings, proj, text_projection and logit_scale DO decay."""
    return not any(nd in name for nd in NO_DECAY)


def warmup_linear_lambda(step: int, warmup_steps: int, t_total: int) -> float:
    """EasyNLPWarmupLinearSchedule.lr_lambda (optimizers.py:191-204)."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    return max(0.0, float(t_total - step) / float(max(1.0, t_total - warmup_steps)))


def clip_grad_norm(grads: List[Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ semantics (trainer.py:325): returns the total
    norm and scales grads in place by min(1, max_norm / (norm + 1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
               weight_decay: float, beta1=0.9, beta2=0.999, eps=1e-6) -> None:
    """One in-place step of the reference's AdamW (optimizers.py:437-462):
    eps is added to sqrt(v) BEFORE bias correction; decoupled decay is applied
    after the Adam update with the scheduled lr."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def trainable_names(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k, v in sd.items() if v.is_floating_point()]


def train_step(sd: Dict[str, Tensor], cfg: dict, pixels: Tensor, ids: Tensor, opt_state: dict,
               lr: float, weight_decay: float = 1e-4, max_grad_norm: float = 1.0) -> dict:
    """forward + loss + backward + clip + AdamW over every float tensor in ``sd``
    (in place).  Parameters that receive no gradient (bert.pooler.*) are skipped,
    as the reference optimizer skips ``p.grad is None`` (optimizers.py:420-421)."""
    names = trainable_names(sd)
    params = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(params)
    out = clip_forward(full, cfg, pixels, ids)
    loss = clip_loss(out["logits_per_text"])
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    gmap = {k: g for k, g in zip(names, grads) if g is not None}
    gnorm = clip_grad_norm(list(gmap.values()), max_grad_norm)
    opt_state["step"] = opt_state.get("step", 0) + 1
    for k, g in gmap.items():
        if k not in opt_state:
            opt_state[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
        m, v = opt_state[k]
        adamw_step(sd[k], g, m, v, opt_state["step"], lr, weight_decay if uses_weight_decay(k) else 0.0)
    return {"loss": loss.detach(), "grad_norm": gnorm, "grads": gmap,
            "logits_per_text": out["logits_per_text"].detach()}


# --------------------------------------------------------------------------- retrieval
def recall_at_k(text_embeds: Tensor, image_embeds: Tensor, ks=(1, 5, 10)) -> Dict[int, int]:
    """CLIPEvaluator's text->image hit counts (evaluator.py:47-61): query idx is a hit at K
    when idx is among the first K entries of torch.sort(agreement[idx], descending=True)."""
    agreement = text_embeds @ image_embeds.t()
    hits = {k: 0 for k in ks}
    for idx in range(agreement.shape[0]):
        _, ridx = torch.sort(agreement[idx], descending=True)
        for k in ks:
            if idx in ridx[:k]:
                hits[k] += 1
    return hits


def rank_of_match(text_embeds: Tensor, image_embeds: Tensor) -> Tensor:
    """Number of gallery items scoring strictly higher than the matching item, per query
    (tie-free restatement used for large N; hit@K <=> rank < K when there are no ties)."""
    agreement = text_embeds @ image_embeds.t()
    diag = agreement.diagonal().unsqueeze(1)
    return (agreement > diag).sum(dim=1)


# =========================================================================== huggingface_clip branch
# appzoo/clip/model.py:73-104,128-144: text = RobertaModel (modelzoo/models/roberta/modeling_roberta.py:65-575: pad-aware position ids
# :1497-1510, BERT-style post-LN encoder, tanh pooler), image = CLIPVisionModel (modelzoo/models/clip/modeling_clip.py:112-140,173-334,
# 731-776) whose pooled output is DETACHED (model.py:142), biased text_projection / vision_projection Linears, logit_scale [1].
def hf_tiny_config() -> dict:
    """nested config.json of the reference's huggingface_clip branch at a small shape (2 heads of 64 per tower)"""
    return {"text_config": dict(vocab_size=512, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                                max_position_embeddings=64, hidden_act="gelu", layer_norm_eps=1e-12, pad_token_id=0, type_vocab_size=2,
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
            "vision_config": dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, image_size=64,
                                  patch_size=16, hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0),
            "projection_dim": 128}


def hf_init_state_dict(raw_cfg: dict, seed: int = 1234, scale_boost: float = 1.0) -> Dict[str, Tensor]:
    """seeded random checkpoint with the reference's huggingface_clip key names"""
    g = torch.Generator().manual_seed(seed)
    t = raw_cfg["text_config"]; v = raw_cfg["vision_config"]; E = raw_cfg.get("projection_dim", 512)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd: Dict[str, Tensor] = {}
    W = v["hidden_size"]; P = v["patch_size"]; Iv = v["intermediate_size"]; n_tok = (v["image_size"] // P) ** 2 + 1
    H = t["hidden_size"]; I = t["intermediate_size"]; r = 0.02
    sd["logit_scale"] = torch.tensor([math.log(1 / 0.07)])
    sd["text_projection.weight"] = randn(E, H, std=H ** -0.5); sd["text_projection.bias"] = 0.02 * randn(E)
    sd["vision_projection.weight"] = randn(E, W, std=W ** -0.5); sd["vision_projection.bias"] = 0.02 * randn(E)
    p = "vision_encoder.vision_model."
    sd[p + "embeddings.class_embedding"] = randn(W, std=W ** -0.5)
    sd[p + "embeddings.patch_embedding.weight"] = randn(W, 3, P, P, std=(3 * P * P) ** -0.5)
    sd[p + "embeddings.position_embedding.weight"] = randn(n_tok, W, std=W ** -0.5)
    sd[p + "embeddings.position_ids"] = torch.arange(n_tok).unsqueeze(0)

    def ln(prefix, d):
        sd[prefix + ".weight"] = 1.0 + 0.1 * randn(d); sd[prefix + ".bias"] = 0.1 * randn(d)

    ln(p + "pre_layrnorm", W)
    for i in range(v["num_hidden_layers"]):
        q = p + f"encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj"):
            sd[q + f"self_attn.{nm}.weight"] = randn(W, W, std=W ** -0.5) * scale_boost; sd[q + f"self_attn.{nm}.bias"] = 0.02 * randn(W)
        sd[q + "self_attn.out_proj.weight"] = randn(W, W, std=W ** -0.5 * (2 * v["num_hidden_layers"]) ** -0.5); sd[q + "self_attn.out_proj.bias"] = 0.02 * randn(W)
        ln(q + "layer_norm1", W)
        sd[q + "mlp.fc1.weight"] = randn(Iv, W, std=(2 * W) ** -0.5); sd[q + "mlp.fc1.bias"] = 0.02 * randn(Iv)
        sd[q + "mlp.fc2.weight"] = randn(W, Iv, std=W ** -0.5 * (2 * v["num_hidden_layers"]) ** -0.5); sd[q + "mlp.fc2.bias"] = 0.02 * randn(W)
        ln(q + "layer_norm2", W)
    ln(p + "post_layernorm", W)
    p = "text_encoder."
    sd[p + "embeddings.position_ids"] = torch.arange(t["max_position_embeddings"]).unsqueeze(0)
    sd[p + "embeddings.word_embeddings.weight"] = randn(t["vocab_size"], H, std=r); sd[p + "embeddings.word_embeddings.weight"][t["pad_token_id"]].zero_()
    sd[p + "embeddings.position_embeddings.weight"] = randn(t["max_position_embeddings"], H, std=r)
    sd[p + "embeddings.position_embeddings.weight"][t["pad_token_id"]].zero_()
    sd[p + "embeddings.token_type_embeddings.weight"] = randn(t["type_vocab_size"], H, std=r)
    ln(p + "embeddings.LayerNorm", H)
    for i in range(t["num_hidden_layers"]):
        q = p + f"encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            sd[q + f"attention.self.{nm}.weight"] = randn(H, H, std=r) * scale_boost; sd[q + f"attention.self.{nm}.bias"] = 0.02 * randn(H)
        sd[q + "attention.output.dense.weight"] = randn(H, H, std=r); sd[q + "attention.output.dense.bias"] = 0.02 * randn(H)
        ln(q + "attention.output.LayerNorm", H)
        sd[q + "intermediate.dense.weight"] = randn(I, H, std=r); sd[q + "intermediate.dense.bias"] = 0.02 * randn(I)
        sd[q + "output.dense.weight"] = randn(H, I, std=r); sd[q + "output.dense.bias"] = 0.02 * randn(H)
        ln(q + "output.LayerNorm", H)
    sd[p + "pooler.dense.weight"] = randn(H, H, std=H ** -0.5); sd[p + "pooler.dense.bias"] = 0.02 * randn(H)
    return sd


def hf_vision_forward(sd: Dict[str, Tensor], v: dict, pixels: Tensor) -> Tensor:
    """CLIPVisionTransformer.forward -> pooled = post_layernorm(last_hidden[:, 0])  (modeling_clip.py:731-776)"""
    p = "vision_encoder.vision_model."
    W = v["hidden_size"]; heads = v["num_attention_heads"]
    act = quick_gelu if v.get("hidden_act", "quick_gelu") == "quick_gelu" else F.gelu
    x = F.conv2d(pixels, sd[p + "embeddings.patch_embedding.weight"], stride=v["patch_size"])
    B = x.shape[0]
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd[p + "embeddings.class_embedding"].expand(B, 1, W), x], dim=1) + sd[p + "embeddings.position_embedding.weight"]
    x = F.layer_norm(x, (W,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], 1e-5)
    for i in range(v["num_hidden_layers"]):
        q = p + f"encoder.layers.{i}."
        h = F.layer_norm(x, (W,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], 1e-5)
        w_in = torch.cat([sd[q + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")], 0)
        b_in = torch.cat([sd[q + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")], 0)
        x = x + _mha(h, w_in, b_in, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"], heads)   # q * scale == scores / sqrt(dh)
        h = F.layer_norm(x, (W,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], 1e-5)
        h = act(h @ sd[q + "mlp.fc1.weight"].t() + sd[q + "mlp.fc1.bias"])
        x = x + (h @ sd[q + "mlp.fc2.weight"].t() + sd[q + "mlp.fc2.bias"])
    return F.layer_norm(x[:, 0, :], (W,), sd[p + "post_layernorm.weight"], sd[p + "post_layernorm.bias"], 1e-5)


def hf_text_forward(sd: Dict[str, Tensor], t: dict, ids: Tensor, token_type_ids: Optional[Tensor] = None,
                    attention_mask: Optional[Tensor] = None) -> Tensor:
    """RobertaModel(...)[1] = tanh pooler output (modeling_roberta.py:100-130,559-575,1497-1510)"""
    p = "text_encoder."
    H = t["hidden_size"]; heads = t["num_attention_heads"]; pad = t.get("pad_token_id", 0); eps = t.get("layer_norm_eps", 1e-12)
    B, L = ids.shape
    m = ids.ne(pad).int()
    pos = (torch.cumsum(m, dim=1) * m).long() + pad
    tt = token_type_ids if token_type_ids is not None else torch.zeros_like(ids)
    x = (sd[p + "embeddings.word_embeddings.weight"][ids] + sd[p + "embeddings.token_type_embeddings.weight"][tt]
         + sd[p + "embeddings.position_embeddings.weight"][pos])
    x = F.layer_norm(x, (H,), sd[p + "embeddings.LayerNorm.weight"], sd[p + "embeddings.LayerNorm.bias"], eps)
    am = attention_mask if attention_mask is not None else ids.ne(pad).long()
    mask = (1.0 - am.to(x.dtype))[:, None, None, :] * -10000.0
    for i in range(t["num_hidden_layers"]):
        q = p + f"encoder.layer.{i}."
        w_in = torch.cat([sd[q + f"attention.self.{n}.weight"] for n in ("query", "key", "value")], 0)
        b_in = torch.cat([sd[q + f"attention.self.{n}.bias"] for n in ("query", "key", "value")], 0)
        a = _mha(x, w_in, b_in, sd[q + "attention.output.dense.weight"], sd[q + "attention.output.dense.bias"], heads, mask)
        x = F.layer_norm(a + x, (H,), sd[q + "attention.output.LayerNorm.weight"], sd[q + "attention.output.LayerNorm.bias"], eps)
        h = F.gelu(x @ sd[q + "intermediate.dense.weight"].t() + sd[q + "intermediate.dense.bias"])
        h = h @ sd[q + "output.dense.weight"].t() + sd[q + "output.dense.bias"]
        x = F.layer_norm(h + x, (H,), sd[q + "output.LayerNorm.weight"], sd[q + "output.LayerNorm.bias"], eps)
    return torch.tanh(x[:, 0] @ sd[p + "pooler.dense.weight"].t() + sd[p + "pooler.dense.bias"])


def hf_clip_forward(sd: Dict[str, Tensor], raw_cfg: dict, pixels: Optional[Tensor], ids: Optional[Tensor],
                    token_type_ids: Optional[Tensor] = None, attention_mask: Optional[Tensor] = None) -> dict:
    """CLIPApp.forward, huggingface_clip branch (appzoo/clip/model.py:128-150)"""
    image_embeds = text_embeds = None
    if ids is not None:
        tfeat = hf_text_forward(sd, raw_cfg["text_config"], ids, token_type_ids, attention_mask)
        tfeat = tfeat @ sd["text_projection.weight"].t() + sd["text_projection.bias"]
        text_embeds = tfeat / tfeat.norm(dim=-1, keepdim=True)
    if pixels is not None:
        pooled = hf_vision_forward(sd, raw_cfg["vision_config"], pixels).detach()          # model.py:142
        f = pooled @ sd["vision_projection.weight"].t() + sd["vision_projection.bias"]
        image_embeds = f / f.norm(dim=-1, keepdim=True)
    out = {"image_embeds": image_embeds, "text_embeds": text_embeds}
    if image_embeds is not None and text_embeds is not None:
        lpt = (text_embeds @ image_embeds.t()) * sd["logit_scale"].exp()
        out["logits_per_text"] = lpt; out["logits_per_image"] = lpt.T
    return out


# =========================================================================== open_clip branch
# appzoo/clip/model.py:56-63: OPEN_CLIP (modelzoo/models/clip/modeling_openclip.py:255-383) = the same VisualTransformer + a pre-LN text
# Transformer with a causal mask (:346-352), token + positional embeddings, ln_final and EOT-argmax pooling (:358-371).
def openclip_tiny_config() -> dict:
    return dict(model_type="open_clip", embed_dim=128, image_resolution=64, vision_layers=2, vision_width=128, vision_patch_size=16,
                context_length=24, vocab_size=512, transformer_width=128, transformer_heads=2, transformer_layers=2)


def openclip_init_state_dict(cfg: dict, seed: int = 1234, scale_boost: float = 1.0) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd: Dict[str, Tensor] = {}
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]; E = cfg["embed_dim"]; Wt = cfg["transformer_width"]
    n_tok = (cfg["image_resolution"] // P) ** 2 + 1
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    sd["positional_embedding"] = randn(cfg["context_length"], Wt, std=0.01)
    sd["text_projection"] = randn(Wt, E, std=Wt ** -0.5)
    sd["token_embedding.weight"] = randn(cfg["vocab_size"], Wt, std=0.02)
    sd["visual.class_embedding"] = randn(W, std=W ** -0.5)
    sd["visual.positional_embedding"] = randn(n_tok, W, std=W ** -0.5)
    sd["visual.proj"] = randn(W, E, std=W ** -0.5)
    sd["visual.conv1.weight"] = randn(W, 3, P, P, std=(3 * P * P) ** -0.5)

    def ln(prefix, d):
        sd[prefix + ".weight"] = 1.0 + 0.1 * randn(d); sd[prefix + ".bias"] = 0.1 * randn(d)

    ln("visual.ln_pre", W); ln("visual.ln_post", W); ln("ln_final", Wt)
    for pre, n, w in (("visual.transformer.resblocks.", cfg["vision_layers"], W), ("transformer.resblocks.", cfg["transformer_layers"], Wt)):
        for i in range(n):
            p = f"{pre}{i}."
            sd[p + "attn.in_proj_weight"] = randn(3 * w, w, std=w ** -0.5) * scale_boost; sd[p + "attn.in_proj_bias"] = 0.02 * randn(3 * w)
            sd[p + "attn.out_proj.weight"] = randn(w, w, std=w ** -0.5 * (2 * n) ** -0.5); sd[p + "attn.out_proj.bias"] = 0.02 * randn(w)
            ln(p + "ln_1", w); ln(p + "ln_2", w)
            sd[p + "mlp.c_fc.weight"] = randn(4 * w, w, std=(2 * w) ** -0.5); sd[p + "mlp.c_fc.bias"] = 0.02 * randn(4 * w)
            sd[p + "mlp.c_proj.weight"] = randn(w, 4 * w, std=w ** -0.5 * (2 * n) ** -0.5); sd[p + "mlp.c_proj.bias"] = 0.02 * randn(w)
    return sd


def openclip_text_forward(sd: Dict[str, Tensor], cfg: dict, ids: Tensor) -> Tensor:
    """OPEN_CLIP.encode_text (modeling_openclip.py:358-371) -> [B, embed_dim] (un-normalised)"""
    Wt = cfg["transformer_width"]; heads = cfg["transformer_heads"]
    B, L = ids.shape
    x = sd["token_embedding.weight"][ids] + sd["positional_embedding"]
    causal = torch.full((L, L), float("-inf")).triu_(1)
    for i in range(cfg["transformer_layers"]):
        p = f"transformer.resblocks.{i}."
        h = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        x = x + _mha(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"], sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], heads, causal)
        h = _ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = quick_gelu(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        x = x + (h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"])
    x = _ln(x, sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(B), ids.argmax(dim=-1)] @ sd["text_projection"]


def openclip_forward(sd: Dict[str, Tensor], cfg: dict, pixels: Optional[Tensor], ids: Optional[Tensor]) -> dict:
    image_embeds = text_embeds = None
    if pixels is not None:
        f = vit_forward(sd, cfg, pixels)
        image_embeds = f / f.norm(dim=-1, keepdim=True)
    if ids is not None:
        t = openclip_text_forward(sd, cfg, ids)
        text_embeds = t / t.norm(dim=-1, keepdim=True)
    out = {"image_embeds": image_embeds, "text_embeds": text_embeds}
    if image_embeds is not None and text_embeds is not None:
        lpt = (text_embeds @ image_embeds.t()) * sd["logit_scale"].exp()
        out["logits_per_text"] = lpt; out["logits_per_image"] = lpt.T
    return out


# =========================================================================== Wukong (sibling application wukong_clip)
# appzoo/wukong_clip/model.py:22-88 + modelzoo/models/wukong/modeling_wukong.py:234-413: the same ViT under `visual_encoder.` and a causal
# pre-LN TextTransformer under `text_encoder.` (token table = the bare parameter `embedding_table`), LayerNorm eps 1e-7 everywhere,
# features pooled at the [SEP] (id 102) position (:349,359).  Keys as in the application's checkpoint minus the `model.` prefix.
def wukong_tiny_config() -> dict:
    return {"model": {"visual": dict(input_resolution=64, patch_size=16, width=128, layers=2, heads=2, output_dim=128),
                      "text": dict(context_length=32, vocab_size=512, output_dim=128, width=128, layers=2, heads=2)}}


def wukong_flat_config(raw: dict) -> dict:
    v = raw["model"]["visual"]; t = raw["model"]["text"]
    return dict(embed_dim=v["output_dim"], image_resolution=v["input_resolution"], vision_layers=v["layers"], vision_width=v["width"],
                vision_patch_size=v["patch_size"], vocab_size=t["vocab_size"], context_length=t["context_length"],
                transformer_width=t["width"], transformer_heads=t["heads"], transformer_layers=t["layers"])


def wukong_init_state_dict(raw: dict, seed: int = 1234, scale_boost: float = 1.0) -> Dict[str, Tensor]:
    """the open_clip tiny initialisation under Wukong's key names"""
    cfg = wukong_flat_config(raw)
    src = openclip_init_state_dict(cfg, seed, scale_boost)
    sd: Dict[str, Tensor] = {}
    for k, v in src.items():
        if k == "logit_scale":
            sd[k] = v
        elif k.startswith("visual."):
            sd["visual_encoder." + k[len("visual."):]] = v
        elif k == "token_embedding.weight":
            sd["text_encoder.embedding_table"] = v
        else:
            sd["text_encoder." + k] = v
    return sd


def wukong_text_forward(sd: Dict[str, Tensor], cfg: dict, ids: Tensor, sep_id: int = 102) -> Tensor:
    """TextTransformer.forward (modeling_wukong.py:347-361) -> [B, output_dim] (un-normalised); one [SEP] per row"""
    heads = cfg["transformer_heads"]; t = "text_encoder."
    B, L = ids.shape
    x = sd[t + "embedding_table"][ids] + sd[t + "positional_embedding"]
    causal = torch.full((L, L), float("-inf")).triu_(1)
    for i in range(cfg["transformer_layers"]):
        p = f"{t}transformer.resblocks.{i}."
        h = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-7)
        x = x + _mha(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"], sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], heads, causal)
        h = _ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-7)
        h = quick_gelu(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        x = x + (h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"])
    x = _ln(x, sd[t + "ln_final.weight"], sd[t + "ln_final.bias"], 1e-7)
    rows, cols = (ids == sep_id).nonzero(as_tuple=True)
    return x[rows, cols] @ sd[t + "text_projection"]


def wukong_forward(sd: Dict[str, Tensor], raw: dict, pixels: Optional[Tensor], ids: Optional[Tensor]) -> dict:
    cfg = wukong_flat_config(raw)
    image_features = text_features = None
    if pixels is not None:
        f = vit_forward(sd, cfg, pixels, pre="visual_encoder.", eps=1e-7)
        image_features = f / f.norm(p=2, dim=-1, keepdim=True)
    if ids is not None:
        tx = wukong_text_forward(sd, cfg, ids)
        text_features = tx / tx.norm(p=2, dim=-1, keepdim=True)
    out = {"image_features": image_features, "text_features": text_features, "logit_scale": sd["logit_scale"].exp()}
    if image_features is not None and text_features is not None:
        out["logits_per_text"] = out["logit_scale"] * text_features @ image_features.t()
    return out
