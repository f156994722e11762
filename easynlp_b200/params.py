"""Flat parameter storage for the chinese_clip model: ONE fp32 master buffer (+ grads, Adam moments, bf16 shadow)
with named views, laid out as  [ weight-decay group | no-decay group | tensors that never receive a gradient ].

The names and shapes are the reference's checkpoint schema (modeling_chineseclip.py:219-233,316;
modeling_bert.py:72-129,145-147,264-346,535-541); the decay grouping is the substring rule of
easynlp/core/optimizers.py:490,519-523.  The layout lets clip+AdamW run as one multi-tensor kernel per group and
keeps BERT's separate query/key/value matrices adjacent so the QKV projection is one [3H, H] GEMM operand.
"""
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")
ALIGN = 64  # elements: every tensor starts on a 128-B line in the bf16 shadow (TMA rows = whole sectors) and 256 B in fp32


def uses_weight_decay(name: str) -> bool:
    return not any(nd in name for nd in NO_DECAY)


def param_schema(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """(name -> shape) in checkpoint order for model_type == chinese_clip with a ViT visual tower."""
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]; E = cfg["embed_dim"]
    n_tok = (cfg["image_resolution"] // P) ** 2 + 1
    H = cfg["text_hidden_size"]; I = cfg["text_intermediate_size"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["text_projection"] = (H, E)
    s["logit_scale"] = ()
    s["visual.class_embedding"] = (W,)
    s["visual.positional_embedding"] = (n_tok, W)
    s["visual.proj"] = (W, E)
    s["visual.conv1.weight"] = (W, 3, P, P)
    s["visual.ln_pre.weight"] = (W,); s["visual.ln_pre.bias"] = (W,)
    for i in range(cfg["vision_layers"]):
        p = f"visual.transformer.resblocks.{i}."
        s[p + "attn.in_proj_weight"] = (3 * W, W); s[p + "attn.in_proj_bias"] = (3 * W,)
        s[p + "attn.out_proj.weight"] = (W, W); s[p + "attn.out_proj.bias"] = (W,)
        s[p + "ln_1.weight"] = (W,); s[p + "ln_1.bias"] = (W,)
        s[p + "mlp.c_fc.weight"] = (4 * W, W); s[p + "mlp.c_fc.bias"] = (4 * W,)
        s[p + "mlp.c_proj.weight"] = (W, 4 * W); s[p + "mlp.c_proj.bias"] = (W,)
        s[p + "ln_2.weight"] = (W,); s[p + "ln_2.bias"] = (W,)
    s["visual.ln_post.weight"] = (W,); s["visual.ln_post.bias"] = (W,)
    s["bert.embeddings.word_embeddings.weight"] = (cfg["vocab_size"], H)
    s["bert.embeddings.position_embeddings.weight"] = (cfg["text_max_position_embeddings"], H)
    s["bert.embeddings.token_type_embeddings.weight"] = (cfg["text_type_vocab_size"], H)
    s["bert.embeddings.LayerNorm.weight"] = (H,); s["bert.embeddings.LayerNorm.bias"] = (H,)
    for i in range(cfg["text_num_hidden_layers"]):
        p = f"bert.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            s[p + f"attention.self.{nm}.weight"] = (H, H)
        for nm in ("query", "key", "value"):
            s[p + f"attention.self.{nm}.bias"] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H); s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,); s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H); s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I); s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,); s[p + "output.LayerNorm.bias"] = (H,)
    s["bert.pooler.dense.weight"] = (H, H); s["bert.pooler.dense.bias"] = (H,)
    return s


def hf_param_schema(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """(name -> shape) for the huggingface_clip branch (appzoo/clip/model.py:82-104): `text_encoder.*` = RobertaModel
    (modelzoo/models/roberta/modeling_roberta.py), `vision_encoder.*` = CLIPVisionModel (modelzoo/models/clip/modeling_clip.py:112-140,
    173-334,731-838), biased `text_projection` / `vision_projection` Linears and `logit_scale` [1].  cfg = the flat engine config built by
    easynlp_b200.engine.hf_engine_config.  q/k/v projection matrices (and biases) are adjacent so that the QKV projection is one operand."""
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]; E = cfg["embed_dim"]; Iv = cfg["vision_intermediate_size"]
    n_tok = (cfg["image_resolution"] // P) ** 2 + 1
    H = cfg["text_hidden_size"]; I = cfg["text_intermediate_size"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["logit_scale"] = (1,)
    s["text_projection.weight"] = (E, H); s["text_projection.bias"] = (E,)
    s["vision_projection.weight"] = (E, W); s["vision_projection.bias"] = (E,)
    v = "vision_encoder.vision_model."
    s[v + "embeddings.class_embedding"] = (W,)
    s[v + "embeddings.patch_embedding.weight"] = (W, 3, P, P)
    s[v + "embeddings.position_embedding.weight"] = (n_tok, W)
    s[v + "pre_layrnorm.weight"] = (W,); s[v + "pre_layrnorm.bias"] = (W,)
    for i in range(cfg["vision_layers"]):
        p = v + f"encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj"):
            s[p + f"self_attn.{nm}.weight"] = (W, W)
        for nm in ("q_proj", "k_proj", "v_proj"):
            s[p + f"self_attn.{nm}.bias"] = (W,)
        s[p + "self_attn.out_proj.weight"] = (W, W); s[p + "self_attn.out_proj.bias"] = (W,)
        s[p + "layer_norm1.weight"] = (W,); s[p + "layer_norm1.bias"] = (W,)
        s[p + "mlp.fc1.weight"] = (Iv, W); s[p + "mlp.fc1.bias"] = (Iv,)
        s[p + "mlp.fc2.weight"] = (W, Iv); s[p + "mlp.fc2.bias"] = (W,)
        s[p + "layer_norm2.weight"] = (W,); s[p + "layer_norm2.bias"] = (W,)
    s[v + "post_layernorm.weight"] = (W,); s[v + "post_layernorm.bias"] = (W,)
    t = "text_encoder."
    s[t + "embeddings.word_embeddings.weight"] = (cfg["vocab_size"], H)
    s[t + "embeddings.position_embeddings.weight"] = (cfg["text_max_position_embeddings"], H)
    s[t + "embeddings.token_type_embeddings.weight"] = (cfg["text_type_vocab_size"], H)
    s[t + "embeddings.LayerNorm.weight"] = (H,); s[t + "embeddings.LayerNorm.bias"] = (H,)
    for i in range(cfg["text_num_hidden_layers"]):
        p = t + f"encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            s[p + f"attention.self.{nm}.weight"] = (H, H)
        for nm in ("query", "key", "value"):
            s[p + f"attention.self.{nm}.bias"] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H); s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,); s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H); s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I); s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,); s[p + "output.LayerNorm.bias"] = (H,)
    s[t + "pooler.dense.weight"] = (H, H); s[t + "pooler.dense.bias"] = (H,)
    return s


def openclip_param_schema(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """(name -> shape) of model_type == open_clip (OPEN_CLIP, modelzoo/models/clip/modeling_openclip.py:255-312): the same
    VisualTransformer under `visual.*` and a pre-LN text Transformer with a causal mask under `transformer.*`; checkpoint keys carry the
    `open_clip.` prefix (appzoo/clip/model.py:61-62)."""
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]; E = cfg["embed_dim"]
    n_tok = (cfg["image_resolution"] // P) ** 2 + 1
    Wt = cfg["transformer_width"]; Lc = cfg["context_length"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["positional_embedding"] = (Lc, Wt)
    s["text_projection"] = (Wt, E)
    s["logit_scale"] = ()
    s["visual.class_embedding"] = (W,)
    s["visual.positional_embedding"] = (n_tok, W)
    s["visual.proj"] = (W, E)
    s["visual.conv1.weight"] = (W, 3, P, P)
    s["visual.ln_pre.weight"] = (W,); s["visual.ln_pre.bias"] = (W,)
    for pre, n, w in (("visual.transformer.resblocks.", cfg["vision_layers"], W), ("transformer.resblocks.", cfg["transformer_layers"], Wt)):
        for i in range(n):
            p = f"{pre}{i}."
            s[p + "attn.in_proj_weight"] = (3 * w, w); s[p + "attn.in_proj_bias"] = (3 * w,)
            s[p + "attn.out_proj.weight"] = (w, w); s[p + "attn.out_proj.bias"] = (w,)
            s[p + "ln_1.weight"] = (w,); s[p + "ln_1.bias"] = (w,)
            s[p + "mlp.c_fc.weight"] = (4 * w, w); s[p + "mlp.c_fc.bias"] = (4 * w,)
            s[p + "mlp.c_proj.weight"] = (w, 4 * w); s[p + "mlp.c_proj.bias"] = (w,)
            s[p + "ln_2.weight"] = (w,); s[p + "ln_2.bias"] = (w,)
    s["visual.ln_post.weight"] = (W,); s["visual.ln_post.bias"] = (W,)
    s["token_embedding.weight"] = (cfg["vocab_size"], Wt)
    s["ln_final.weight"] = (Wt,); s["ln_final.bias"] = (Wt,)
    return s


def wukong_param_schema(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """(name -> shape) of WukongModel (modelzoo/models/wukong/modeling_wukong.py:268-289,311-336,363-380) without the `model.` prefix
    the application's state dict carries (appzoo/wukong_clip/model.py:55): a VisualTransformer under `visual_encoder.*` and a causal
    pre-LN TextTransformer under `text_encoder.*` whose token table is the bare parameter `embedding_table`."""
    W = cfg["vision_width"]; P = cfg["vision_patch_size"]; E = cfg["embed_dim"]
    n_tok = (cfg["image_resolution"] // P) ** 2 + 1
    Wt = cfg["transformer_width"]; Lc = cfg["context_length"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["logit_scale"] = ()
    v = "visual_encoder."; t = "text_encoder."
    s[v + "class_embedding"] = (W,)
    s[v + "positional_embedding"] = (n_tok, W)
    s[v + "proj"] = (W, E)
    s[v + "conv1.weight"] = (W, 3, P, P)
    s[v + "ln_pre.weight"] = (W,); s[v + "ln_pre.bias"] = (W,)
    for pre, n, w in ((v + "transformer.resblocks.", cfg["vision_layers"], W), (t + "transformer.resblocks.", cfg["transformer_layers"], Wt)):
        for i in range(n):
            p = f"{pre}{i}."
            s[p + "attn.in_proj_weight"] = (3 * w, w); s[p + "attn.in_proj_bias"] = (3 * w,)
            s[p + "attn.out_proj.weight"] = (w, w); s[p + "attn.out_proj.bias"] = (w,)
            s[p + "ln_1.weight"] = (w,); s[p + "ln_1.bias"] = (w,)
            s[p + "mlp.c_fc.weight"] = (4 * w, w); s[p + "mlp.c_fc.bias"] = (4 * w,)
            s[p + "mlp.c_proj.weight"] = (w, 4 * w); s[p + "mlp.c_proj.bias"] = (w,)
            s[p + "ln_2.weight"] = (w,); s[p + "ln_2.bias"] = (w,)
    s[v + "ln_post.weight"] = (W,); s[v + "ln_post.bias"] = (W,)
    s[t + "text_projection"] = (Wt, E)
    s[t + "ln_final.weight"] = (Wt,); s[t + "ln_final.bias"] = (Wt,)
    s[t + "embedding_table"] = (cfg["vocab_size"], Wt)
    s[t + "positional_embedding"] = (Lc, Wt)
    return s


# the pooler is computed-but-unused by chinese_clip (modeling_chineseclip.py:349 takes [0]); its parameters never get a
# gradient, so the reference optimizer skips them (optimizers.py:420-421) -- they sit outside the updated range.
NO_GRAD = ("bert.pooler.dense.weight", "bert.pooler.dense.bias")


def _numel(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def _pad(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


class ParamStore:
    def __init__(self, cfg: dict, device="cuda", with_optimizer_state: bool = True):
        self.cfg = cfg
        self.kind = cfg.get("model_type", "chinese_clip")
        if self.kind == "huggingface_clip":
            # the image tower is frozen by the reference's `.detach()` (appzoo/clip/model.py:142): its parameters never get a gradient
            self.schema = hf_param_schema(cfg)
            self.no_grad = tuple(n for n in self.schema if n.startswith("vision_encoder."))
            self.buffers = {"text_encoder.embeddings.position_ids": cfg["text_max_position_embeddings"],
                            "vision_encoder.vision_model.embeddings.position_ids": (cfg["image_resolution"] // cfg["vision_patch_size"]) ** 2 + 1}
        elif self.kind in ("open_clip", "wukong"):
            self.schema = openclip_param_schema(cfg) if self.kind == "open_clip" else wukong_param_schema(cfg)
            self.no_grad = ()
            self.buffers = {}
        else:
            self.schema = param_schema(cfg)
            self.no_grad = NO_GRAD
        if self.kind not in ("huggingface_clip", "open_clip", "wukong"):
            # buffer exported by the reference BertEmbeddings (modeling_bert.py:87)
            self.buffers = {"bert.embeddings.position_ids": cfg["text_max_position_embeddings"]}
        NO_GRAD_ = set(self.no_grad)
        decay = [n for n in self.schema if n not in NO_GRAD_ and uses_weight_decay(n)]
        nodecay = [n for n in self.schema if n not in NO_GRAD_ and not uses_weight_decay(n)]
        frozen = [n for n in self.schema if n in NO_GRAD_]
        self.offsets: Dict[str, int] = {}
        off = 0
        for group in (decay, nodecay, frozen):
            for n in group:
                self.offsets[n] = off
                off += _pad(max(1, _numel(self.schema[n])))
            if group is decay:
                self.n_decay = off
            elif group is nodecay:
                self.n_trainable = off
        self.n_total = off
        self.device = torch.device(device)
        self.master = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        self.shadow = torch.zeros(self.n_total, dtype=torch.bfloat16, device=self.device)
        self.exp_avg = self.exp_avg_sq = None
        if with_optimizer_state:
            self.exp_avg = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
            self.exp_avg_sq = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        self.step = 0
        self.version = 0          # bumped whenever the master weights change (load / optimizer step): derived operand caches key on it

    # ---- views
    def _view(self, buf, name, shape=None, numel=None):
        shape = self.schema[name] if shape is None else shape
        n = _numel(shape) if numel is None else numel
        o = self.offsets[name]
        return buf[o:o + n].view(shape)

    def p(self, name, shape=None):
        return self._view(self.master, name, shape)

    def w(self, name, shape=None):
        """bf16 shadow copy (GEMM operand)."""
        return self._view(self.shadow, name, shape)

    def g(self, name, shape=None):
        return self._view(self.grad, name, shape)

    def m(self, name, shape=None):
        """Adam first moment of one tensor (view into the flat buffer)."""
        return self._view(self.exp_avg, name, shape)

    def v(self, name, shape=None):
        return self._view(self.exp_avg_sq, name, shape)

    def names(self) -> List[str]:
        return list(self.schema.keys())

    def trainable_names(self) -> List[str]:
        ng = set(self.no_grad)
        return [n for n in self.schema if n not in ng]

    def ranges_for(self, prefix: str) -> List[Tuple[int, int]]:
        """Contiguous [start, end) element ranges of the flat gradient covering every trainable tensor whose name starts with
        `prefix` (padding up to the next tensor included).  One layer's tensors are adjacent inside each decay group, so a layer is
        at most two ranges -- what the overlapped gradient all-reduce sends as soon as that layer's backward has run."""
        spans = sorted((self.offsets[n], self.offsets[n] + _pad(max(1, _numel(self.schema[n])))) for n in self.trainable_names()
                       if n.startswith(prefix))
        out: List[Tuple[int, int]] = []
        for a, b in spans:
            if out and out[-1][1] == a:
                out[-1] = (out[-1][0], b)
            else:
                out.append((a, b))
        return out

    # ---- checkpoint I/O (reference key names; `chinese_clip.` prefix handled by the caller)
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        missing = []
        for n, shape in self.schema.items():
            if n not in sd:
                missing.append(n)
                continue
            t = sd[n]
            if n == "logit_scale" and t.numel() == 1:
                t = t.reshape(shape)          # scalar in chinese_clip checkpoints, [1] in huggingface_clip ones
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"shape mismatch for {n}: checkpoint {tuple(t.shape)} vs model {tuple(shape)}")
            self.p(n).copy_(t.to(device=self.device, dtype=torch.float32))
        if strict and missing:
            raise KeyError(f"missing keys: {missing[:5]}...")
        self.version += 1
        self.refresh_shadow()
        return missing

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for n in self.schema:
            out[n] = self.p(n).detach().clone()
        for name, n in self.buffers.items():
            out[name] = torch.arange(n, device=self.device).unsqueeze(0)
        return out

    def refresh_shadow(self):
        from . import ops
        ops.cast_bf16(self.master, self.shadow)
