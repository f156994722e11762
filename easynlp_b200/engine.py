"""ClipEngine: host-side schedule of the CLIP training / encoding step over the clipk kernels.

Everything numerical happens in hand-written sm_100a kernels behind the C ABI (include/clipk.h); this file only owns
buffers (torch tensors as device memory) and the order of launches, i.e. it is the from-scratch counterpart of the
autograd graph PyTorch builds for

    CLIPApp.forward / compute_loss      easynlp/appzoo/clip/model.py:106-164
    CHINESE_CLIP.forward                easynlp/modelzoo/models/clip/modeling_chineseclip.py:343-365
    VisualTransformer / ResidualAttentionBlock   .../modeling_chineseclip.py:170-253   (pre-LN, QuickGELU)
    BertModel (embeddings + post-LN encoder)     easynlp/modelzoo/models/bert/modeling_bert.py:72-541
    CLIPVisionModel / RobertaModel (huggingface_clip branch)   .../clip/modeling_clip.py:731-838, .../roberta/modeling_roberta.py:65-575
    OPEN_CLIP (causal text transformer, EOT pooling)           .../clip/modeling_openclip.py:255-383
    WukongModel (same shape, eps 1e-7, [SEP] pooling)          .../wukong/modeling_wukong.py:234-413
    Text2VideoRetrieval's frame pooling                        easynlp/appzoo/text2video_retrieval/model.py:82-105
    clip_grad_norm_ + AdamW.step        easynlp/core/trainer.py:315-337, easynlp/core/optimizers.py:405-464

`cfg["model_type"]` picks the kind: "chinese_clip" (default), "huggingface_clip", "open_clip", "wukong".  The towers share the code:
one pre-LN block stack (`_blocks_forward/_blocks_backward`) serves every ViT and the causal text towers through a parameter-name
table; `bert_forward/backward` serves BertModel and RobertaModel.

Numerics: bf16 GEMM/attention operands with fp32 accumulation, fp32 residual stream / LayerNorm statistics /
embeddings / logits / loss, fp32 master weights and gradients.  Layout: token-major [B*L, d] activations.
BERT-tower dropout (hidden / attention-probs) is fused into the LayerNorm and attention kernels as Philox masks keyed by a
device-resident forward-pass counter: a training forward draws fresh masks, its backward regenerates them ("dropout").
"""
import os
from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops
from .params import ParamStore

SM_TARGET = 148 * 2


def _splits_for(m, n, k):
    tiles = ((m + 127) // 128) * ((n + 255) // 256 if (n % 256 == 0 or n > 512) else (n + 127) // 128)
    s = max(1, SM_TARGET // max(1, tiles))
    return max(1, min(s, max(1, k // 256)))


def hf_engine_config(raw: dict, embed_dim: int) -> dict:
    """Flat engine config of the huggingface_clip branch from the reference's nested config.json ({'text_config': CLIPTextConfig kwargs,
    'vision_config': CLIPVisionConfig kwargs}, appzoo/clip/model.py:82-85; defaults of modelzoo/models/clip/configuration_clip.py:88-120,
    203-235).  embed_dim = rows of text_projection.weight in the checkpoint (model.py:93-96)."""
    t = dict(vocab_size=21128, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
             max_position_embeddings=512, hidden_act="gelu", layer_norm_eps=1e-12, pad_token_id=0, type_vocab_size=2,
             hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    t.update(raw.get("text_config", {}))
    v = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, image_size=224, patch_size=32,
             hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0)
    v.update(raw.get("vision_config", {}))
    if float(v.get("attention_dropout", 0.0) or 0.0) != 0.0:
        raise NotImplementedError("vision attention_dropout > 0 is not on the hot path (the reference default is 0)")
    return dict(model_type="huggingface_clip", embed_dim=int(embed_dim), image_resolution=v["image_size"], vision_layers=v["num_hidden_layers"],
                vision_width=v["hidden_size"], vision_patch_size=v["patch_size"], vision_heads=v["num_attention_heads"],
                vision_intermediate_size=v["intermediate_size"], vision_hidden_act=v["hidden_act"],
                vocab_size=t["vocab_size"], text_hidden_size=t["hidden_size"], text_intermediate_size=t["intermediate_size"],
                text_num_hidden_layers=t["num_hidden_layers"], text_num_attention_heads=t["num_attention_heads"],
                text_max_position_embeddings=t["max_position_embeddings"], text_type_vocab_size=t["type_vocab_size"],
                text_hidden_act=t["hidden_act"], text_layer_norm_eps=t["layer_norm_eps"], text_pad_token_id=t["pad_token_id"],
                text_hidden_dropout_prob=t["hidden_dropout_prob"], text_attention_probs_dropout_prob=t["attention_probs_dropout_prob"],
                text_initializer_range=t.get("initializer_range", 0.02))


def wukong_engine_config(raw: dict) -> dict:
    """flat engine config from a Wukong `config.json` ({"model": {"visual": {...}, "text": {...}}}: the constructor arguments of
    VisualTransformer / TextTransformer, modeling_wukong.py:269-277,312-319; appzoo/wukong_clip/model.py:51-55)"""
    v = raw["model"]["visual"]; t = raw["model"]["text"]
    if v["heads"] * 64 != v["width"] or t["heads"] * 64 != t["width"]:
        raise NotImplementedError("clipk attention needs head_dim 64 in both Wukong towers")
    if v["output_dim"] != t["output_dim"]:
        raise ValueError("Wukong towers must project to the same output_dim")
    return dict(model_type="wukong", embed_dim=v["output_dim"], image_resolution=v["input_resolution"], vision_layers=v["layers"],
                vision_width=v["width"], vision_patch_size=v["patch_size"], vocab_size=t["vocab_size"], context_length=t["context_length"],
                transformer_width=t["width"], transformer_heads=t["heads"], transformer_layers=t["layers"])


class ClipEngine:
    def __init__(self, cfg: dict, device="cuda", with_optimizer_state: bool = True):
        if isinstance(cfg.get("vision_layers"), (tuple, list)):
            raise NotImplementedError("ModifiedResNet visual towers are outside the hot path (ViT only)")
        self.cfg = cfg
        self.kind = cfg.get("model_type", "chinese_clip")
        self.hf = self.kind == "huggingface_clip"
        self.wk = self.kind == "wukong"
        self.oc = self.kind in ("open_clip", "wukong")          # pre-LN causal text transformer (OPEN_CLIP; Wukong's TextTransformer)
        if self.kind not in ("chinese_clip", "huggingface_clip", "open_clip", "wukong"):
            raise NotImplementedError(f"model_type {self.kind!r}")
        if self.oc:      # OPEN_CLIP ctor arguments (modeling_openclip.py:256-271) -> the text-tower keys this engine uses
            cfg = dict(cfg, text_hidden_size=cfg["transformer_width"], text_intermediate_size=4 * cfg["transformer_width"],
                       text_num_attention_heads=cfg["transformer_heads"], text_num_hidden_layers=cfg["transformer_layers"],
                       text_max_position_embeddings=cfg["context_length"])
            self.cfg = cfg
        self.dev = torch.device(device)
        self.W = cfg["vision_width"]; self.P = cfg["vision_patch_size"]; self.R = cfg["image_resolution"]
        self.E = cfg["embed_dim"]; self.g = self.R // self.P; self.Lv = self.g * self.g + 1
        self.Hv = self.W // 64                                   # modeling_chineseclip.py:289
        self.Iv = cfg.get("vision_intermediate_size", 4 * self.W)
        self.H = cfg["text_hidden_size"]; self.I = cfg["text_intermediate_size"]; self.Ht = cfg["text_num_attention_heads"]
        self.nv = cfg["vision_layers"]; self.nt = cfg["text_num_hidden_layers"]
        if self.H != self.Ht * 64 or self.W % 128 or self.H % 128 or self.E % 128:
            raise NotImplementedError("clipk kernels need head_dim 64 and widths that are multiples of 128")
        if self.hf and cfg["vision_heads"] * 64 != self.W:
            raise NotImplementedError("clipk attention needs head_dim 64 in the vision tower")
        self.kdim = 3 * self.P * self.P
        self.kdim_pad = (self.kdim + 7) // 8 * 8                 # ViT-L/14: 588 -> 592 (TMA rows are 16-byte multiples)
        if cfg.get("text_hidden_act", "gelu") != "gelu":
            raise NotImplementedError("text tower activation must be erf-GELU")
        self.vit_act = L.EPI_QUICK_GELU
        if self.hf:
            va = cfg.get("vision_hidden_act", "quick_gelu")
            if va not in ("quick_gelu", "gelu"):
                raise NotImplementedError(f"vision hidden_act {va!r}")
            self.vit_act = L.EPI_QUICK_GELU if va == "quick_gelu" else L.EPI_ERF_GELU
        self.text_eps = float(cfg.get("text_layer_norm_eps", 1e-12))   # modeling_chineseclip.py:311 / CLIPTextConfig.layer_norm_eps
        # eps of the pre-LN blocks: nn.LayerNorm default 1e-5 (modeling_chineseclip.py:170-176, modeling_openclip.py:108-121),
        # CLIPVisionConfig.layer_norm_eps (configuration_clip.py:211), 1e-7 in the Wukong towers (modeling_wukong.py:242-248,284-288,329)
        self.ln_eps = float(cfg.get("vision_layer_norm_eps", 1e-5)) if self.hf else (1e-7 if self.wk else 1e-5)
        self.pad_id = int(cfg.get("text_pad_token_id", 0))
        # parameter names of the two branches
        if self.hf:
            v = "vision_encoder.vision_model."
            self.vn = {"conv": v + "embeddings.patch_embedding.weight", "cls": v + "embeddings.class_embedding",
                       "pos": v + "embeddings.position_embedding.weight", "ln_pre": v + "pre_layrnorm", "ln_post": v + "post_layernorm",
                       "layer": v + "encoder.layers.{}.", "ln1": "layer_norm1", "ln2": "layer_norm2", "qkv_w": "self_attn.q_proj.weight",
                       "qkv_b": "self_attn.q_proj.bias", "out": "self_attn.out_proj", "fc1": "mlp.fc1", "fc2": "mlp.fc2"}
            self.tp = "text_encoder."
        else:
            v = "visual_encoder." if self.wk else "visual."
            self.vn = {"conv": v + "conv1.weight", "cls": v + "class_embedding", "pos": v + "positional_embedding",
                       "ln_pre": v + "ln_pre", "ln_post": v + "ln_post", "layer": v + "transformer.resblocks.{}.", "ln1": "ln_1",
                       "ln2": "ln_2", "qkv_w": "attn.in_proj_weight", "qkv_b": "attn.in_proj_bias", "out": "attn.out_proj",
                       "fc1": "mlp.c_fc", "fc2": "mlp.c_proj", "proj": v + "proj"}
            self.tp = "bert."
        # OPEN_CLIP's text tower: the same residual block as the ViT, under `transformer.resblocks.{i}.` (modeling_openclip.py:296-301);
        # Wukong's TextTransformer keeps it under `text_encoder.` with an `embedding_table` parameter (modeling_wukong.py:311-336)
        t = "text_encoder." if self.wk else ""
        self.on = {"layer": t + "transformer.resblocks.{}.", "ln1": "ln_1", "ln2": "ln_2", "qkv_w": "attn.in_proj_weight", "qkv_b": "attn.in_proj_bias",
                   "out": "attn.out_proj", "fc1": "mlp.c_fc", "fc2": "mlp.c_proj", "tok": t + ("embedding_table" if self.wk else "token_embedding.weight"),
                   "pos": t + "positional_embedding", "ln_final": t + "ln_final", "proj": t + "text_projection"}
        self.sep_id = int(cfg.get("sep_token_id", 102))          # Wukong pools the [SEP] position: `(x == 102).nonzero()` (modeling_wukong.py:349)
        self.params = ParamStore(cfg, device, with_optimizer_state)
        self._buf: Dict[tuple, torch.Tensor] = {}
        self._saved = None
        self._reducer = None      # OverlappedGradReducer while an eager multi-GPU backward is running
        self.norm_and_coef = torch.zeros(2, device=self.dev)
        self._norm_ws = torch.zeros(1024, dtype=torch.float64, device=self.dev)
        # device-resident optimizer step counter and {lr, step size}
        self._dev_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # device-resident count of TRAINING forward passes = the dropout stream offset: every forward(train) draws fresh masks (also
        # each micro-batch of a gradient-accumulation window and every replay of a captured graph); backward reuses the value
        self._dev_pass = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._dev_hyper = torch.zeros(2, dtype=torch.float32, device=self.dev)
        # BERT-tower dropout (nn.Dropout in modeling_bert.py:85,128,238,267,345); the ViT tower has none
        self.p_hidden = float(cfg.get("text_hidden_dropout_prob", 0.0) or 0.0)
        self.p_attn = float(cfg.get("text_attention_probs_dropout_prob", 0.0) or 0.0)
        self.dropout_seed = 0x5EED_C11B
        self._drops = {}
        self._conv_pad = None; self._conv_pad_version = -1
        self._peer = None; self._peer_key = None; self._peer_on = False      # peer-memory collectives of the contrastive head (distributed.PeerGroup)

    # ------------------------------------------------------------------ buffers
    def buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._buf[key] = t
        return t

    def zbuf(self, name, shape, dtype):
        """like buf, but zero-filled when first created (for operands whose padding rows must stay zero)"""
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            self._buf[key] = t
        return t

    def bf(self, name, *shape):
        return self.buf(name, shape, torch.bfloat16)

    def f32(self, name, *shape):
        return self.buf(name, shape, torch.float32)

    # ------------------------------------------------------------------ ViT
    # ------------------------------------------------------------------ pre-LN residual blocks (ViT towers, OPEN_CLIP's text tower)
    def _blocks_forward(self, x, nm, n_layers, B, Ltok, W, I, Hh, save, tg, act, causal):
        """ResidualAttentionBlock stack (modeling_chineseclip.py:184-216 / modeling_openclip.py:123-160 / CLIPEncoderLayer
        modeling_clip.py:283-334): x + attn(ln_1(x)), then + mlp(ln_2(.)).  Every projection GEMM writes its branch output as bf16 (plain
        epilogue); the fp32 residual add is fused into the LayerNorm that follows it (x_new = x + branch is stored by that kernel for
        backward / the next residual).  Returns (saved layer dicts, pending residual, last stored x, branch buffer): the caller's final
        LayerNorm performs the last pending add."""
        P_ = self.params; M = B * Ltok
        ybr = self.bf(tg + "ybr", M, W)              # branch output (attention out-proj / MLP c_proj), reused
        pending = None                               # residual stream owed to the next LayerNorm
        layers = []
        for i in range(n_layers):
            p = nm["layer"].format(i)
            tag = f"{tg}{i}." if save else tg + "t."
            ly = {}
            ly["h"] = self.bf(tag + "h", M, W); ly["m1"] = self.f32(tag + "m1", M); ly["r1"] = self.f32(tag + "r1", M)
            g1, b1 = P_.p(p + nm["ln1"] + ".weight"), P_.p(p + nm["ln1"] + ".bias")
            if pending is None:
                ops.layernorm_fwd(x, g1, b1, self.ln_eps, ly["h"], None, ly["m1"], ly["r1"])
            else:       # x = x1_prev + c_proj(...) of the previous block
                x_new = self.f32(f"{tg}x.{i}" if save else f"{tg}x.t{i % 2}", M, W)
                ops.layernorm_fwd(pending, g1, b1, self.ln_eps, ly["h"], None, ly["m1"], ly["r1"], add=ybr, x_out=x_new)
                x = x_new
            ly["x_in"] = x
            ly["qkv"] = self.bf(tag + "qkv", M, 3 * W)
            ops.gemm(ly["h"], P_.w(p + nm["qkv_w"], (3 * W, W)), ly["qkv"], bias=P_.p(p + nm["qkv_b"], (3 * W,)))
            ly["ctx"] = self.bf(tag + "ctx", M, W); ly["lse"] = self.f32(tag + "lse", B * Hh * Ltok)
            ops.attention_fwd(ly["qkv"], None, ly["ctx"], ly["lse"], B, Ltok, Hh, causal=causal)
            ops.gemm(ly["ctx"], P_.w(p + nm["out"] + ".weight"), ybr, bias=P_.p(p + nm["out"] + ".bias"))
            ly["x1"] = self.f32(tag + "x1", M, W)
            ly["h2"] = self.bf(tag + "h2", M, W); ly["m2"] = self.f32(tag + "m2", M); ly["r2"] = self.f32(tag + "r2", M)
            ops.layernorm_fwd(x, P_.p(p + nm["ln2"] + ".weight"), P_.p(p + nm["ln2"] + ".bias"), self.ln_eps, ly["h2"], None, ly["m2"], ly["r2"],
                              add=ybr, x_out=ly["x1"])
            # "z" holds act'(z) (the activation's derivative) saved for backward, "a" the activation
            ly["z"] = self.bf(tag + "z", M, I); ly["a"] = self.bf(tag + "a", M, I)
            ops.gemm(ly["h2"], P_.w(p + nm["fc1"] + ".weight"), ly["z"], bias=P_.p(p + nm["fc1"] + ".bias"), mode=act, out2=ly["a"])
            ops.gemm(ly["a"], P_.w(p + nm["fc2"] + ".weight"), ybr, bias=P_.p(p + nm["fc2"] + ".bias"))
            pending = ly["x1"]
            layers.append(ly)
        return layers, pending, x, ybr

    def _blocks_backward(self, layers, dX, dXb, nm, n_layers, B, Ltok, W, I, Hh, tg, causal):
        """backward of _blocks_forward.  In: dX (fp32) / dXb (bf16) = gradient of the stack's output (its bias-gradient of the last c_proj
        already taken by the caller's LayerNorm backward).  Out: dX = gradient of the stack's input."""
        P_ = self.params; M = B * Ltok
        dz = self.bf(tg + "dz", M, I); dh = self.bf(tg + "dh", M, W); dctx = self.bf(tg + "dctx", M, W); dqkv = self.bf(tg + "dqkv", M, 3 * W)
        for i in reversed(range(n_layers)):
            p = nm["layer"].format(i)
            ly = layers[i]
            # MLP
            ops.gemm(dXb, ly["a"], P_.g(p + nm["fc2"] + ".weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(W, I, M))
            ops.gemm(dXb, P_.w(p + nm["fc2"] + ".weight"), dz, b_mn_major=1, mode=L.EPI_MUL_AUX, aux=ly["z"],
                     colsum=P_.g(p + nm["fc1"] + ".bias"))            # dz = (dX W) o act'(z); its column sums = d(c_fc.bias)
            ops.gemm(dz, ly["h2"], P_.g(p + nm["fc1"] + ".weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(I, W, M))
            ops.gemm(dz, P_.w(p + nm["fc1"] + ".weight"), dh, b_mn_major=1)
            dX1 = self.f32(tg + "dX.b", M, W); dX1b = self.bf(tg + "dXb.b", M, W)
            ops.layernorm_bwd(dh, ly["x1"], P_.p(p + nm["ln2"] + ".weight"), ly["m2"], ly["r2"], dx_add=dX, dx_f32=dX1, dx_bf16=dX1b,
                              dgamma=P_.g(p + nm["ln2"] + ".weight"), dbeta=P_.g(p + nm["ln2"] + ".bias"), dbias=P_.g(p + nm["out"] + ".bias"))
            # attention
            ops.gemm(dX1b, ly["ctx"], P_.g(p + nm["out"] + ".weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(W, W, M))
            ops.gemm(dX1b, P_.w(p + nm["out"] + ".weight"), dctx, b_mn_major=1)
            ops.attention_bwd(ly["qkv"], None, ly["ctx"], ly["lse"], dctx, dqkv, B, Ltok, Hh, dqkv_colsum=P_.g(p + nm["qkv_b"], (3 * W,)), causal=causal)
            ops.gemm(dqkv, ly["h"], P_.g(p + nm["qkv_w"], (3 * W, W)), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(3 * W, W, M))
            ops.gemm(dqkv, P_.w(p + nm["qkv_w"], (3 * W, W)), dh, b_mn_major=1)
            bias_prev = P_.g(nm["layer"].format(i - 1) + nm["fc2"] + ".bias") if i > 0 else None
            ops.layernorm_bwd(dh, ly["x_in"], P_.p(p + nm["ln1"] + ".weight"), ly["m1"], ly["r1"], dx_add=dX1, dx_f32=dX, dx_bf16=dXb,
                              dgamma=P_.g(p + nm["ln1"] + ".weight"), dbeta=P_.g(p + nm["ln1"] + ".bias"), dbias=bias_prev)
            self._grads_ready(p)
        return dX

    def _conv_operand(self):
        """patch-embedding weight as the [W, kdim_pad] bf16 GEMM operand (zero-padded copy when 3*P*P is not a multiple of 8)"""
        P_ = self.params
        if self.kdim_pad == self.kdim:
            return P_.w(self.vn["conv"], (self.W, self.kdim))
        if self._conv_pad is None or self._conv_pad_version != P_.version:
            if self._conv_pad is None:
                self._conv_pad = torch.zeros((self.W, self.kdim_pad), dtype=torch.bfloat16, device=self.dev)
            self._conv_pad[:, :self.kdim].copy_(P_.w(self.vn["conv"], (self.W, self.kdim)))
            self._conv_pad_version = P_.version
        return self._conv_pad

    def vit_forward(self, pixels: torch.Tensor, save: bool):
        """VisualTransformer.forward (modeling_chineseclip.py:219-253) / CLIPVisionTransformer.forward (modeling_clip.py:731-776): the two
        towers are the same pre-LN ViT; the HF one stores q/k/v separately (adjacent here -> one packed projection), takes its MLP width
        and activation from the config and ends in a biased `vision_projection` Linear instead of the `proj` matrix."""
        P_ = self.params; W = self.W; B = pixels.shape[0]; Lv = self.Lv; M = B * Lv; Hh = self.Hv; vn = self.vn; Iv = self.Iv
        assert pixels.dtype == torch.float32 and pixels.is_contiguous() and pixels.shape[1:] == (3, self.R, self.R)
        npatch = B * self.g * self.g
        patches = self.zbuf("v.patches", (npatch, self.kdim_pad), torch.bfloat16)
        ops.im2col_patches(pixels, patches, B, self.R, self.P)
        patch_out = self.f32("v.patch_out", npatch, W)
        ops.gemm(patches, self._conv_operand(), patch_out)
        x0 = self.f32("v.x0", M, W)
        ops.vit_assemble(patch_out, P_.p(vn["cls"]), P_.p(vn["pos"]), x0, B, Lv, W)
        x = self.f32("v.x.0", M, W)
        st = {"B": B, "mean0": self.f32("v.mean0", M), "rstd0": self.f32("v.rstd0", M), "layers": []}
        ops.layernorm_fwd(x0, P_.p(vn["ln_pre"] + ".weight"), P_.p(vn["ln_pre"] + ".bias"), self.ln_eps, None, x, st["mean0"], st["rstd0"])
        layers, pending, x, ybr = self._blocks_forward(x, vn, self.nv, B, Lv, W, Iv, Hh, save, "v.", self.vit_act, causal=False)
        st["layers"] = layers
        # ln_post on the CLS rows of x_final = x1_last + c_proj(...): the add is fused here too; x_cls [B, W] is kept for backward
        st["x_cls"] = self.f32("v.x_cls", B, W)
        st["pooled"] = self.bf("v.pooled", B, W); st["mp"] = self.f32("v.mp", B); st["rp"] = self.f32("v.rp", B)
        gp, bp = P_.p(vn["ln_post"] + ".weight"), P_.p(vn["ln_post"] + ".bias")
        if pending is None:      # zero-layer tower (degenerate configs): no pending residual
            ops.layernorm_fwd(x, gp, bp, self.ln_eps, st["pooled"], None, st["mp"], st["rp"], rows=B, ldx=Lv * W)
            st["x_cls"] = None; st["x_final"] = x
        else:
            ops.layernorm_fwd(pending, gp, bp, self.ln_eps, st["pooled"], None, st["mp"], st["rp"],
                              rows=B, ldx=Lv * W, add=ybr, ldadd=Lv * W, x_out=st["x_cls"])
        st["feat"] = self.f32("v.feat", B, self.E)
        if self.hf:   # image_embeds = vision_projection(pooled.detach())  (appzoo/clip/model.py:142-143)
            ops.gemm(st["pooled"], P_.w("vision_projection.weight"), st["feat"], bias=P_.p("vision_projection.bias"))
        else:
            ops.gemm(st["pooled"], P_.w(vn["proj"]), st["feat"], b_mn_major=1)
        st["embeds"] = self.f32("v.embeds", B, self.E); st["norm"] = self.f32("v.norm", B)
        if self._peer_on:      # fused normalise + all-gather: the embeddings land in every rank's image gallery (csrc/peer.cu)
            self._peer.l2norm_allgather(st["feat"], st["embeds"], st["norm"], "image")
        else:
            ops.l2norm_fwd(st["feat"], st["embeds"], st["norm"], B, self.E)
        return st

    def vit_head_backward_hf(self, st, d_embeds: torch.Tensor):
        """huggingface_clip: the image tower is frozen by `.detach()` (model.py:142) -- only vision_projection receives a gradient"""
        P_ = self.params; B = st["B"]; E = self.E
        dfeat = self.f32("v.dfeat", B, E); dfeat_b = self.bf("v.dfeat_b", B, E)
        ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], dfeat, dfeat_b, B, E)
        ops.gemm(dfeat_b, st["pooled"], P_.g("vision_projection.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)   # dW[E,W] += dfeat^T pooled
        ops.colsum(dfeat, P_.g("vision_projection.bias"), B, E)
        self._grads_ready("vision_projection.")

    def vit_backward(self, st, d_embeds: torch.Tensor):
        P_ = self.params; W = self.W; B = st["B"]; Lv = self.Lv; M = B * Lv; Hh = self.Hv; E = self.E; vn = self.vn; Iv = self.Iv
        if self.hf:
            return self.vit_head_backward_hf(st, d_embeds)
        dfeat_b = self.bf("v.dfeat_b", B, E)
        ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], None, dfeat_b, B, E)
        # d proj[W,E] += pooled^T dfeat ; dpooled = dfeat proj^T
        ops.gemm(st["pooled"], dfeat_b, P_.g(vn["proj"]), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
        dpooled = self.f32("v.dpooled", B, W)
        ops.gemm(dfeat_b, P_.w(vn["proj"]), dpooled)
        dX = self.f32("v.dX.a", M, W); dXb = self.bf("v.dXb.a", M, W)
        dX.zero_()
        last = st["layers"][-1] if self.nv else None
        bias_prev = P_.g(vn["layer"].format(self.nv - 1) + vn["fc2"] + ".bias") if self.nv else None
        x_post, ld_post = (st["x_cls"], W) if st.get("x_cls") is not None else (st["x_final"], Lv * W)
        ops.layernorm_bwd(dpooled, x_post, P_.p(vn["ln_post"] + ".weight"), st["mp"], st["rp"], dx_f32=dX,
                          dgamma=P_.g(vn["ln_post"] + ".weight"), dbeta=P_.g(vn["ln_post"] + ".bias"), dbias=bias_prev,
                          rows=B, ldx=ld_post, lddx=Lv * W)
        ops.cast_bf16(dX, dXb)
        dX = self._blocks_backward(st["layers"], dX, dXb, vn, self.nv, B, Lv, W, Iv, Hh, "v.", causal=False)
        # ln_pre, token assembly, patch embedding
        dx0 = self.f32("v.dx0", M, W)
        ops.layernorm_bwd(dX, self.f32("v.x0", M, W), P_.p(vn["ln_pre"] + ".weight"), st["mean0"], st["rstd0"], dx_f32=dx0,
                          dgamma=P_.g(vn["ln_pre"] + ".weight"), dbeta=P_.g(vn["ln_pre"] + ".bias"))
        ops.colsum(dx0, P_.g(vn["pos"]).view(-1), B, Lv * W)
        ops.colsum(dx0, P_.g(vn["cls"]), B, W, ldx=Lv * W)
        npatch = B * self.g * self.g; kdim = 3 * self.P * self.P
        dpatch = self.bf("v.dpatch", npatch, W)
        ops.vit_assemble_bwd(dx0, dpatch, B, Lv, W)
        patches = self.zbuf("v.patches", (npatch, self.kdim_pad), torch.bfloat16)
        if self.kdim_pad == kdim:
            ops.gemm(dpatch, patches, P_.g(vn["conv"], (W, kdim)), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD, splits=_splits_for(W, kdim, npatch))
        else:       # ViT-*/14: the operand rows are padded 588 -> 592; the weight gradient goes through a padded scratch tile
            dpad = self.f32("v.dconv_pad", W, self.kdim_pad)
            dpad.zero_()
            ops.gemm(dpatch, patches, dpad, a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD, splits=_splits_for(W, self.kdim_pad, npatch))
            P_.g(vn["conv"], (W, kdim)).add_(dpad[:, :kdim])

    # ------------------------------------------------------------------ BERT
    def _drop(self, train: bool, p: float, site: int):
        """clipk_dropout_t of one call site (None in eval mode / p == 0); the structs are cached so their addresses stay valid"""
        if not train or p <= 0.0:
            return None
        key = (p, site)
        d = self._drops.get(key)
        if d is None:
            d = ops.make_dropout(p, self.dropout_seed, site, self._dev_pass)
            self._drops[key] = d
        return d

    # ------------------------------------------------------------------ OPEN_CLIP text tower
    def octext_forward(self, ids: torch.Tensor, save: bool):
        """OPEN_CLIP.encode_text (modeling_openclip.py:358-371): token + positional embedding -> causal pre-LN transformer -> ln_final ->
        features of the EOT token (the highest id of each sequence) @ text_projection"""
        P_ = self.params; W = self.H; B, Lt = ids.shape; M = B * Lt; Hh = self.Ht
        if Lt != self.cfg["context_length"]:
            raise ValueError(f"open_clip texts must be padded to context_length={self.cfg['context_length']} (got {Lt})")
        st = {"B": B, "Lt": Lt, "ids": ids}
        st["pos_ids"] = self.buf("o.pos_ids", (B, Lt), torch.int32)
        st["pos_ids"].copy_(torch.arange(Lt, device=self.dev, dtype=torch.int32).expand(B, Lt))
        ztype = self.zbuf("o.ztype", (1, W), torch.float32)                       # no token-type table on this tower
        x = self.f32("o.x.0", M, W)
        ops.embed_gather(ids, st["pos_ids"], None, None, P_.p(self.on["tok"]), P_.p(self.on["pos"]), ztype, x, None, -1)
        layers, pending, x, ybr = self._blocks_forward(x, self.on, self.nt, B, Lt, W, 4 * W, Hh, save, "o.", L.EPI_QUICK_GELU, causal=True)
        st["layers"] = layers
        # ln_final on every row of x_final = x1_last + c_proj(...), then the EOT rows are gathered for the projection
        st["x_fin"] = self.f32("o.x_fin", M, W); st["hf"] = self.bf("o.hfin", M, W); st["mf"] = self.f32("o.mf", M); st["rf"] = self.f32("o.rf", M)
        g, b = P_.p(self.on["ln_final"] + ".weight"), P_.p(self.on["ln_final"] + ".bias")
        if pending is None:
            ops.layernorm_fwd(x, g, b, self.ln_eps, st["hf"], None, st["mf"], st["rf"]); st["x_fin"] = x
        else:
            ops.layernorm_fwd(pending, g, b, self.ln_eps, st["hf"], None, st["mf"], st["rf"], add=ybr, x_out=st["x_fin"])
        st["eot"] = self.buf("o.eot", (B,), torch.int32)
        if self.wk:
            ops.find_token_rows(ids, self.sep_id, st["eot"])       # exactly one [SEP] per text (WukongCLIPDataset.tokenize, data.py:180-187)
        else:
            ops.argmax_rows(ids, st["eot"])
        st["pooled"] = self.bf("o.pooled", B, W)
        ops.gather_rows_bf16(st["hf"], st["eot"], st["pooled"], B, Lt, W)
        st["feat"] = self.f32("t.feat", B, self.E)
        ops.gemm(st["pooled"], P_.w(self.on["proj"]), st["feat"], b_mn_major=1)
        st["embeds"] = self.f32("t.embeds", B, self.E); st["norm"] = self.f32("t.norm", B)
        if self._peer_on:
            self._peer.l2norm_allgather(st["feat"], st["embeds"], st["norm"], "text")
        else:
            ops.l2norm_fwd(st["feat"], st["embeds"], st["norm"], B, self.E)
        return st

    def octext_backward(self, st, d_embeds: torch.Tensor):
        P_ = self.params; W = self.H; B = st["B"]; Lt = st["Lt"]; M = B * Lt; Hh = self.Ht; E = self.E
        dfeat_b = self.bf("t.dfeat_b", B, E)
        ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], None, dfeat_b, B, E)
        ops.gemm(st["pooled"], dfeat_b, P_.g(self.on["proj"]), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
        dpooled = self.f32("o.dpooled", B, W)
        ops.gemm(dfeat_b, P_.w(self.on["proj"]), dpooled)
        dhf = self.f32("o.dhf", M, W)
        dhf.zero_()
        ops.scatter_rows_f32(dpooled, st["eot"], dhf, B, Lt, W)                    # only the EOT rows carry a gradient into ln_final
        dX = self.f32("o.dX.a", M, W); dXb = self.bf("o.dXb.a", M, W)
        bias_prev = P_.g(self.on["layer"].format(self.nt - 1) + self.on["fc2"] + ".bias") if self.nt else None
        ops.layernorm_bwd(dhf, st["x_fin"], P_.p(self.on["ln_final"] + ".weight"), st["mf"], st["rf"], dx_f32=dX, dx_bf16=dXb,
                          dgamma=P_.g(self.on["ln_final"] + ".weight"), dbeta=P_.g(self.on["ln_final"] + ".bias"), dbias=bias_prev)
        dX = self._blocks_backward(st["layers"], dX, dXb, self.on, self.nt, B, Lt, W, 4 * W, Hh, "o.", causal=True)
        ztype_g = self.zbuf("o.ztype_g", (1, W), torch.float32)
        ops.embed_gather_bwd(st["ids"], st["pos_ids"], None, dX, P_.g(self.on["tok"]), P_.g(self.on["pos"]), ztype_g, -1)
        self._grads_ready(self.on["tok"]); self._grads_ready(self.on["pos"]); self._grads_ready(self.on["ln_final"] + ".")

    def bert_forward(self, ids: torch.Tensor, save: bool, train: bool = False, token_type_ids=None, attention_mask=None):
        """BertModel (chinese_clip: mask = ids != 0, modeling_chineseclip.py:347-349) or RobertaModel (huggingface_clip: pad-aware
        position ids, the batch's token_type_ids / attention_mask, tanh pooler; appzoo/clip/model.py:128-137)"""
        if self.oc:
            return self.octext_forward(ids, save)
        P_ = self.params; H = self.H; I = self.I; B, Lt = ids.shape; M = B * Lt; Hh = self.Ht; tp = self.tp
        assert ids.dtype == torch.int64 and ids.is_contiguous()
        if Lt > self.cfg["text_max_position_embeddings"] - (self.pad_id + 1 if self.hf else 0):
            raise ValueError("sequence longer than the position table")
        eps = self.text_eps
        st = {"B": B, "Lt": Lt, "ids": ids, "layers": [], "train": train}
        st["e"] = self.f32("t.e", M, H); st["mask"] = self.f32("t.mask", M)
        if self.hf:
            st["pos_ids"] = self.buf("t.pos_ids", (B, Lt), torch.int32)
            st["type_ids"] = token_type_ids.to(self.dev).long().contiguous() if token_type_ids is not None else None
            am = attention_mask.to(self.dev).long().contiguous() if attention_mask is not None else None
            ops.position_ids(ids, st["pos_ids"], self.pad_id)
            ops.embed_gather(ids, st["pos_ids"], st["type_ids"], am, P_.p(tp + "embeddings.word_embeddings.weight"),
                             P_.p(tp + "embeddings.position_embeddings.weight"), P_.p(tp + "embeddings.token_type_embeddings.weight"),
                             st["e"], st["mask"], self.pad_id)
        else:
            ops.bert_embed(ids.view(-1), P_.p("bert.embeddings.word_embeddings.weight"), P_.p("bert.embeddings.position_embeddings.weight"),
                           P_.p("bert.embeddings.token_type_embeddings.weight"), st["e"], M, Lt, H, self.cfg["vocab_size"], key_mask=st["mask"])
        x = self.f32("t.x.0", M, H); xb = self.bf("t.xb.0", M, H)
        st["me"] = self.f32("t.me", M); st["re"] = self.f32("t.re", M)
        ops.layernorm_fwd(st["e"], P_.p(tp + "embeddings.LayerNorm.weight"), P_.p(tp + "embeddings.LayerNorm.bias"), eps, xb, x,
                          st["me"], st["re"], drop=self._drop(train, self.p_hidden, 1), drop_mode=2)
        tbr = self.bf("t.tbr", M, H)                 # bf16 branch output of attention.output.dense / output.dense
        for i in range(self.nt):
            p = f"{tp}encoder.layer.{i}."
            tag = f"t.{i}." if save else "t.t."
            ly = {"x_in": x, "xb_in": xb}
            ly["qkv"] = self.bf(tag + "qkv", M, 3 * H)
            ops.gemm(xb, P_.w(p + "attention.self.query.weight", (3 * H, H)), ly["qkv"], bias=P_.p(p + "attention.self.query.bias", (3 * H,)))
            ly["ctx"] = self.bf(tag + "ctx", M, H); ly["lse"] = self.f32(tag + "lse", B * Hh * Lt)
            ops.attention_fwd(ly["qkv"], st["mask"], ly["ctx"], ly["lse"], B, Lt, Hh, drop=self._drop(train, self.p_attn, 16 * (i + 1)))
            ops.gemm(ly["ctx"], P_.w(p + "attention.output.dense.weight"), tbr, bias=P_.p(p + "attention.output.dense.bias"))
            ly["s1"] = self.f32(tag + "s1", M, H)
            ly["y1"] = self.f32(tag + "y1", M, H); ly["y1b"] = self.bf(tag + "y1b", M, H)
            ly["m1"] = self.f32(tag + "m1", M); ly["r1"] = self.f32(tag + "r1", M)
            ops.layernorm_fwd(x, P_.p(p + "attention.output.LayerNorm.weight"), P_.p(p + "attention.output.LayerNorm.bias"), eps,
                              ly["y1b"], ly["y1"], ly["m1"], ly["r1"], add=tbr, x_out=ly["s1"],      # LN(dropout(dense(ctx)) + input)
                              drop=self._drop(train, self.p_hidden, 16 * (i + 1) + 1), drop_mode=1)
            ly["z"] = self.bf(tag + "z", M, I); ly["a"] = self.bf(tag + "a", M, I)                  # z = gelu'(.) saved for backward
            ops.gemm(ly["y1b"], P_.w(p + "intermediate.dense.weight"), ly["z"], bias=P_.p(p + "intermediate.dense.bias"),
                     mode=L.EPI_ERF_GELU, out2=ly["a"])
            ops.gemm(ly["a"], P_.w(p + "output.dense.weight"), tbr, bias=P_.p(p + "output.dense.bias"))
            ly["s2"] = self.f32(tag + "s2", M, H)
            x = self.f32(f"t.x.{i + 1}" if save else f"t.x.t{i % 2}", M, H)
            xb = self.bf(f"t.xb.{i + 1}" if save else f"t.xb.t{i % 2}", M, H)
            ly["m2"] = self.f32(tag + "m2", M); ly["r2"] = self.f32(tag + "r2", M)
            ops.layernorm_fwd(ly["y1"], P_.p(p + "output.LayerNorm.weight"), P_.p(p + "output.LayerNorm.bias"), eps, xb, x, ly["m2"], ly["r2"],
                              add=tbr, x_out=ly["s2"], drop=self._drop(train, self.p_hidden, 16 * (i + 1) + 2), drop_mode=1)
            st["layers"].append(ly)
        st["xb_final"] = xb
        st["feat"] = self.f32("t.feat", B, self.E)
        cls_rows = xb.view(B, Lt * H)[:, :H]                           # x[:, 0, :]  (modeling_chineseclip.py:350)
        if self.hf:
            # text_embeds = text_projection(pooler_output), pooler_output = tanh(dense(h[:, 0]))  (model.py:135-136, modeling_roberta.py:559-575)
            st["pool_pre"] = self.f32("t.pool_pre", B, H); st["pool"] = self.f32("t.pool", B, H); st["pool_b"] = self.bf("t.pool_b", B, H)
            ops.gemm(cls_rows, P_.w(tp + "pooler.dense.weight"), st["pool_pre"], bias=P_.p(tp + "pooler.dense.bias"))
            ops.tanh_fwd(st["pool_pre"], st["pool"], st["pool_b"])
            ops.gemm(st["pool_b"], P_.w("text_projection.weight"), st["feat"], bias=P_.p("text_projection.bias"))
        else:
            ops.gemm(cls_rows, P_.w("text_projection"), st["feat"], b_mn_major=1)
        st["embeds"] = self.f32("t.embeds", B, self.E); st["norm"] = self.f32("t.norm", B)
        if self._peer_on:
            self._peer.l2norm_allgather(st["feat"], st["embeds"], st["norm"], "text")
        else:
            ops.l2norm_fwd(st["feat"], st["embeds"], st["norm"], B, self.E)
        return st

    def bert_backward(self, st, d_embeds: torch.Tensor):
        if self.oc:
            return self.octext_backward(st, d_embeds)
        P_ = self.params; H = self.H; I = self.I; B = st["B"]; Lt = st["Lt"]; M = B * Lt; Hh = self.Ht; E = self.E; tp = self.tp
        train = st.get("train", False)
        dfeat_b = self.bf("t.dfeat_b", B, E)
        cls_rows = st["xb_final"].view(B, Lt * H)[:, :H]
        dOut = self.f32("t.dOut", M, H)
        dOut.zero_()
        if self.hf:
            dfeat = self.f32("t.dfeat", B, E)
            ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], dfeat, dfeat_b, B, E)
            # text_projection (Linear [E,H] + bias) <- tanh pooler (dense [H,H] + bias) <- CLS hidden state
            ops.gemm(dfeat_b, st["pool_b"], P_.g("text_projection.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
            ops.colsum(dfeat, P_.g("text_projection.bias"), B, E)
            dpool = self.f32("t.dpool", B, H); dpre = self.f32("t.dpre", B, H); dpre_b = self.bf("t.dpre_b", B, H)
            ops.gemm(dfeat_b, P_.w("text_projection.weight"), dpool, b_mn_major=1)
            ops.tanh_bwd(dpool, st["pool"], dpre, dpre_b)
            ops.gemm(dpre_b, cls_rows, P_.g(tp + "pooler.dense.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
            ops.colsum(dpre, P_.g(tp + "pooler.dense.bias"), B, H)
            ops.gemm(dpre_b, P_.w(tp + "pooler.dense.weight"), dOut.view(B, Lt * H)[:, :H], b_mn_major=1)
            self._grads_ready("text_projection."); self._grads_ready(tp + "pooler.")
        else:
            ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], None, dfeat_b, B, E)
            ops.gemm(cls_rows, dfeat_b, P_.g("text_projection"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
            ops.gemm(dfeat_b, P_.w("text_projection"), dOut.view(B, Lt * H)[:, :H])
        dy, dy_add = dOut, None
        ds2 = self.f32("t.ds2", M, H); ds2b = self.bf("t.ds2b", M, H); ds1 = self.f32("t.ds1", M, H); ds1b = self.bf("t.ds1b", M, H)
        dz = self.bf("t.dz", M, I); dpart = self.bf("t.dpart", M, H); dctx = self.bf("t.dctx", M, H); dqkv = self.bf("t.dqkv", M, 3 * H)
        dxp = self.bf("t.dxp", M, H)
        for i in reversed(range(self.nt)):
            p = f"{tp}encoder.layer.{i}."
            ly = st["layers"][i]
            ops.layernorm_bwd(dy, ly["s2"], P_.p(p + "output.LayerNorm.weight"), ly["m2"], ly["r2"], dy_add=dy_add, dx_f32=ds2, dx_bf16=ds2b,
                              dgamma=P_.g(p + "output.LayerNorm.weight"), dbeta=P_.g(p + "output.LayerNorm.bias"), dbias=P_.g(p + "output.dense.bias"),
                              drop=self._drop(train, self.p_hidden, 16 * (i + 1) + 2), drop_mode=1)
            ops.gemm(ds2b, ly["a"], P_.g(p + "output.dense.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD, splits=_splits_for(H, I, M))
            ops.gemm(ds2b, P_.w(p + "output.dense.weight"), dz, b_mn_major=1, mode=L.EPI_MUL_AUX, aux=ly["z"],
                     colsum=P_.g(p + "intermediate.dense.bias"))
            ops.gemm(dz, ly["y1b"], P_.g(p + "intermediate.dense.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD, splits=_splits_for(I, H, M))
            ops.gemm(dz, P_.w(p + "intermediate.dense.weight"), dpart, b_mn_major=1)
            ops.layernorm_bwd(dpart, ly["s1"], P_.p(p + "attention.output.LayerNorm.weight"), ly["m1"], ly["r1"], dy_add=ds2, dx_f32=ds1,
                              dx_bf16=ds1b, dgamma=P_.g(p + "attention.output.LayerNorm.weight"),
                              dbeta=P_.g(p + "attention.output.LayerNorm.bias"), dbias=P_.g(p + "attention.output.dense.bias"),
                              drop=self._drop(train, self.p_hidden, 16 * (i + 1) + 1), drop_mode=1)
            ops.gemm(ds1b, ly["ctx"], P_.g(p + "attention.output.dense.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(H, H, M))
            ops.gemm(ds1b, P_.w(p + "attention.output.dense.weight"), dctx, b_mn_major=1)
            ops.attention_bwd(ly["qkv"], st["mask"], ly["ctx"], ly["lse"], dctx, dqkv, B, Lt, Hh, drop=self._drop(train, self.p_attn, 16 * (i + 1)),
                              dqkv_colsum=P_.g(p + "attention.self.query.bias", (3 * H,)))   # query|key|value biases are adjacent (params.py)
            ops.gemm(dqkv, ly["xb_in"], P_.g(p + "attention.self.query.weight", (3 * H, H)), a_mn_major=1, b_mn_major=1,
                     mode=L.EPI_ATOMIC_ADD, splits=_splits_for(3 * H, H, M))
            ops.gemm(dqkv, P_.w(p + "attention.self.query.weight", (3 * H, H)), dxp, b_mn_major=1)
            # d(layer input) = dxp (bf16) + ds1 (f32): consumed by the LayerNorm backward of the layer below, which writes ds2
            # (a different buffer) before ds1 is overwritten again
            dy, dy_add = dxp, ds1
            self._grads_ready(p)
        de = self.f32("t.de", M, H)
        ops.layernorm_bwd(dy, st["e"], P_.p(tp + "embeddings.LayerNorm.weight"), st["me"], st["re"], dy_add=dy_add, dx_f32=de,
                          dgamma=P_.g(tp + "embeddings.LayerNorm.weight"), dbeta=P_.g(tp + "embeddings.LayerNorm.bias"),
                          drop=self._drop(train, self.p_hidden, 1), drop_mode=2)
        if self.hf:
            ops.embed_gather_bwd(st["ids"], st["pos_ids"], st["type_ids"], de, P_.g(tp + "embeddings.word_embeddings.weight"),
                                 P_.g(tp + "embeddings.position_embeddings.weight"), P_.g(tp + "embeddings.token_type_embeddings.weight"), self.pad_id)
        else:
            ops.bert_embed_bwd(st["ids"].view(-1), de, P_.g("bert.embeddings.word_embeddings.weight"), M, H, self.cfg["vocab_size"])
            ops.colsum(de, P_.g("bert.embeddings.position_embeddings.weight").view(-1)[:Lt * H], B, Lt * H)
            ops.colsum(de, P_.g("bert.embeddings.token_type_embeddings.weight")[0], M, H)
        self._grads_ready(tp + "embeddings.")

    # ------------------------------------------------------------------ contrastive head
    def loss_forward(self, text_embeds, image_embeds, gallery_image=None, gallery_text=None, label_offset=0, want_logits=True):
        """Two CE strips: local texts vs the image gallery, local images vs the text gallery (appzoo/clip/model.py:148-164).  With no
        gallery given the local batch is the gallery (world size 1: identical to the reference's [B,B] loss).
        The dots come from the tcgen05 GEMM on bf16 hi/lo splits of the fp32 embeddings (K = 3E, fp32-level logits; loss.cu)."""
        gi = image_embeds if gallery_image is None else gallery_image
        gt = text_embeds if gallery_text is None else gallery_text
        B = text_embeds.shape[0]; G = gi.shape[0]; E = self.E
        Gp = (G + 7) // 8 * 8                                   # GEMM N / leading dimensions are multiples of 8; padding rows stay zero
        ls = self.params.p("logit_scale").view(-1)
        Ts = self.bf("l.Ts", B, 3 * E); Is = self.bf("l.Is", B, 3 * E)
        GIs = self.zbuf("l.GIs", (Gp, 3 * E), torch.bfloat16); GTs = self.zbuf("l.GTs", (Gp, 3 * E), torch.bfloat16)
        ops.split_bf16x3(text_embeds, Ts, 0); ops.split_bf16x3(image_embeds, Is, 0)
        ops.split_bf16x3(gi, GIs, 1); ops.split_bf16x3(gt, GTs, 1)
        S_t = self.f32("l.S_t", B, Gp); S_i = self.f32("l.S_i", B, Gp)
        ops.gemm(Ts, GIs, S_t); ops.gemm(Is, GTs, S_i)
        st = {"T": text_embeds, "I": image_embeds, "off": label_offset, "G": G, "Gp": Gp, "Ts": Ts, "Is": Is, "GIs": GIs, "GTs": GTs,
              "S_t": S_t, "S_i": S_i}
        st["lse_t"] = self.f32("l.lse_t", B); st["lse_i"] = self.f32("l.lse_i", B)
        rows_t = self.f32("l.rows_t", B); rows_i = self.f32("l.rows_i", B)
        ops.ce_rows_fwd(S_t, ls, label_offset, st["lse_t"], rows_t, B, G)     # S_* now hold the scaled logits
        ops.ce_rows_fwd(S_i, ls, label_offset, st["lse_i"], rows_i, B, G)
        st["loss_sum"] = self.f32("l.loss", 1)       # sum over LOCAL rows of both directions / (2 G)
        ops.reduce_sum(rows_t, B, 1.0 / (2 * G), st["loss_sum"], False)
        ops.reduce_sum(rows_i, B, 1.0 / (2 * G), st["loss_sum"], True)
        # scaled logits of the two strips: local texts x image gallery, local images x text gallery.  World size 1: S_i == S_t^T.
        st["logits"] = S_t[:, :G] if want_logits else None
        st["logits_img"] = S_i[:, :G] if want_logits else None
        return st

    def loss_backward(self, st, grad_scale: float = 1.0, local_gallery: bool = True):
        """d(grad_scale * loss) w.r.t. the local embeddings.  With local_gallery (world size 1) the gallery IS the local batch and
        its gradient accumulates straight onto dI / dT; otherwise the gallery gradients (all G rows) are returned separately so
        the distributed wrapper can reduce-scatter them to their owners."""
        B = st["T"].shape[0]; G = st["G"]; Gp = st["Gp"]; E = self.E
        coef = grad_scale / (2.0 * G)
        ls = self.params.p("logit_scale").view(-1); dls = self.params.g("logit_scale").view(1)
        dS_t = self.bf("l.dS_t", B, Gp); dS_i = self.bf("l.dS_i", B, Gp)
        ops.ce_rows_bwd(st["S_t"], ls, st["lse_t"], st["off"], coef, dS_t, B, G, dscale_log=dls)
        ops.ce_rows_bwd(st["S_i"], ls, st["lse_i"], st["off"], coef, dS_i, B, G, dscale_log=dls)
        hi = lambda t: t[:, :E]                                  # the bf16 "hi" block of a split operand
        dT = self.f32("l.dT", B, E); dI = self.f32("l.dI", B, E)
        ops.gemm(dS_t, hi(st["GIs"]), dT, b_mn_major=1)          # own rows as queries: dT = dS_t GI, dI = dS_i GT
        ops.gemm(dS_i, hi(st["GTs"]), dI, b_mn_major=1)
        if local_gallery:
            # gallery rows: dGI = dS_t^T T (+= onto dI), dGT = dS_i^T I (+= onto dT)
            ops.gemm(dS_t[:, :G], hi(st["Ts"]), dI, a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
            ops.gemm(dS_i[:, :G], hi(st["Is"]), dT, a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
            return dT, dI, None, None
        if st.get("peer"):      # gallery gradients go straight into the peer buffer, from where their owners pull them
            dGI, dGT = self._peer.grad_views()
        else:
            dGI = self.f32("l.dGI", G, E); dGT = self.f32("l.dGT", G, E)
        ops.gemm(dS_t[:, :G], hi(st["Ts"]), dGI, a_mn_major=1, b_mn_major=1)
        ops.gemm(dS_i[:, :G], hi(st["Is"]), dGT, a_mn_major=1, b_mn_major=1)
        return dT, dI, dGI, dGT

    # ------------------------------------------------------------------ public steps
    def encode(self, pixels: Optional[torch.Tensor], ids: Optional[torch.Tensor], token_type_ids=None, attention_mask=None):
        """feat=True path of CLIPApp.forward (model.py:145-146): embeddings only, no activations kept."""
        out = {"image_embeds": None, "text_embeds": None}
        if pixels is not None:
            out["image_embeds"] = self.vit_forward(pixels, save=False)["embeds"]
        if ids is not None:
            out["text_embeds"] = self.bert_forward(ids, save=False, token_type_ids=token_type_ids, attention_mask=attention_mask)["embeds"]
        return out

    def _video_pool(self, v, video_masks, B, T):
        """frame embeddings [B*T, E] (each l2-normalised) -> masked mean over the frames -> l2-normalised video embeddings [B, E]
        (appzoo/text2video_retrieval/model.py:82-88)"""
        E = self.E
        vm = video_masks.to(self.dev).long().contiguous()
        vf = self.f32("vd.feat", B, E)
        ops.frame_pool_fwd(v["embeds"], vm, vf, B, T, E)
        ve = self.f32("vd.embeds", B, E); vn_ = self.f32("vd.norm", B)
        ops.l2norm_fwd(vf, ve, vn_, B, E)
        return {"mask": vm, "T": T, "embeds": ve, "norm": vn_}

    def forward(self, pixels, ids, save=True, want_logits=True, distributed=False, train=None, token_type_ids=None, attention_mask=None,
                video_masks=None):
        """Both towers + the contrastive head.  distributed=True: all-gather the embedding shards over the default process
        group and take the loss over the GLOBAL batch (labels offset by rank * local_B); 'loss' is then this rank's share
        (sum over ranks = global loss) and 'logits_per_text' the local [b, G] strip."""
        from . import distributed as D
        if train is None:
            train = save           # training step <=> activations are kept; dropout is active only then
        if train and (self.p_hidden > 0.0 or self.p_attn > 0.0):
            ops.counter_add(self._dev_pass, 1)       # new dropout masks for this pass (its backward sees the same value)
        dist_on = bool(distributed) and D.world_size() > 1
        B = pixels.shape[0]
        # Training steps on several GPUs exchange the embedding shards through peer memory: the l2-normalise kernel of each tower stores
        # its rows straight into every rank's gallery (fused producer + all-gather over NVLink), backward pulls the gallery gradients
        # (distributed.PeerGroup, csrc/peer.cu).  Forward-only calls and backends without peer access use torch.distributed collectives.
        self._peer_on = False
        if dist_on and save and not getattr(self, "_no_peer", False):
            if self._peer_key != (B, self.E):
                self._peer = D.PeerGroup.create(B, self.E, self.dev); self._peer_key = (B, self.E)
            self._peer_on = self._peer is not None
        if video_masks is not None:      # Text2VideoRetrieval: [B, T, 3, R, R] frames through the image tower, masked mean over the frames
            if dist_on:
                raise NotImplementedError("video batches use the single-process loss")
            Bv, T = pixels.shape[0], pixels.shape[1]
            v = self.vit_forward(pixels.reshape(Bv * T, *pixels.shape[2:]).contiguous(), save and not self.hf)
            v["video"] = self._video_pool(v, video_masks, Bv, T)
            img_embeds = v["video"]["embeds"]
        else:
            v = self.vit_forward(pixels, save and not self.hf)       # huggingface_clip: the image tower is frozen -> no activations kept
            img_embeds = v["embeds"]
        t = self.bert_forward(ids, save, train=train, token_type_ids=token_type_ids, attention_mask=attention_mask)
        if dist_on:
            Wd = D.world_size()
            if self._peer_on:
                self._peer.sync(D.PeerGroup.CH_GALLERY)
                gi, gt = self._peer.gallery_views()
            else:
                gi = D.gather_rows(v["embeds"], self.f32("l.gi", Wd * B, self.E))
                gt = D.gather_rows(t["embeds"], self.f32("l.gt", Wd * B, self.E))
            l = self.loss_forward(t["embeds"], v["embeds"], gi, gt, label_offset=D.get_rank() * B, want_logits=want_logits)
            l["dist"] = True; l["peer"] = self._peer_on
        else:
            l = self.loss_forward(t["embeds"], img_embeds, want_logits=want_logits)
            l["dist"] = False; l["peer"] = False
        self._peer_on = False
        self._saved = (v, t, l) if save else None
        if video_masks is not None:
            return {"video_embeds": v["video"]["embeds"], "text_embeds": t["embeds"], "logits_per_text": l["logits"], "logits_per_image": l["logits_img"],
                    "loss": l["loss_sum"], "distributed": l["dist"]}
        return {"image_embeds": v["embeds"], "text_embeds": t["embeds"], "logits_per_text": l["logits"], "logits_per_image": l["logits_img"],
                "loss": l["loss_sum"], "distributed": l["dist"]}

    def zero_grad(self):
        self.params.grad.zero_()

    def backward(self, grad_scale: float = 1.0):
        """Gradient of grad_scale * loss into params.grad (accumulating)."""
        if self._saved is None:
            raise RuntimeError("backward() needs a forward(save=True) first")
        v, t, l = self._saved
        if l["dist"]:
            from . import distributed as D
            dT, dI, dGI, dGT = self.loss_backward(l, grad_scale, local_gallery=False)
            B = dT.shape[0]
            if l.get("peer"):      # reduce-scatter by peer loads: this rank's rows of every rank's gallery gradient, summed onto dI / dT
                self._peer.sync(D.PeerGroup.CH_GRADS)
                self._peer.reduce_rows("image", dI, accumulate=True)
                self._peer.reduce_rows("text", dT, accumulate=True)
            else:
                ops.axpy(D.reduce_scatter_rows(dGI, self.f32("l.rsI", B, self.E)), dI)
                ops.axpy(D.reduce_scatter_rows(dGT, self.f32("l.rsT", B, self.E)), dT)
        else:
            dT, dI, _, _ = self.loss_backward(l, grad_scale, local_gallery=True)
        self.bert_backward(t, dT)
        if v.get("video") is not None:      # video embeds <- l2norm <- masked frame mean <- per-frame l2-normalised image embeds
            vd = v["video"]; Bv = vd["embeds"].shape[0]; T = vd["T"]
            dvf = self.f32("vd.dfeat", Bv, self.E); dframe = self.f32("vd.dframe", Bv * T, self.E)
            ops.l2norm_bwd(dI, vd["embeds"], vd["norm"], dvf, None, Bv, self.E)
            ops.frame_pool_bwd(dvf, vd["mask"], dframe, Bv, T, self.E)
            dI = dframe
        self.vit_backward(v, dI)
        self._saved = None

    def allreduce_grads(self):
        """SUM over ranks of the flat fp32 gradient (the loss is already normalised by the global batch)."""
        from . import distributed as D
        D.allreduce_sum_(self.params.grad)

    def _grads_ready(self, prefix: str):
        """backward hook: every gradient under `prefix` is final -> hand its slices to the overlapped all-reduce (if one is active)"""
        if self._reducer is not None:
            self._reducer.ready(self.params.ranges_for(prefix))

    def set_step(self, n: int):
        """Number of optimizer steps already taken (checkpoint resume): the host counter AND the device-resident one that feeds Adam's
        bias correction, the on-device learning-rate schedule and the dropout stream."""
        self.params.step = int(n)
        self._dev_step.fill_(int(n))
        self._dev_pass.fill_(int(n))

    def set_micro_step(self, n: int):
        """Number of training forward passes already made (resume with gradient accumulation): position of the dropout stream."""
        self._dev_pass.fill_(int(n))

    def optimizer_step(self, lr: float, weight_decay: float = 1e-4, max_grad_norm: float = 1.0, warmup_steps: int = 0, t_total: int = 0):
        """clip_grad_norm_(max_grad_norm) + AdamW(betas 0.9/0.999, eps 1e-6) with the reference's decay grouping.
        The step counter and {lr, bias-corrected step size} live on the device (clipk_adam_schedule), so the same launches can be
        replayed from a CUDA graph.  t_total == 0: `lr` is the learning rate of THIS step (host-side scheduler, Trainer);
        t_total > 0: `lr` is the base rate and the reference's warmup-linear schedule is evaluated on the device."""
        P_ = self.params
        P_.step += 1
        n_tr, n_dec = P_.n_trainable, P_.n_decay
        ops.adam_schedule(self._dev_step, self._dev_hyper, float(lr), int(warmup_steps), int(t_total))
        coef = None
        if max_grad_norm and max_grad_norm > 0:
            ops.grad_norm(P_.grad, n_tr, float(max_grad_norm), self._norm_ws, self.norm_and_coef)
            coef = self.norm_and_coef[1:]
        ops.adamw_step(P_.master[:n_dec], P_.grad[:n_dec], P_.exp_avg[:n_dec], P_.exp_avg_sq[:n_dec], P_.shadow[:n_dec], n_dec,
                       lr, weight_decay, 0, coef, dev_hyper=self._dev_hyper)
        ops.adamw_step(P_.master[n_dec:n_tr], P_.grad[n_dec:n_tr], P_.exp_avg[n_dec:n_tr], P_.exp_avg_sq[n_dec:n_tr],
                       P_.shadow[n_dec:n_tr], n_tr - n_dec, lr, 0.0, 0, coef, dev_hyper=self._dev_hyper)
        P_.version += 1

    # ------------------------------------------------------------------ whole training step, optionally as ONE CUDA graph
    def _step_body(self, pixels, ids, hp):
        self.zero_grad()
        out = self.forward(pixels, ids, save=True, want_logits=hp["want_logits"], distributed=hp["distributed"],
                           token_type_ids=hp.get("token_type_ids"), attention_mask=hp.get("attention_mask"))
        if hp["allreduce"] and hp.get("overlap", False):
            # eager multi-GPU step: per-layer gradient slices are all-reduced while the rest of the backward pass runs
            from . import distributed as D
            self._reducer = D.OverlappedGradReducer(self.params.grad, self.params.n_trainable)
            try:
                self.backward(hp["grad_scale"])
                self._reducer.finish()
            finally:
                self._reducer = None
        else:
            self.backward(hp["grad_scale"])
            if hp["allreduce"]:
                self.allreduce_grads()
        self.optimizer_step(hp["lr"], hp["weight_decay"], hp["max_grad_norm"], hp["warmup_steps"], hp["t_total"])
        return out

    def train_step(self, pixels, ids, lr, weight_decay=1e-4, max_grad_norm=1.0, warmup_steps=0, t_total=0, distributed=False,
                   want_logits=False, use_graph=True, grad_scale=None, allreduce=None, token_type_ids=None, attention_mask=None):
        """zero_grad -> forward -> backward -> (grad all-reduce) -> clip + AdamW.  `pixels` / `ids` may live on the host (pinned):
        they are copied into static device buffers.  After two eager calls (which size every buffer and set kernel attributes) the
        whole step is captured once into a CUDA graph and replayed: ~2.6 k kernel launches collapse into one cudaGraphLaunch.
        With use_graph, `lr` must be the BASE rate and the schedule (warmup_steps, t_total) is evaluated on the device.
        Returns the dict of forward(); tensors are engine-owned buffers, valid until the next call."""
        from . import distributed as D
        if allreduce is None:
            allreduce = D.world_size() > 1
        if grad_scale is None:      # local-loss data parallelism averages gradients like DDP; the global loss is already normalised
            grad_scale = 1.0 if (distributed or D.world_size() == 1) else 1.0 / D.world_size()
        B, Lt = ids.shape
        extra = self.hf and (token_type_ids is not None, attention_mask is not None)
        key = (B, Lt, bool(allreduce), float(lr), float(weight_decay), float(max_grad_norm), int(warmup_steps), int(t_total), bool(distributed), bool(want_logits), float(grad_scale), extra)
        st = getattr(self, "_gs", None)
        if st is None or st["key"] != key:
            st = {"key": key, "calls": 0, "graph": None, "out": None,
                  "pixels": torch.empty((B, 3, self.R, self.R), dtype=torch.float32, device=self.dev),
                  "ids": torch.empty((B, Lt), dtype=torch.int64, device=self.dev)}
            self._gs = st
        st["pixels"].copy_(pixels, non_blocking=True)
        st["ids"].copy_(ids, non_blocking=True)
        tt = am = None
        if self.hf:       # the text tower of the huggingface_clip branch reads the batch's token_type_ids / attention_mask (static buffers too)
            if token_type_ids is not None:
                tt = st.setdefault("tt", torch.empty((B, Lt), dtype=torch.int64, device=self.dev)); tt.copy_(token_type_ids, non_blocking=True)
            if attention_mask is not None:
                am = st.setdefault("am", torch.empty((B, Lt), dtype=torch.int64, device=self.dev)); am.copy_(attention_mask, non_blocking=True)
        hp = {"token_type_ids": tt, "attention_mask": am, "lr": lr, "weight_decay": weight_decay, "max_grad_norm": max_grad_norm, "warmup_steps": warmup_steps, "t_total": t_total,
              "distributed": distributed, "want_logits": want_logits, "grad_scale": grad_scale, "allreduce": allreduce,
              "overlap": bool(allreduce) and not use_graph and os.environ.get("CLIPK_NO_OVERLAP", "0") != "1"}
        # the flag epochs of the peer-memory exchange are host-side launch arguments: a replayed graph would reuse them, so a captured
        # multi-GPU step takes the torch.distributed collectives instead
        self._no_peer = bool(use_graph)
        if not use_graph or st["calls"] < 2:
            st["calls"] += 1
            return self._step_body(st["pixels"], st["ids"], hp)
        if st["graph"] is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            host_step = self.params.step
            with torch.cuda.graph(g):
                st["out"] = self._step_body(st["pixels"], st["ids"], hp)
            self.params.step = host_step      # capture does not execute
            st["graph"] = g
        self.params.step += 1
        st["graph"].replay()
        return st["out"]
