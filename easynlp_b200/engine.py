"""ClipEngine: host-side schedule of the chinese_clip training / encoding step over the clipk kernels.

Everything numerical happens in hand-written sm_100a kernels behind the C ABI (include/clipk.h); this file only owns
buffers (torch tensors as device memory) and the order of launches, i.e. it is the from-scratch counterpart of the
autograd graph PyTorch builds for

    CLIPApp.forward / compute_loss      easynlp/appzoo/clip/model.py:106-164
    CHINESE_CLIP.forward                easynlp/modelzoo/models/clip/modeling_chineseclip.py:343-365
    VisualTransformer / ResidualAttentionBlock   .../modeling_chineseclip.py:170-253   (pre-LN, QuickGELU)
    BertModel (embeddings + post-LN encoder)     easynlp/modelzoo/models/bert/modeling_bert.py:72-541
    clip_grad_norm_ + AdamW.step        easynlp/core/trainer.py:315-337, easynlp/core/optimizers.py:405-464

Numerics: bf16 GEMM/attention operands with fp32 accumulation, fp32 residual stream / LayerNorm statistics /
embeddings / logits / loss, fp32 master weights and gradients.  Layout: token-major [B*L, d] activations.
BERT-tower dropout (hidden / attention-probs) is fused into the LayerNorm and attention kernels as Philox masks keyed by a
device-resident forward-pass counter: a training forward draws fresh masks, its backward regenerates them ("dropout").
"""
import math
from typing import Dict, Optional

import os

import torch

from . import _lib as L
from . import ops
from .params import ParamStore

SM_TARGET = 148 * 2


def _splits_for(m, n, k):
    tiles = ((m + 127) // 128) * ((n + 255) // 256 if (n % 256 == 0 or n > 512) else (n + 127) // 128)
    s = max(1, SM_TARGET // max(1, tiles))
    return max(1, min(s, max(1, k // 256)))


class ClipEngine:
    def __init__(self, cfg: dict, device="cuda", with_optimizer_state: bool = True):
        if isinstance(cfg.get("vision_layers"), (tuple, list)):
            raise NotImplementedError("ModifiedResNet visual towers are outside the hot path (ViT only)")
        self.cfg = cfg
        self.dev = torch.device(device)
        self.W = cfg["vision_width"]; self.P = cfg["vision_patch_size"]; self.R = cfg["image_resolution"]
        self.E = cfg["embed_dim"]; self.g = self.R // self.P; self.Lv = self.g * self.g + 1
        self.Hv = self.W // 64                                   # modeling_chineseclip.py:289
        self.H = cfg["text_hidden_size"]; self.I = cfg["text_intermediate_size"]; self.Ht = cfg["text_num_attention_heads"]
        self.nv = cfg["vision_layers"]; self.nt = cfg["text_num_hidden_layers"]
        if self.H != self.Ht * 64 or self.W % 128 or self.H % 128 or self.E % 128:
            raise NotImplementedError("clipk kernels need head_dim 64 and widths that are multiples of 128")
        if (3 * self.P * self.P) % 8:
            raise NotImplementedError("patch dim 3*P*P must be a multiple of 8 (pad the patch matrix for ViT-*/14)")
        if cfg.get("text_hidden_act", "gelu") != "gelu":
            raise NotImplementedError("text tower activation must be erf-GELU")
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
    def vit_forward(self, pixels: torch.Tensor, save: bool):
        P_ = self.params; W = self.W; B = pixels.shape[0]; Lv = self.Lv; M = B * Lv; Hh = self.Hv
        assert pixels.dtype == torch.float32 and pixels.is_contiguous() and pixels.shape[1:] == (3, self.R, self.R)
        npatch = B * self.g * self.g
        kdim = 3 * self.P * self.P
        patches = self.bf("v.patches", npatch, kdim)
        ops.im2col_patches(pixels, patches, B, self.R, self.P)
        patch_out = self.f32("v.patch_out", npatch, W)
        ops.gemm(patches, P_.w("visual.conv1.weight", (W, kdim)), patch_out)
        x0 = self.f32("v.x0", M, W)
        ops.vit_assemble(patch_out, P_.p("visual.class_embedding"), P_.p("visual.positional_embedding"), x0, B, Lv, W)
        x = self.f32("v.x.0", M, W)
        st = {"B": B, "mean0": self.f32("v.mean0", M), "rstd0": self.f32("v.rstd0", M), "layers": []}
        ops.layernorm_fwd(x0, P_.p("visual.ln_pre.weight"), P_.p("visual.ln_pre.bias"), 1e-5, None, x, st["mean0"], st["rstd0"])
        # Residual stream: every projection GEMM writes its branch output as bf16 (plain epilogue); the fp32 residual add is fused
        # into the LayerNorm that follows it (x_new = x + branch is stored by that kernel for backward / the next residual).
        ybr = self.bf("v.ybr", M, W)                 # branch output (attention out-proj / MLP c_proj), reused
        pending = None                               # (x_res, add): residual add owed to the next LayerNorm
        for i in range(self.nv):
            p = f"visual.transformer.resblocks.{i}."
            tag = f"v.{i}." if save else "v.t."
            ly = {}
            ly["h"] = self.bf(tag + "h", M, W); ly["m1"] = self.f32(tag + "m1", M); ly["r1"] = self.f32(tag + "r1", M)
            if pending is None:
                ops.layernorm_fwd(x, P_.p(p + "ln_1.weight"), P_.p(p + "ln_1.bias"), 1e-5, ly["h"], None, ly["m1"], ly["r1"])
            else:       # x = x1_prev + c_proj(...) of the previous block
                x_new = self.f32(f"v.x.{i}" if save else f"v.x.t{i % 2}", M, W)
                ops.layernorm_fwd(pending, P_.p(p + "ln_1.weight"), P_.p(p + "ln_1.bias"), 1e-5, ly["h"], None, ly["m1"], ly["r1"],
                                  add=ybr, x_out=x_new)
                x = x_new
            ly["x_in"] = x
            ly["qkv"] = self.bf(tag + "qkv", M, 3 * W)
            ops.gemm(ly["h"], P_.w(p + "attn.in_proj_weight"), ly["qkv"], bias=P_.p(p + "attn.in_proj_bias"))
            ly["ctx"] = self.bf(tag + "ctx", M, W); ly["lse"] = self.f32(tag + "lse", B * Hh * Lv)
            ops.attention_fwd(ly["qkv"], None, ly["ctx"], ly["lse"], B, Lv, Hh)
            ops.gemm(ly["ctx"], P_.w(p + "attn.out_proj.weight"), ybr, bias=P_.p(p + "attn.out_proj.bias"))
            ly["x1"] = self.f32(tag + "x1", M, W)
            ly["h2"] = self.bf(tag + "h2", M, W); ly["m2"] = self.f32(tag + "m2", M); ly["r2"] = self.f32(tag + "r2", M)
            ops.layernorm_fwd(x, P_.p(p + "ln_2.weight"), P_.p(p + "ln_2.bias"), 1e-5, ly["h2"], None, ly["m2"], ly["r2"],
                              add=ybr, x_out=ly["x1"])
            # "z" holds act'(z) (QuickGELU derivative) saved for backward, "a" the activation
            ly["z"] = self.bf(tag + "z", M, 4 * W); ly["a"] = self.bf(tag + "a", M, 4 * W)
            ops.gemm(ly["h2"], P_.w(p + "mlp.c_fc.weight"), ly["z"], bias=P_.p(p + "mlp.c_fc.bias"), mode=L.EPI_QUICK_GELU, out2=ly["a"])
            ops.gemm(ly["a"], P_.w(p + "mlp.c_proj.weight"), ybr, bias=P_.p(p + "mlp.c_proj.bias"))
            pending = ly["x1"]
            st["layers"].append(ly)
        # ln_post on the CLS rows of x_final = x1_last + c_proj(...): the add is fused here too; x_cls [B, W] is kept for backward
        st["x_cls"] = self.f32("v.x_cls", B, W)
        st["pooled"] = self.bf("v.pooled", B, W); st["mp"] = self.f32("v.mp", B); st["rp"] = self.f32("v.rp", B)
        if pending is None:      # zero-layer tower (degenerate configs): no pending residual
            ops.layernorm_fwd(x, P_.p("visual.ln_post.weight"), P_.p("visual.ln_post.bias"), 1e-5, st["pooled"], None, st["mp"], st["rp"],
                              rows=B, ldx=Lv * W)
            st["x_cls"] = None; st["x_final"] = x
        else:
            ops.layernorm_fwd(pending, P_.p("visual.ln_post.weight"), P_.p("visual.ln_post.bias"), 1e-5, st["pooled"], None, st["mp"], st["rp"],
                              rows=B, ldx=Lv * W, add=ybr, ldadd=Lv * W, x_out=st["x_cls"])
        st["feat"] = self.f32("v.feat", B, self.E)
        ops.gemm(st["pooled"], P_.w("visual.proj"), st["feat"], b_mn_major=1)
        st["embeds"] = self.f32("v.embeds", B, self.E); st["norm"] = self.f32("v.norm", B)
        ops.l2norm_fwd(st["feat"], st["embeds"], st["norm"], B, self.E)
        return st

    def vit_backward(self, st, d_embeds: torch.Tensor):
        P_ = self.params; W = self.W; B = st["B"]; Lv = self.Lv; M = B * Lv; Hh = self.Hv; E = self.E
        dfeat_b = self.bf("v.dfeat_b", B, E)
        ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], None, dfeat_b, B, E)
        # d proj[W,E] += pooled^T dfeat ; dpooled = dfeat proj^T
        ops.gemm(st["pooled"], dfeat_b, P_.g("visual.proj"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
        dpooled = self.f32("v.dpooled", B, W)
        ops.gemm(dfeat_b, P_.w("visual.proj"), dpooled)
        dX = self.f32("v.dX.a", M, W); dXb = self.bf("v.dXb.a", M, W)
        dX.zero_()
        last = st["layers"][-1] if self.nv else None
        bias_prev = P_.g(f"visual.transformer.resblocks.{self.nv - 1}.mlp.c_proj.bias") if self.nv else None
        x_post, ld_post = (st["x_cls"], W) if st.get("x_cls") is not None else (st["x_final"], Lv * W)
        ops.layernorm_bwd(dpooled, x_post, P_.p("visual.ln_post.weight"), st["mp"], st["rp"], dx_f32=dX,
                          dgamma=P_.g("visual.ln_post.weight"), dbeta=P_.g("visual.ln_post.bias"), dbias=bias_prev,
                          rows=B, ldx=ld_post, lddx=Lv * W)
        ops.cast_bf16(dX, dXb)
        dz = self.bf("v.dz", M, 4 * W); dh = self.bf("v.dh", M, W); dctx = self.bf("v.dctx", M, W); dqkv = self.bf("v.dqkv", M, 3 * W)
        for i in reversed(range(self.nv)):
            p = f"visual.transformer.resblocks.{i}."
            ly = st["layers"][i]
            # MLP
            ops.gemm(dXb, ly["a"], P_.g(p + "mlp.c_proj.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(W, 4 * W, M))
            ops.gemm(dXb, P_.w(p + "mlp.c_proj.weight"), dz, b_mn_major=1, mode=L.EPI_MUL_AUX, aux=ly["z"],
                     colsum=P_.g(p + "mlp.c_fc.bias"))            # dz = (dX W) o act'(z); its column sums = d(c_fc.bias)
            ops.gemm(dz, ly["h2"], P_.g(p + "mlp.c_fc.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(4 * W, W, M))
            ops.gemm(dz, P_.w(p + "mlp.c_fc.weight"), dh, b_mn_major=1)
            dX1 = self.f32("v.dX.b", M, W); dX1b = self.bf("v.dXb.b", M, W)
            ops.layernorm_bwd(dh, ly["x1"], P_.p(p + "ln_2.weight"), ly["m2"], ly["r2"], dx_add=dX, dx_f32=dX1, dx_bf16=dX1b,
                              dgamma=P_.g(p + "ln_2.weight"), dbeta=P_.g(p + "ln_2.bias"), dbias=P_.g(p + "attn.out_proj.bias"))
            # attention
            ops.gemm(dX1b, ly["ctx"], P_.g(p + "attn.out_proj.weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(W, W, M))
            ops.gemm(dX1b, P_.w(p + "attn.out_proj.weight"), dctx, b_mn_major=1)
            ops.attention_bwd(ly["qkv"], None, ly["ctx"], ly["lse"], dctx, dqkv, B, Lv, Hh, dqkv_colsum=P_.g(p + "attn.in_proj_bias"))
            ops.gemm(dqkv, ly["h"], P_.g(p + "attn.in_proj_weight"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD,
                     splits=_splits_for(3 * W, W, M))
            ops.gemm(dqkv, P_.w(p + "attn.in_proj_weight"), dh, b_mn_major=1)
            bias_prev = P_.g(f"visual.transformer.resblocks.{i - 1}.mlp.c_proj.bias") if i > 0 else None
            ops.layernorm_bwd(dh, ly["x_in"], P_.p(p + "ln_1.weight"), ly["m1"], ly["r1"], dx_add=dX1, dx_f32=dX, dx_bf16=dXb,
                              dgamma=P_.g(p + "ln_1.weight"), dbeta=P_.g(p + "ln_1.bias"), dbias=bias_prev)
            self._grads_ready(p)
        # ln_pre, token assembly, patch embedding
        dx0 = self.f32("v.dx0", M, W)
        ops.layernorm_bwd(dX, self.f32("v.x0", M, W), P_.p("visual.ln_pre.weight"), st["mean0"], st["rstd0"], dx_f32=dx0,
                          dgamma=P_.g("visual.ln_pre.weight"), dbeta=P_.g("visual.ln_pre.bias"))
        ops.colsum(dx0, P_.g("visual.positional_embedding").view(-1), B, Lv * W)
        ops.colsum(dx0, P_.g("visual.class_embedding"), B, W, ldx=Lv * W)
        npatch = B * self.g * self.g; kdim = 3 * self.P * self.P
        dpatch = self.bf("v.dpatch", npatch, W)
        ops.vit_assemble_bwd(dx0, dpatch, B, Lv, W)
        ops.gemm(dpatch, self.bf("v.patches", npatch, kdim), P_.g("visual.conv1.weight", (W, kdim)), a_mn_major=1, b_mn_major=1,
                 mode=L.EPI_ATOMIC_ADD, splits=_splits_for(W, kdim, npatch))

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

    def bert_forward(self, ids: torch.Tensor, save: bool, train: bool = False):
        P_ = self.params; H = self.H; I = self.I; B, Lt = ids.shape; M = B * Lt; Hh = self.Ht
        assert ids.dtype == torch.int64 and ids.is_contiguous()
        if Lt > self.cfg["text_max_position_embeddings"]:
            raise ValueError("sequence longer than the position table")
        eps = 1e-12                                                   # modeling_chineseclip.py:311
        st = {"B": B, "Lt": Lt, "ids": ids, "layers": [], "train": train}
        st["e"] = self.f32("t.e", M, H); st["mask"] = self.f32("t.mask", M)
        ops.bert_embed(ids.view(-1), P_.p("bert.embeddings.word_embeddings.weight"), P_.p("bert.embeddings.position_embeddings.weight"),
                       P_.p("bert.embeddings.token_type_embeddings.weight"), st["e"], M, Lt, H, self.cfg["vocab_size"], key_mask=st["mask"])
        x = self.f32("t.x.0", M, H); xb = self.bf("t.xb.0", M, H)
        st["me"] = self.f32("t.me", M); st["re"] = self.f32("t.re", M)
        ops.layernorm_fwd(st["e"], P_.p("bert.embeddings.LayerNorm.weight"), P_.p("bert.embeddings.LayerNorm.bias"), eps, xb, x,
                          st["me"], st["re"], drop=self._drop(train, self.p_hidden, 1), drop_mode=2)
        tbr = self.bf("t.tbr", M, H)                 # bf16 branch output of attention.output.dense / output.dense
        for i in range(self.nt):
            p = f"bert.encoder.layer.{i}."
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
        ops.gemm(cls_rows, P_.w("text_projection"), st["feat"], b_mn_major=1)
        st["embeds"] = self.f32("t.embeds", B, self.E); st["norm"] = self.f32("t.norm", B)
        ops.l2norm_fwd(st["feat"], st["embeds"], st["norm"], B, self.E)
        return st

    def bert_backward(self, st, d_embeds: torch.Tensor):
        P_ = self.params; H = self.H; I = self.I; B = st["B"]; Lt = st["Lt"]; M = B * Lt; Hh = self.Ht; E = self.E
        train = st.get("train", False)
        dfeat_b = self.bf("t.dfeat_b", B, E)
        ops.l2norm_bwd(d_embeds, st["embeds"], st["norm"], None, dfeat_b, B, E)
        cls_rows = st["xb_final"].view(B, Lt * H)[:, :H]
        ops.gemm(cls_rows, dfeat_b, P_.g("text_projection"), a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD)
        dOut = self.f32("t.dOut", M, H)
        dOut.zero_()
        ops.gemm(dfeat_b, P_.w("text_projection"), dOut.view(B, Lt * H)[:, :H])
        dy, dy_add = dOut, None
        ds2 = self.f32("t.ds2", M, H); ds2b = self.bf("t.ds2b", M, H); ds1 = self.f32("t.ds1", M, H); ds1b = self.bf("t.ds1b", M, H)
        dz = self.bf("t.dz", M, I); dpart = self.bf("t.dpart", M, H); dctx = self.bf("t.dctx", M, H); dqkv = self.bf("t.dqkv", M, 3 * H)
        dxp = self.bf("t.dxp", M, H)
        for i in reversed(range(self.nt)):
            p = f"bert.encoder.layer.{i}."
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
        ops.layernorm_bwd(dy, st["e"], P_.p("bert.embeddings.LayerNorm.weight"), st["me"], st["re"], dy_add=dy_add, dx_f32=de,
                          dgamma=P_.g("bert.embeddings.LayerNorm.weight"), dbeta=P_.g("bert.embeddings.LayerNorm.bias"),
                          drop=self._drop(train, self.p_hidden, 1), drop_mode=2)
        ops.bert_embed_bwd(st["ids"].view(-1), de, P_.g("bert.embeddings.word_embeddings.weight"), M, H, self.cfg["vocab_size"])
        ops.colsum(de, P_.g("bert.embeddings.position_embeddings.weight").view(-1)[:Lt * H], B, Lt * H)
        ops.colsum(de, P_.g("bert.embeddings.token_type_embeddings.weight")[0], M, H)
        self._grads_ready("bert.embeddings.")

    # ------------------------------------------------------------------ contrastive head
    def loss_forward(self, text_embeds, image_embeds, gallery_image=None, gallery_text=None, label_offset=0, want_logits=True):
        """Two CE strips: local texts vs the image gallery, local images vs the text gallery (appzoo/clip/model.py:148-164).  With no
        gallery given the local batch is the gallery (world size 1: identical to the reference's [B,B] loss).
        The dots come from the tcgen05 GEMM on bf16 hi/lo splits of the fp32 embeddings (K = 3E, fp32-level logits; loss.cu)."""
        gi = image_embeds if gallery_image is None else gallery_image
        gt = text_embeds if gallery_text is None else gallery_text
        B = text_embeds.shape[0]; G = gi.shape[0]; E = self.E
        Gp = (G + 7) // 8 * 8                                   # GEMM N / leading dimensions are multiples of 8; padding rows stay zero
        ls = self.params.p("logit_scale")
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
        ls = self.params.p("logit_scale"); dls = self.params.g("logit_scale").view(1)
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
        dGI = self.f32("l.dGI", G, E); dGT = self.f32("l.dGT", G, E)
        ops.gemm(dS_t[:, :G], hi(st["Ts"]), dGI, a_mn_major=1, b_mn_major=1)
        ops.gemm(dS_i[:, :G], hi(st["Is"]), dGT, a_mn_major=1, b_mn_major=1)
        return dT, dI, dGI, dGT

    # ------------------------------------------------------------------ public steps
    def encode(self, pixels: Optional[torch.Tensor], ids: Optional[torch.Tensor]):
        """feat=True path of CLIPApp.forward (model.py:145-146): embeddings only, no activations kept."""
        out = {"image_embeds": None, "text_embeds": None}
        if pixels is not None:
            out["image_embeds"] = self.vit_forward(pixels, save=False)["embeds"]
        if ids is not None:
            out["text_embeds"] = self.bert_forward(ids, save=False)["embeds"]
        return out

    def forward(self, pixels, ids, save=True, want_logits=True, distributed=False, train=None):
        """Both towers + the contrastive head.  distributed=True: all-gather the embedding shards over the default process
        group and take the loss over the GLOBAL batch (labels offset by rank * local_B); 'loss' is then this rank's share
        (sum over ranks = global loss) and 'logits_per_text' the local [b, G] strip."""
        from . import distributed as D
        if train is None:
            train = save           # training step <=> activations are kept; dropout is active only then
        if train and (self.p_hidden > 0.0 or self.p_attn > 0.0):
            ops.counter_add(self._dev_pass, 1)       # new dropout masks for this pass (its backward sees the same value)
        v = self.vit_forward(pixels, save)
        t = self.bert_forward(ids, save, train=train)
        if distributed and D.world_size() > 1:
            B = pixels.shape[0]; Wd = D.world_size()
            gi = D.gather_rows(v["embeds"], self.f32("l.gi", Wd * B, self.E))
            gt = D.gather_rows(t["embeds"], self.f32("l.gt", Wd * B, self.E))
            l = self.loss_forward(t["embeds"], v["embeds"], gi, gt, label_offset=D.get_rank() * B, want_logits=want_logits)
            l["dist"] = True
        else:
            l = self.loss_forward(t["embeds"], v["embeds"], want_logits=want_logits)
            l["dist"] = False
        self._saved = (v, t, l) if save else None
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
            ops.axpy(D.reduce_scatter_rows(dGI, self.f32("l.rsI", B, self.E)), dI)
            ops.axpy(D.reduce_scatter_rows(dGT, self.f32("l.rsT", B, self.E)), dT)
        else:
            dT, dI, _, _ = self.loss_backward(l, grad_scale, local_gallery=True)
        self.bert_backward(t, dT)
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

    # ------------------------------------------------------------------ whole training step, optionally as ONE CUDA graph
    def _step_body(self, pixels, ids, hp):
        self.zero_grad()
        out = self.forward(pixels, ids, save=True, want_logits=hp["want_logits"], distributed=hp["distributed"])
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
                   want_logits=False, use_graph=True, grad_scale=None, allreduce=None):
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
        key = (B, Lt, bool(allreduce), float(lr), float(weight_decay), float(max_grad_norm), int(warmup_steps), int(t_total), bool(distributed), bool(want_logits), float(grad_scale))
        st = getattr(self, "_gs", None)
        if st is None or st["key"] != key:
            st = {"key": key, "calls": 0, "graph": None, "out": None,
                  "pixels": torch.empty((B, 3, self.R, self.R), dtype=torch.float32, device=self.dev),
                  "ids": torch.empty((B, Lt), dtype=torch.int64, device=self.dev)}
            self._gs = st
        st["pixels"].copy_(pixels, non_blocking=True)
        st["ids"].copy_(ids, non_blocking=True)
        hp = {"lr": lr, "weight_decay": weight_decay, "max_grad_norm": max_grad_norm, "warmup_steps": warmup_steps, "t_total": t_total,
              "distributed": distributed, "want_logits": want_logits, "grad_scale": grad_scale, "allreduce": allreduce,
              "overlap": bool(allreduce) and not use_graph and os.environ.get("CLIPK_NO_OVERLAP", "0") != "1"}
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
