"""Synthetic checkpoints and batches for benchmarks / smoke runs (no datasets or pretrained weights are reachable here).

Shapes follow SURVEY.md 8(d): pixels ~ N(0,1) fp32 [B,3,R,R] (what CLIPDataset's normalisation yields, appzoo/clip/data.py:102-135),
ids int64 [B,L] with [CLS]=101 first, uniform tokens for the first len_i ~ U{8..L} positions and 0-padding after, so the
``ids != 0`` mask path (modeling_chineseclip.py:347-348) is exercised.  Weight init follows the reference's init laws
(modeling_chineseclip.py:226-233,316-341; BERT normal(0, initializer_range))."""
import math

import torch


def random_state_dict_hf(cfg: dict, seed: int = 1234, device="cpu"):
    """random-init checkpoint of the huggingface_clip branch (flat engine config from engine.hf_engine_config): normal(0, 0.02) matrices,
    width^-1/2 projections / class / position embeddings, unit LayerNorm gains, zero biases, zero padding rows"""
    from .params import hf_param_schema
    g = torch.Generator(device=device).manual_seed(seed)
    W = cfg["vision_width"]; H = cfg["text_hidden_size"]
    sd = {}
    for name, shape in hf_param_schema(cfg).items():
        if name == "logit_scale":
            t = torch.full(shape, math.log(1 / 0.07), device=device)
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("LayerNorm.weight"):
            t = torch.ones(shape, device=device)
        elif name.endswith("bias"):
            t = torch.zeros(shape, device=device)
        else:
            std = 0.02
            if name.startswith("vision_projection") or "class_embedding" in name or "position_embedding.weight" in name:
                std = W ** -0.5
            elif name.startswith("text_projection") or "pooler" in name:
                std = H ** -0.5
            elif "patch_embedding" in name:
                std = (3 * cfg["vision_patch_size"] ** 2) ** -0.5
            t = torch.randn(shape, generator=g, device=device) * std
        sd[name] = t
    pad = cfg.get("text_pad_token_id", 0)
    sd["text_encoder.embeddings.word_embeddings.weight"][pad].zero_()
    sd["text_encoder.embeddings.position_embeddings.weight"][pad].zero_()
    return sd


def random_state_dict(cfg: dict, seed: int = 1234, device="cpu"):
    if cfg.get("model_type") == "huggingface_clip":
        return random_state_dict_hf(cfg, seed, device)
    from .params import param_schema
    g = torch.Generator(device=device).manual_seed(seed)
    W = cfg["vision_width"]; H = cfg["text_hidden_size"]; r = cfg["text_initializer_range"]
    nl = max(1, cfg["vision_layers"])
    sd = {}
    for name, shape in param_schema(cfg).items():
        if name == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07), device=device)
        elif name.endswith("LayerNorm.weight") or ".ln_" in name and name.endswith(".weight"):
            t = torch.ones(shape, device=device)
        elif name.endswith("bias"):
            t = torch.zeros(shape, device=device)
        else:
            if name.startswith("visual."):
                if name.endswith("conv1.weight"):
                    std = (3 * cfg["vision_patch_size"] ** 2) ** -0.5
                elif "out_proj" in name or "c_proj" in name:
                    std = W ** -0.5 * (2 * nl) ** -0.5
                elif "c_fc" in name:
                    std = (2 * W) ** -0.5
                else:
                    std = W ** -0.5
            elif name == "text_projection":
                std = H ** -0.5
            else:
                std = r
            t = torch.randn(shape, generator=g, device=device) * std
        sd[name] = t
    sd["bert.embeddings.word_embeddings.weight"][0].zero_()
    return sd


def synthetic_batch(cfg: dict, batch: int, seq_len: int = 77, seed: int = 1234, device="cpu", pin: bool = False):
    g = torch.Generator(device=device).manual_seed(seed)
    R = cfg["image_resolution"]; V = cfg["vocab_size"]
    pixels = torch.randn(batch, 3, R, R, generator=g, device=device)
    lens = torch.randint(min(8, seq_len), seq_len + 1, (batch,), generator=g, device=device)
    ids = torch.randint(1, V, (batch, seq_len), generator=g, device=device)
    ids[:, 0] = min(101, V - 1)
    pos = torch.arange(seq_len, device=device).unsqueeze(0)
    ids = torch.where(pos < lens.unsqueeze(1), ids, torch.zeros_like(ids))
    if pin and device == "cpu":
        pixels = pixels.pin_memory(); ids = ids.pin_memory()
    return pixels, ids
