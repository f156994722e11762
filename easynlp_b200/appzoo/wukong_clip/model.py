"""WukongCLIP -- drop-in for easynlp/appzoo/wukong_clip/model.py:22-73 (+ WukongModel, modelzoo/models/wukong/modeling_wukong.py:234-413):
a ViT image tower and a causal pre-LN text transformer (LayerNorm eps 1e-7, token table `embedding_table`, pooling at the [SEP] = 102
position), both on the clipk kernels through the engine's `wukong` kind -- a sibling application of the CLIP path.

Contract kept from the reference: `config.json` = {"model": {"visual": {...}, "text": {...}}}, `pytorch_model.bin` keys under
`model.visual_encoder.` / `model.text_encoder.` / `model.logit_scale` (:363-380; the .pt / pickle loaders of WukongModel are conversion
utilities and are not reproduced); `forward(inputs)` returns the TUPLE `({'image_features', 'text_features', 'logit_scale'}, [])` with
either modality optional (:58-71); `compute_loss` is the symmetric InfoNCE over `logit_scale * image_features @ text_features.T`
(:73-88) -- evaluated by the fused loss kernels during forward() when both modalities are present, with the engine's hand-written
backward behind `loss.backward()`."""
import json
import os

import torch

from ..clip.model import CLIPApp, _ClipLossFn
from ...engine import ClipEngine, wukong_engine_config


class WukongConfig:
    """configuration_wukong.py:25-38: the raw JSON object under `.data`, serialised as it came"""
    model_type = "wukong"

    def __init__(self, config_obj):
        self.data = config_obj

    def to_json_string(self):
        return json.dumps(self.data)


class WukongCLIP(CLIPApp):

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, user_defined_parameters={}, **kwargs):
        return cls(pretrained_model_name_or_path, user_defined_parameters)

    def __init__(self, pretrained_model_name_or_path=None, user_defined_parameters=None, **kwargs):
        torch.nn.Module.__init__(self)
        self.engine = None
        if pretrained_model_name_or_path is None:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("easynlp_b200.WukongCLIP needs a CUDA device (B200): there is no CPU fallback")
        path = pretrained_model_name_or_path
        with open(os.path.join(path, "config.json"), "r") as f:
            self.raw_config = json.load(f)
        self.config = WukongConfig(self.raw_config)
        checkpoint = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        self.model_type = "wukong"
        self.prefix = "model."
        self.distributed_loss = False
        self.engine = ClipEngine(wukong_engine_config(self.raw_config), device=kwargs.get("device", "cuda"))
        sd = {k[len(self.prefix):]: v for k, v in checkpoint.items() if k.startswith(self.prefix)}
        missing = [n for n in self.engine.params.names() if n not in sd]
        if missing:       # WukongModel loads each tower with load_state_dict(strict=True) (modeling_wukong.py:406-410)
            raise KeyError(f"Wukong checkpoint lacks {len(missing)} tensors, e.g. {missing[:3]}")
        self.engine.params.load_state_dict(sd, strict=False)
        self._wrap_params()

    def forward(self, inputs, feat=None):
        eng = self.engine; dev = eng.dev
        pix = inputs.get("pixel_values"); ids = inputs.get("input_ids")
        if pix is not None:
            pix = pix.to(dev, non_blocking=True).float().contiguous()
        if ids is not None:
            if not ids.is_cuda and not bool(((ids == eng.sep_id).sum(1) == 1).all()):
                # `x[(ids == 102).nonzero()]` (modeling_wukong.py:349,359) yields one row per [SEP]; anything but one per text breaks the batch
                raise ValueError(f"every Wukong text needs exactly one [SEP] (id {eng.sep_id})")
            ids = ids.to(dev, non_blocking=True).long().contiguous()
        scale = eng.params.p("logit_scale").reshape(()).exp()
        if pix is not None and ids is not None:
            out = eng.forward(pix, ids, save=self.training and torch.is_grad_enabled())
            self._last_loss = out["loss"]
            return {"image_features": out["image_embeds"].clone(), "text_features": out["text_embeds"].clone(), "logit_scale": scale}, []
        self._last_loss = None
        out = eng.encode(pix, ids)
        return {"image_features": out["image_embeds"].clone() if out.get("image_embeds") is not None else None,
                "text_features": out["text_embeds"].clone() if out.get("text_embeds") is not None else None, "logit_scale": scale}, []

    def compute_loss(self, forward_outputs, label_ids=None, **kwargs):
        if isinstance(forward_outputs, tuple):       # the Trainer hands over what forward() returned
            forward_outputs = forward_outputs[0]
        if getattr(self, "_last_loss", None) is None:
            raise RuntimeError("compute_loss needs a forward() over both modalities")
        if self.training and torch.is_grad_enabled():
            return {"loss": _ClipLossFn.apply(self, next(iter(self._plist.values())), self._last_loss)}
        return {"loss": self._last_loss.detach().clone().view(())}
