"""CLIPApp -- drop-in for easynlp/appzoo/clip/model.py:40-164 whose math runs in the clipk sm_100a kernels.  All three branches of the
reference are implemented: model_type == "chinese_clip" (ViT + BertModel, model.py:64-72), "open_clip" (ViT + causal text transformer
with EOT pooling, model.py:56-63) and the default "huggingface_clip" branch (CLIPVisionModel, frozen by `.detach()`, + RobertaModel
with its tanh pooler and biased projections, model.py:73-104,128-144); ModifiedResNet visual towers raise.  Same constructor /
from_pretrained / forward(inputs, feat) / compute_loss contract, same `.config` wrapper, same
state_dict key names (`chinese_clip.` prefix resp. `text_encoder.` / `vision_encoder.` / `*_projection.`), so Trainer /
CLIPEvaluator / CLIPPredictor written against the reference keep working.  There is no CPU or PyTorch fallback."""
import json
import os
from collections import OrderedDict

import torch

from ..application import Application
from ...engine import ClipEngine, hf_engine_config

PREFIX = "chinese_clip."


class Config_Wrapper:
    """model.py:32-38 of the reference: the Trainer needs `.to_json_string()` and a `__dict__`."""

    def __init__(self, json_data):
        self.json_data = json_data

    def to_json_string(self):
        return json.dumps(self.json_data, ensure_ascii=False)


class _ClipLossFn(torch.autograd.Function):
    """Bridges torch.autograd (`loss.backward()` in a Trainer) to the engine's hand-written backward."""

    @staticmethod
    def forward(ctx, app, anchor, loss_value):
        ctx.app = app
        return loss_value.detach().clone().view(())

    @staticmethod
    def backward(ctx, grad_out):
        app = ctx.app
        app.engine.backward(grad_scale=float(grad_out))
        app._publish_grads()
        return None, None, None


class CLIPApp(Application):

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, user_defined_parameters={}, **kwargs):
        return cls(pretrained_model_name_or_path, user_defined_parameters)

    def __init__(self, pretrained_model_name_or_path=None, user_defined_parameters=None, **kwargs):
        super().__init__()
        self.engine = None
        if pretrained_model_name_or_path is None:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("easynlp_b200.CLIPApp needs a CUDA device (B200): there is no CPU fallback")
        path = pretrained_model_name_or_path
        with open(os.path.join(path, "config.json"), "r") as f:
            self.raw_config = json.load(f)
        mt = self.raw_config.get("model_type")
        ckpt = os.path.join(path, "pytorch_model.bin")
        self.config = Config_Wrapper(self.raw_config)
        checkpoint = torch.load(ckpt, map_location="cpu")
        if mt == "open_clip":
            # OPEN_CLIP(**config) with a ViT visual tower and the causal text transformer (model.py:56-63, modeling_openclip.py:255-383)
            self.model_type = "open_clip"
            self.prefix = "open_clip."
            cfg = {k: v for k, v in self.raw_config.items()}
        elif mt == "chinese_clip":
            self.model_type = "chinese_clip"
            self.prefix = PREFIX
            cfg = {k: v for k, v in self.raw_config.items()}
        else:
            # every other config takes the reference's huggingface_clip branch (model.py:73-104): nested text_config / vision_config,
            # checkpoint keys used as they are, projection width read off the checkpoint
            self.model_type = "huggingface_clip"
            self.prefix = ""
            if "text_projection.weight" not in checkpoint:
                raise KeyError("huggingface_clip checkpoint without text_projection.weight")
            cfg = hf_engine_config(self.raw_config, checkpoint["text_projection.weight"].shape[0])
        self.engine = ClipEngine(cfg, device=kwargs.get("device", "cuda"))
        self.engine.params.load_state_dict({(k[len(self.prefix):] if self.prefix and k.startswith(self.prefix) else k): v for k, v in checkpoint.items()},
                                           strict=False)
        self._wrap_params()
        self.distributed_loss = bool((user_defined_parameters or {}).get("app_parameters", {}).get("global_contrastive", False)) \
            if isinstance(user_defined_parameters, dict) else False

    # ------------------------------------------------------------------ parameters (reference names, shared storage)
    def _wrap_params(self):
        P = self.engine.params
        pre = getattr(self, "prefix", PREFIX)
        self.prefix = pre
        self._plist = OrderedDict()
        for n in P.names():
            self._plist[pre + n] = torch.nn.Parameter(P.p(n), requires_grad=True)

    def _publish_grads(self):
        P = self.engine.params
        for n in P.trainable_names():
            self._plist[self.prefix + n].grad = P.g(n)

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        for n, p in self._plist.items():
            yield (prefix + ("." if prefix else "") + n, p)

    def parameters(self, recurse=True):
        for _, p in self.named_parameters():
            yield p

    def state_dict(self, *args, **kwargs):
        sd = self.engine.params.state_dict()
        return OrderedDict((self.prefix + k, v) for k, v in sd.items())

    def load_state_dict(self, state_dict, strict=True):
        pre = self.prefix
        self.engine.params.load_state_dict({(k[len(pre):] if pre and k.startswith(pre) else k): v for k, v in state_dict.items()}, strict=False)

    def zero_grad(self, set_to_none=False):
        self.engine.zero_grad()

    def to(self, *args, **kwargs):      # parameters live on the GPU from construction
        return self

    def cuda(self, device=None):
        return self

    # ------------------------------------------------------------------ forward / loss
    def forward(self, inputs, feat=None):
        dev = self.engine.dev
        # like the reference (model.py:116-123) the batch dict is updated in place with device tensors
        inputs["pixel_values"] = inputs["pixel_values"].to(dev, non_blocking=True) if inputs.get("pixel_values") is not None else None
        inputs["input_ids"] = inputs["input_ids"].to(dev, non_blocking=True) if inputs.get("input_ids") is not None else None
        pix = inputs["pixel_values"]; ids = inputs["input_ids"]
        if pix is not None:
            pix = pix.float().contiguous()
        if ids is not None:
            ids = ids.long().contiguous()
        # the huggingface_clip branch feeds the batch's token_type_ids / attention_mask to the text tower (model.py:132-134); chinese_clip
        # ignores them and masks with ids != 0 (quirk A.4-3)
        tt = am = None
        if self.model_type == "huggingface_clip" and ids is not None:
            tt = inputs.get("token_type_ids"); am = inputs.get("attention_mask")
            tt = tt if torch.is_tensor(tt) else None; am = am if torch.is_tensor(am) else None
        if feat is True:    # outputs are copies: the engine reuses its buffers on the next call
            return {k: (v.clone() if v is not None else None) for k, v in self.engine.encode(pix, ids, token_type_ids=tt, attention_mask=am).items()}
        assert pix is not None and ids is not None, "text and image cannot both be None!"
        out = self.engine.forward(pix, ids, save=self.training and torch.is_grad_enabled(), distributed=self.distributed_loss,
                                  token_type_ids=tt, attention_mask=am)
        lpt = out["logits_per_text"].clone()
        self._last_loss = out["loss"]
        # world size 1 (and local-loss data parallelism): logits_per_image is the transpose view, as in the reference (model.py:149).
        # Global-batch mode: the rank holds two [b, G] strips -- its texts against the gathered images and its images against the
        # gathered texts; the second one IS its rows of the image->text logits (not a transpose of the first).
        lpi = out["logits_per_image"].clone() if out.get("distributed") else lpt.T
        return {"logits_per_text": lpt, "logits_per_image": lpi, "image_embeds": out["image_embeds"].clone(),
                "text_embeds": out["text_embeds"].clone()}

    def compute_loss(self, forward_outputs, label_ids, **kwargs):
        """(CE(S) + CE(S^T)) / 2 -- already evaluated by the fused CE-strip kernels during forward()."""
        anchor = next(iter(self._plist.values()))
        if self.training and torch.is_grad_enabled():
            return {"loss": _ClipLossFn.apply(self, anchor, self._last_loss)}
        return {"loss": self._last_loss.detach().clone().view(())}
