"""Text2VideoRetrieval ('clip4clip') -- drop-in for easynlp/appzoo/text2video_retrieval/model.py:38-120: an OPEN_CLIP checkpoint whose image
tower encodes the T frames of every video; the per-frame embeddings are l2-normalised, mean-pooled under `video_masks` and normalised
again (:82-88,98-104), then the same symmetric InfoNCE as the CLIP application.  A sibling application of the CLIP path
that reuses the open_clip engine; the frame pooling is clipk_frame_pool_fwd/bwd.  Inputs / outputs keep the reference's dict keys
(`pixel_values` [B,T,3,H,W], `video_masks` [B,T], `input_ids`; `logits_per_text`, `logits_per_video`, `video_embeds`, `text_embeds`)."""
import torch

from ..clip.model import CLIPApp


class Text2VideoRetrieval(CLIPApp):

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, user_defined_parameters={}, **kwargs):
        return cls(pretrained_model_name_or_path, user_defined_parameters)

    def __init__(self, pretrained_model_name_or_path=None, user_defined_parameters=None, **kwargs):
        super().__init__(pretrained_model_name_or_path, user_defined_parameters, **kwargs)
        if pretrained_model_name_or_path is not None and self.model_type != "open_clip":
            raise NotImplementedError("Text2VideoRetrieval loads open_clip checkpoints only (text2video_retrieval/model.py:54-61 of the reference)")

    def forward(self, inputs, feat=None):
        dev = self.engine.dev
        pix = inputs.get("pixel_values"); ids = inputs.get("input_ids")
        vm = inputs.get("video_masks")
        if pix is not None:
            inputs["pixel_values"] = pix = pix.to(dev, non_blocking=True).float()
            inputs["video_masks"] = vm = vm.to(dev, non_blocking=True)
            B, T = pix.shape[0], pix.shape[1]
            inputs["pixel_values"] = pix.view(B * T, *pix.shape[2:])          # the reference flattens the batch dict entry too (:73-74)
        if ids is not None:
            inputs["input_ids"] = ids = ids.to(dev, non_blocking=True).long().contiguous()
        eng = self.engine
        if feat is True:
            out = {"video_embeds": None, "text_embeds": None}
            if pix is not None:
                v = eng.vit_forward(pix.reshape(B * T, *pix.shape[2:]).contiguous(), save=False)
                out["video_embeds"] = eng._video_pool(v, vm, B, T)["embeds"].clone()
            if ids is not None:
                out["text_embeds"] = eng.bert_forward(ids, save=False)["embeds"].clone()
            return out
        assert pix is not None and ids is not None
        o = eng.forward(pix, ids, save=self.training and torch.is_grad_enabled(), video_masks=vm)
        lpt = o["logits_per_text"].clone()
        self._last_loss = o["loss"]
        return {"logits_per_text": lpt, "logits_per_video": lpt.T, "video_embeds": o["video_embeds"].clone(), "text_embeds": o["text_embeds"].clone()}
