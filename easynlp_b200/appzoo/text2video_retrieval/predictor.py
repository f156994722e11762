"""Text2VideoRetrievalPredictor -- drop-in for easynlp/appzoo/text2video_retrieval/predictor.py:32-143.  As in the reference the modality is
chosen by the NAME of the first_sequence column: 'text' rows are BPE-tokenised (:82-85), 'image' rows name a directory of frames (:87-104);
output rows carry 'text_feat' / 'video_feat' (tab-joined floats, or float32 arrays with feature_format='numpy')."""
import json
import os

import torch

from ...bpe_tokenizer import SimpleTokenizer, openclip_tokenize
from ...core.predictor import Predictor
from .data import MAX_FRAMES, frames_to_pixels, load_frames, video_mask
from .model import Text2VideoRetrieval


class Text2VideoRetrievalPredictor(Predictor):
    def __init__(self, model_dir, model_cls=None, user_defined_parameters=None, *args, **kwargs):
        super().__init__()
        with open(os.path.join(model_dir, "config.json"), "r") as f:
            self.raw_config = json.load(f)
        if self.raw_config.get("model_type") != "open_clip":
            raise NotImplementedError("Text2VideoRetrievalPredictor supports open_clip checkpoints only")
        self.model_type = "open_clip"
        self.openclip_tokenizer = SimpleTokenizer(bpe_path=os.path.join(model_dir, "vocab.txt"))
        self.multi_modal = (model_cls or Text2VideoRetrieval).from_pretrained(model_dir)
        self.multi_modal.eval()
        self.first_sequence = kwargs.pop("first_sequence", "first_sequence")
        self.second_sequence = kwargs.pop("second_sequence", "second_sequence")
        self.sequence_length = kwargs.pop("sequence_length", 128)
        self.feature_format = kwargs.pop("feature_format", "text")
        if self.feature_format not in ("text", "numpy"):
            raise ValueError(f"feature_format must be 'text' or 'numpy', got {self.feature_format!r}")
        self.gpu_preprocess = bool(kwargs.pop("gpu_preprocess", True))
        self.max_frames = MAX_FRAMES

    def preprocess(self, in_data):
        if not in_data:
            raise RuntimeError("Input data should not be None.")
        if not isinstance(in_data, list):
            in_data = [in_data]
        for record in in_data:
            content = record.get(self.first_sequence, None)
            if self.first_sequence == "text":
                record["input_ids"] = openclip_tokenize(texts=[content], context_length=77, _tokenizer=self.openclip_tokenizer)
            elif self.first_sequence == "image":
                images, n = load_frames(content, self.max_frames)
                record["pixel_values"] = frames_to_pixels(images, self.gpu_preprocess)
                record["video_masks"] = video_mask(n, len(images))
        return in_data

    def predict(self, in_data):
        output = {}
        if "pixel_values" in in_data[0]:
            output = {"pixel_values": torch.cat([d["pixel_values"] for d in in_data], dim=0),
                      "video_masks": torch.cat([d["video_masks"] for d in in_data], dim=0)}
        if "input_ids" in in_data[0]:
            output = {"input_ids": torch.cat([d["input_ids"] for d in in_data], dim=0)}
        with torch.no_grad():
            return self.multi_modal(output, feat=True)

    def postprocess(self, result):
        for key, col in (("video_embeds", "video_feat"), ("text_embeds", "text_feat")):
            if result.get(key) is not None:
                embs = result[key].detach().float().cpu().numpy()
                if self.feature_format == "numpy":
                    return [{col: emb} for emb in embs]
                return [{col: "\t".join(str(x) for x in emb)} for emb in embs]
        return []
