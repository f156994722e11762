"""WukongCLIPPredictor -- drop-in for easynlp/appzoo/wukong_clip/predictor.py:32-139: rows with a text column (`first_sequence`) or a
base64 image column (`second_sequence`) -> one modality's features per row under 'text_feat' / 'image_feat' (tab-joined floats; a row
carrying both encodes only the text, as predict() overwrites).  `feature_format="numpy"` is the binary sink of the CLIP predictor."""
import os

import torch

from ...core.predictor import Predictor
from ..clip.data import decode_image, preprocess_image
from .data import FullTokenizer, wukong_tokenize
from .model import WukongCLIP


class WukongCLIPPredictor(Predictor):
    def __init__(self, model_dir, model_cls=None, user_defined_parameters=None, *args, **kwargs):
        super().__init__()
        self.tokenizer = FullTokenizer(vocab_file=os.path.join(model_dir, "vocab.txt"))
        self.model = (model_cls or WukongCLIP).from_pretrained(model_dir)
        self.model.eval()
        self.first_sequence = kwargs.pop("first_sequence", "first_sequence")
        self.second_sequence = kwargs.pop("second_sequence", "second_sequence")
        self.sequence_length = kwargs.pop("sequence_length", 128)
        self.feature_format = kwargs.pop("feature_format", "text")
        # decoded RGB images of a batch are resized / cropped / normalised in one GPU call (bit-identical to the host chain,
        # easynlp_b200/image_pipeline.py); gpu_preprocess=False keeps the per-image PIL + numpy chain of the reference
        self.gpu_preprocess = bool(kwargs.pop("gpu_preprocess", True))
        if self.feature_format not in ("text", "numpy"):
            raise ValueError(f"feature_format must be 'text' or 'numpy', got {self.feature_format!r}")

    def tokenize(self, texts, context_length: int = 32):
        return wukong_tokenize(self.tokenizer, texts, context_length)

    def preprocess(self, in_data):
        if not in_data:
            raise RuntimeError("Input data should not be None.")
        if not isinstance(in_data, list):
            in_data = [in_data]
        pending = []
        for record in in_data:
            text = record.get(self.first_sequence, None)
            image = record.get(self.second_sequence, None)
            if text is not None:
                record["input_ids"] = self.tokenize(text)
            if image is not None:
                img = decode_image(image)
                if self.gpu_preprocess and img.mode == "RGB":
                    pending.append((record, img))
                else:
                    record["pixel_values"] = preprocess_image(img)
        if pending:
            from ...image_pipeline import preprocess_images
            batch = preprocess_images([img for _, img in pending])
            for j, (record, _) in enumerate(pending):
                record["pixel_values"] = batch[j:j + 1]
            dev = batch.device
            for record in in_data:       # host-chain leftovers (palette / grey / alpha images) join the batch on the same device
                if record.get("pixel_values") is not None and record["pixel_values"].device != dev:
                    record["pixel_values"] = record["pixel_values"].to(dev)
        return in_data

    def predict(self, in_data):
        output = {}
        if "pixel_values" in in_data[0]:
            output = {"pixel_values": torch.cat([d["pixel_values"] for d in in_data], dim=0)}
        if "input_ids" in in_data[0]:
            output = {"input_ids": torch.cat([d["input_ids"] for d in in_data], dim=0)}
        with torch.no_grad():
            result, _ = self.model(output)
        return result

    def postprocess(self, result):
        for key, col in (("image_features", "image_feat"), ("text_features", "text_feat")):
            if result.get(key) is not None:
                embs = result[key].detach().float().cpu().numpy()
                if self.feature_format == "numpy":
                    return [{col: emb} for emb in embs]
                return [{col: "\t".join(str(x) for x in emb)} for emb in embs]
        return []
