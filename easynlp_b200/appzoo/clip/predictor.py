"""CLIPPredictor -- drop-in for easynlp/appzoo/clip/predictor.py:32-153: rows with a text column (`first_sequence`) or a
base64 image column (`second_sequence`) -> one modality's embedding per row, formatted as tab-joined floats under
'text_feat' / 'image_feat'.  As in the reference, a row carrying both encodes only the text (predict() overwrites).

`feature_format` (new): "text" (default) is the reference's `'\t'.join(str(x))` (predictor.py:140-153, ~10 bytes and a
float->str conversion per value: GBs of text at 1 M vectors); "numpy" hands back float32 arrays under the same keys, which
SimplePredictorManager writes as ONE [rows, E] .npy file when the output file ends in .npy."""
import json
import os

import torch

from ...core.predictor import Predictor
from ...tokenization import BertTokenizer
from ...bpe_tokenizer import SimpleTokenizer, openclip_tokenize
from .data import decode_image, preprocess_image


class CLIPPredictor(Predictor):
    def __init__(self, model_dir, model_cls=None, user_defined_parameters=None, *args, **kwargs):
        super().__init__()
        with open(os.path.join(model_dir, "config.json"), "r") as f:
            self.raw_config = json.load(f)
        mt = self.raw_config.get("model_type")
        self.model_type = mt if mt in ("open_clip", "chinese_clip") else "huggingface_clip"
        if self.model_type == "open_clip":      # byte-level BPE over the gzip'd merges file `vocab.txt` (predictor.py:54-55)
            self.openclip_tokenizer = SimpleTokenizer(bpe_path=os.path.join(model_dir, "vocab.txt"))
        else:
            self.tokenizer = BertTokenizer.from_pretrained(os.path.join(model_dir, "vocab.txt"))
        if model_cls is None:
            from .model import CLIPApp as model_cls
        self.multi_modal = model_cls.from_pretrained(model_dir, user_defined_parameters=user_defined_parameters or {}).cuda()
        self.multi_modal.eval()
        self.first_sequence = kwargs.pop("first_sequence", "first_sequence")
        self.second_sequence = kwargs.pop("second_sequence", "second_sequence")
        self.sequence_length = kwargs.pop("sequence_length", 128)
        self.feature_format = kwargs.pop("feature_format", "text")
        # decoded RGB images of a batch are resized / cropped / normalised in one GPU call (bit-identical to the host chain,
        # easynlp_b200/image_pipeline.py); gpu_preprocess=False keeps the per-image PIL + numpy chain of the reference
        self.gpu_preprocess = bool(kwargs.pop("gpu_preprocess", True))
        if self.feature_format not in ("text", "numpy"):
            raise ValueError(f"feature_format must be 'text' or 'numpy', got {self.feature_format!r}")

    def preprocess(self, in_data):
        if not in_data:
            raise RuntimeError("Input data should not be None.")
        if not isinstance(in_data, list):
            in_data = [in_data]
        max_seq_length = -1
        for record in in_data:
            if "sequence_length" not in record:
                break
            max_seq_length = max(max_seq_length, record["sequence_length"])
        max_seq_length = self.sequence_length if max_seq_length == -1 else max_seq_length
        pending = []
        for record in in_data:
            text = record.get(self.first_sequence, None)
            image = record.get(self.second_sequence, None)
            if text is not None and self.model_type == "open_clip":
                record["input_ids"] = openclip_tokenize(texts=[text], context_length=77, _tokenizer=self.openclip_tokenizer)
            elif text is not None:
                tk = self.tokenizer(text, padding="max_length", truncation=True, max_length=max_seq_length, return_tensors="pt")
                record["input_ids"] = tk["input_ids"]; record["token_type_ids"] = tk["token_type_ids"]; record["attention_mask"] = tk["attention_mask"]
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
            if "token_type_ids" in in_data[0]:
                output["token_type_ids"] = torch.cat([d["token_type_ids"] for d in in_data], dim=0)
                output["attention_mask"] = torch.cat([d["attention_mask"] for d in in_data], dim=0)
        with torch.no_grad():
            return self.multi_modal(output, feat=True)

    def postprocess(self, result):
        fmt = getattr(self, "feature_format", "text")
        for key, col in (("image_embeds", "image_feat"), ("text_embeds", "text_feat")):
            if result.get(key) is not None:
                embs = result[key].detach().float().cpu().numpy()
                if fmt == "numpy":
                    return [{col: emb} for emb in embs]
                return [{col: "\t".join(str(x) for x in emb)} for emb in embs]
        return []
