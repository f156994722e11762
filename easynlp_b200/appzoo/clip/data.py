"""CLIPDataset + image preprocessing -- the data format on the input side of the hot path
(easynlp/appzoo/clip/data.py:29-135 transforms, :165-295 dataset; easynlp/appzoo/dataset.py BaseDataset TSV reader).

Row format (TSV, `input_schema` like "text:str:1,image:str:1"): text column + base64 (urlsafe) encoded image column.
Each example -> {'text': tokenizer output, 'pixel_values': [1,3,224,224] fp32}; batch_fn concatenates and adds the
empty 'label_ids' the Trainer pops (data.py:275-295).  Host-side code (PIL / numpy), like the reference."""
import base64
import json
import os
from io import BytesIO

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from ...tokenization import BertTokenizer
from ...bpe_tokenizer import SimpleTokenizer, openclip_tokenize

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def _resize(image, size=224, resample=Image.BICUBIC):
    """shorter side -> `size`, aspect preserved (data.py:54-74)."""
    if isinstance(size, tuple):
        new_w, new_h = size
    else:
        width, height = image.size
        short, long = (width, height) if width <= height else (height, width)
        if short == size:
            return image
        new_short, new_long = size, int(size * long / short)
        new_w, new_h = (new_short, new_long) if width <= height else (new_long, new_short)
    return image.resize((new_w, new_h), resample)


def _center_crop(image, size=224):
    """data.py:29-52 (pads when the image is smaller than the crop)."""
    if not isinstance(size, tuple):
        size = (size, size)
    w, h = image.size
    ch, cw = size
    top = int((h - ch + 1) * 0.5)
    left = int((w - cw + 1) * 0.5)
    return image.crop((left, top, left + cw, top + ch))


def _normalize(image, mean=CLIP_MEAN, std=CLIP_STD):
    """RGB -> [3,H,W] float32 in [0,1] -> (x - mean) / std (data.py:76-135)."""
    arr = np.array(image.convert("RGB")).astype(np.float32) / 255.0
    arr = arr.transpose(2, 0, 1)
    return (arr - mean[:, None, None]) / std[:, None, None]


def preprocess_image(image) -> torch.Tensor:
    x = _normalize(_center_crop(_resize(image, 224, Image.BICUBIC), 224))
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).unsqueeze(0)


def decode_image(b64: str):
    return Image.open(BytesIO(base64.urlsafe_b64decode(b64)))


def parse_schema(input_schema):
    """'text:str:1,image:str:1' -> ['text', 'image']"""
    return [c.split(":")[0] for c in input_schema.split(",")] if input_schema else []


def collate_pixels(features) -> torch.Tensor:
    """[n, 3, 224, 224] of a batch whose examples carry either host-preprocessed `pixel_values` or a decoded RGB `image` (gpu_preprocess)"""
    todo = [i for i, f in enumerate(features) if "image" in f]
    if not todo:
        return torch.cat([f["pixel_values"] for f in features], dim=0)
    from ...image_pipeline import preprocess_images
    batch = preprocess_images([features[i]["image"] for i in todo])
    if len(todo) == len(features):
        return batch
    rows = [None] * len(features)
    for j, i in enumerate(todo):
        rows[i] = batch[j:j + 1]
    for i, f in enumerate(features):
        if rows[i] is None:
            rows[i] = f["pixel_values"].to(batch.device)
    return torch.cat(rows, dim=0)


class CLIPDataset(Dataset):
    def __init__(self, pretrained_model_name_or_path, data_file, max_seq_length, input_schema=None, first_sequence=None, label_name=None,
                 second_sequence=None, label_enumerate_values=None, user_defined_parameters=None, skip_first_line=False, *args, **kwargs):
        with open(os.path.join(pretrained_model_name_or_path, "config.json"), "r") as f:
            self.raw_config = json.load(f)
        mt = self.raw_config.get("model_type")
        # open_clip tokenises with the byte-level BPE SimpleTokenizer, the other two branches with BertTokenizer (data.py:225-229)
        self.model_type = mt if mt in ("open_clip", "chinese_clip") else "huggingface_clip"
        self.columns = parse_schema(input_schema)
        self.text_col = first_sequence
        self.image_col = second_sequence
        self.label_enumerate_values = label_enumerate_values
        with open(data_file, "r", encoding="utf-8") as f:
            lines = f.read().splitlines()
        if skip_first_line:
            lines = lines[1:]
        self.data_rows = [ln for ln in lines if ln]
        vocab = os.path.join(pretrained_model_name_or_path, "vocab.txt")
        if self.model_type == "open_clip":
            self.openclip_tokenizer = SimpleTokenizer(bpe_path=vocab)       # `vocab.txt` is the gzip'd merges file here (data.py:226)
        else:
            self.tokenizer = BertTokenizer.from_pretrained(vocab)
        self.max_text_length = max_seq_length
        # gpu_preprocess (kwarg or user_defined_parameters['app_parameters']): examples carry the decoded RGB image and batch_fn runs the
        # resize / crop / normalise of the whole batch in one GPU call (easynlp_b200/image_pipeline.py, bit-identical to the host chain).
        # Off by default: batch_fn must then run in the main process (data loader workers = 0).
        ap = (user_defined_parameters or {}).get("app_parameters", {}) if isinstance(user_defined_parameters, dict) else {}
        self.gpu_preprocess = str(kwargs.get("gpu_preprocess", ap.get("gpu_preprocess", False))).lower() in ("1", "true", "yes")

    def __len__(self):
        return len(self.data_rows)

    def __getitem__(self, item):
        fields = self.data_rows[item].split("\t")
        row = {c: v for c, v in zip(self.columns, fields)}
        return self.convert_single_row_to_example(row)

    def convert_single_row_to_example(self, row):
        if self.model_type == "open_clip":      # fixed context length 77 whatever max_seq_length says (data.py:259-261)
            tk = {"input_ids": openclip_tokenize(texts=[row[self.text_col]], context_length=77, _tokenizer=self.openclip_tokenizer)}
        else:
            tk = self.tokenizer([row[self.text_col]], padding="max_length", truncation=True, max_length=self.max_text_length, return_tensors="pt")
        image = decode_image(row[self.image_col])
        if self.gpu_preprocess and image.mode == "RGB":
            image.load()
            return {"text": tk, "image": image}
        return {"text": tk, "pixel_values": preprocess_image(image)}

    def batch_fn(self, features):
        out = {"pixel_values": collate_pixels(features),
               "input_ids": torch.cat([f["text"]["input_ids"] for f in features], dim=0)}
        for k in ("token_type_ids", "attention_mask"):      # absent from BPE rows: the reference then leaves an empty list (data.py:276-294)
            out[k] = torch.cat([f["text"][k] for f in features], dim=0) if all(k in f["text"] for f in features) else []
        out["label_ids"] = []
        return out
