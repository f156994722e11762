"""Text2VideoRetrievalDataset -- drop-in for easynlp/appzoo/text2video_retrieval/data.py:164-279 (host code + the batched GPU image chain).

Row format (TSV): a text column (`first_sequence`) and a column holding a DIRECTORY of extracted frames (`second_sequence`).  Each example:
the frames of the directory, padded with black 224 x 224 frames up to max_frames = 12 (:236-240), each through the CLIP image chain
(bicubic short side 224 -> centre crop -> normalise) -> `pixel_values` [1, T, 3, 224, 224], `video_masks` [1, T] (1 for real frames),
`text` = byte-level BPE ids padded to 77 (open_clip checkpoints only, as in the reference).  Frames are read in sorted name order
(the reference iterates os.listdir order; the masked mean over frames does not depend on it)."""
import json
import os

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from ...bpe_tokenizer import SimpleTokenizer, openclip_tokenize
from ..clip.data import parse_schema, preprocess_image

MAX_FRAMES = 12


def load_frames(frame_dir: str, max_frames: int = MAX_FRAMES, size: int = 224):
    """-> (list of PIL images padded with black frames to max_frames, number of real frames)"""
    images = [Image.open(os.path.join(frame_dir, f)) for f in sorted(os.listdir(frame_dir))]
    n = len(images)
    images += [Image.new("RGB", (size, size), (0, 0, 0)) for _ in range(max_frames - n)]
    return images, n


def frames_to_pixels(images, gpu: bool) -> torch.Tensor:
    """[1, T, 3, 224, 224]; gpu=True: one clipk_preprocess_images call for the video's RGB frames (bit-identical to the host chain)"""
    if gpu and all(im.mode == "RGB" for im in images):
        from ...image_pipeline import preprocess_images
        for im in images:
            im.load()
        return preprocess_images(images).unsqueeze(0)
    return torch.cat([preprocess_image(im) for im in images], dim=0).unsqueeze(0)


def video_mask(n_real: int, n_total: int) -> torch.Tensor:
    m = np.zeros((1, n_total), dtype=np.int64)
    m[0, :n_real] = 1
    return torch.from_numpy(m)


class Text2VideoRetrievalDataset(Dataset):
    def __init__(self, pretrained_model_name_or_path, data_file, max_seq_length, input_schema=None, first_sequence=None, label_name=None,
                 second_sequence=None, label_enumerate_values=None, user_defined_parameters=None, skip_first_line=False, *args, **kwargs):
        with open(os.path.join(pretrained_model_name_or_path, "config.json"), "r") as f:
            self.raw_config = json.load(f)
        if self.raw_config.get("model_type") != "open_clip":
            raise NotImplementedError("Text2VideoRetrievalDataset supports open_clip checkpoints only (text2video_retrieval/data.py:196-210 of the reference)")
        self.model_type = "open_clip"
        self.columns = parse_schema(input_schema)
        self.text_col = first_sequence
        self.image_col = second_sequence
        self.label_enumerate_values = label_enumerate_values
        with open(data_file, "r", encoding="utf-8") as f:
            lines = f.read().splitlines()
        if skip_first_line:
            lines = lines[1:]
        self.data_rows = [ln for ln in lines if ln]
        self.openclip_tokenizer = SimpleTokenizer(bpe_path=os.path.join(pretrained_model_name_or_path, "vocab.txt"))
        self.max_text_length = max_seq_length
        self.max_frames = MAX_FRAMES
        ap = (user_defined_parameters or {}).get("app_parameters", {}) if isinstance(user_defined_parameters, dict) else {}
        self.gpu_preprocess = str(kwargs.get("gpu_preprocess", ap.get("gpu_preprocess", False))).lower() in ("1", "true", "yes")     # see CLIPDataset

    def __len__(self):
        return len(self.data_rows)

    def __getitem__(self, item):
        fields = self.data_rows[item].split("\t")
        return self.convert_single_row_to_example({c: v for c, v in zip(self.columns, fields)})

    def convert_single_row_to_example(self, row):
        images, n = load_frames(row[self.image_col], self.max_frames)
        tk = {"input_ids": openclip_tokenize(texts=[row[self.text_col]], context_length=77, _tokenizer=self.openclip_tokenizer)}
        return {"text": tk, "pixel_values": frames_to_pixels(images, self.gpu_preprocess), "video_masks": video_mask(n, len(images))}

    def batch_fn(self, features):
        dev = next((f["pixel_values"].device for f in features if f["pixel_values"].is_cuda), None)
        return {"pixel_values": torch.cat([f["pixel_values"].to(dev) if dev is not None else f["pixel_values"] for f in features], dim=0),
                "video_masks": torch.cat([f["video_masks"] for f in features], dim=0),
                "input_ids": torch.cat([f["text"]["input_ids"] for f in features], dim=0),
                "token_type_ids": [], "attention_mask": [], "label_ids": []}
