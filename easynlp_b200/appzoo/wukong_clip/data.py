"""WukongCLIPDataset / FullTokenizer -- drop-in for easynlp/appzoo/wukong_clip/data.py:125-241 and bert_tokenizer.py:166-396 (host code,
like the reference).  Rows = (text column `first_sequence`, base64 image column `second_sequence`); the image goes through the CLIP
preprocessing (bicubic short side 224 -> centre crop -> /255 -> mean/std, data.py:32-124 = the CLIP application's) and the text through
`tokenize`: [CLS] + WordPiece ids[:context_length - 2] + [SEP], zero padded to context_length = 32 (data.py:166-187)."""
import os

import torch
from torch.utils.data import Dataset

from ...tokenization import BertTokenizer
from ..clip.data import collate_pixels, decode_image, parse_schema, preprocess_image


class FullTokenizer(BertTokenizer):
    """Google's BERT tokenizer as vendored by the reference (bert_tokenizer.py:166-396).  Differences from the BertTokenizer of the CLIP
    application: special tokens are NOT protected from the punctuation split (no never_split), words up to 200 characters are split
    into word pieces (:340), `tokenize` takes one string."""

    def __init__(self, vocab_file, do_lower_case=True):
        super().__init__(vocab_file, do_lower_case=do_lower_case)
        self.never_split = set()
        self.max_input_chars_per_word = 200
        self._native = False           # the native encoder implements BertTokenizer's rules (never_split, 100 characters)


def wukong_tokenize(tokenizer, texts, context_length: int = 32) -> torch.Tensor:
    if isinstance(texts, str):
        texts = [texts]
    cls_id, sep_id = tokenizer.vocab["[CLS]"], tokenizer.vocab["[SEP]"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        tokens = [cls_id] + tokenizer.convert_tokens_to_ids(tokenizer.tokenize(text))[:context_length - 2] + [sep_id]
        result[i, :len(tokens)] = torch.tensor(tokens)
    return result


class WukongCLIPDataset(Dataset):
    def __init__(self, pretrained_model_name_or_path, data_file, max_seq_length=32, input_schema=None, first_sequence=None, label_name=None,
                 second_sequence=None, label_enumerate_values=None, user_defined_parameters=None, skip_first_line=False, *args, **kwargs):
        self.columns = parse_schema(input_schema)
        self.text_col = first_sequence
        self.image_col = second_sequence
        self.label_enumerate_values = label_enumerate_values
        with open(data_file, "r", encoding="utf-8") as f:
            lines = f.read().splitlines()
        if skip_first_line:
            lines = lines[1:]
        self.data_rows = [ln for ln in lines if ln]
        self.tokenizer = FullTokenizer(vocab_file=os.path.join(pretrained_model_name_or_path, "vocab.txt"))
        self.max_text_length = max_seq_length
        ap = (user_defined_parameters or {}).get("app_parameters", {}) if isinstance(user_defined_parameters, dict) else {}
        self.gpu_preprocess = str(kwargs.get("gpu_preprocess", ap.get("gpu_preprocess", False))).lower() in ("1", "true", "yes")     # see CLIPDataset

    def __len__(self):
        return len(self.data_rows)

    def __getitem__(self, item):
        fields = self.data_rows[item].split("\t")
        return self.convert_single_row_to_example({c: v for c, v in zip(self.columns, fields)})

    def tokenize(self, texts, context_length: int = 32):
        return wukong_tokenize(self.tokenizer, texts, context_length)

    def convert_single_row_to_example(self, row):
        # the reference tokenises with the default context length 32 whatever max_seq_length says (data.py:209)
        tk = {"input_ids": self.tokenize(row[self.text_col])}
        image = decode_image(row[self.image_col])
        if self.gpu_preprocess and image.mode == "RGB":
            image.load()
            return {"text": tk, "image": image}
        return {"text": tk, "pixel_values": preprocess_image(image)}

    def batch_fn(self, features):
        return {"pixel_values": collate_pixels(features),
                "input_ids": torch.cat([f["text"]["input_ids"] for f in features], dim=0)}
