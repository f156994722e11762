"""WordPiece tokenizer for the text side of the CLIP path -- restates the behaviour of the reference's BertTokenizer
(easynlp/modelzoo/models/bert/tokenization_bert.py:67-504: BasicTokenizer + WordpieceTokenizer, do_lower_case=True,
tokenize_chinese_chars=True) for the one call pattern the CLIP app uses (appzoo/clip/data.py:262-264,
appzoo/clip/predictor.py): `tokenizer([text], padding='max_length', truncation=True, max_length=L)`.
Pinned against the reference tokenizer by tests/golden/tokenizer.json."""
import collections
import unicodedata
from typing import Dict, List


def load_vocab(vocab_file: str) -> "collections.OrderedDict[str, int]":
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as reader:
        for index, token in enumerate(reader.readlines()):
            vocab[token.rstrip("\n")] = index
    return vocab


def _is_whitespace(ch):
    if ch in (" ", "\t", "\n", "\r"):
        return True
    return unicodedata.category(ch) == "Zs"


def _is_control(ch):
    if ch in ("\t", "\n", "\r"):
        return False
    return unicodedata.category(ch).startswith("C")


def _is_punctuation(ch):
    cp = ord(ch)
    if (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126):
        return True
    return unicodedata.category(ch).startswith("P")


def _is_chinese_char(cp):
    return ((0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x20000 <= cp <= 0x2A6DF) or (0x2A700 <= cp <= 0x2B73F)
            or (0x2B740 <= cp <= 0x2B81F) or (0x2B820 <= cp <= 0x2CEAF) or (0xF900 <= cp <= 0xFAFF) or (0x2F800 <= cp <= 0x2FA1F))


class BertTokenizer:
    def __init__(self, vocab_file, do_lower_case=True, unk_token="[UNK]", sep_token="[SEP]", pad_token="[PAD]", cls_token="[CLS]",
                 mask_token="[MASK]"):
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = {i: t for t, i in self.vocab.items()}
        self.do_lower_case = do_lower_case
        self.unk_token, self.sep_token, self.pad_token, self.cls_token, self.mask_token = unk_token, sep_token, pad_token, cls_token, mask_token
        self.never_split = {unk_token, sep_token, pad_token, cls_token, mask_token}
        self.max_input_chars_per_word = 100

    @classmethod
    def from_pretrained(cls, path, **kwargs):
        import os
        if os.path.isdir(path):
            path = os.path.join(path, "vocab.txt")
        return cls(path, **kwargs)

    # ---------------------------------------------------------------- basic tokenizer
    def _clean(self, text):
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            out.append(" " if _is_whitespace(ch) else ch)
        return "".join(out)

    def _basic(self, text) -> List[str]:
        text = self._clean(text)
        buf = []
        for ch in text:
            if _is_chinese_char(ord(ch)):
                buf.append(" "); buf.append(ch); buf.append(" ")
            else:
                buf.append(ch)
        tokens = []
        for tok in "".join(buf).strip().split():
            if tok in self.never_split:
                tokens.append(tok)
                continue
            if self.do_lower_case:
                tok = tok.lower()
                tok = "".join(c for c in unicodedata.normalize("NFD", tok) if unicodedata.category(c) != "Mn")
            cur = []
            for ch in tok:
                if _is_punctuation(ch):
                    if cur:
                        tokens.append("".join(cur)); cur = []
                    tokens.append(ch)
                else:
                    cur.append(ch)
            if cur:
                tokens.append("".join(cur))
        return " ".join(tokens).strip().split()

    def _wordpiece(self, token) -> List[str]:
        if len(token) > self.max_input_chars_per_word:
            return [self.unk_token]
        out, start = [], 0
        while start < len(token):
            end = len(token)
            cur = None
            while start < end:
                sub = token[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk_token]
            out.append(cur)
            start = end
        return out

    def tokenize(self, text) -> List[str]:
        out = []
        for tok in self._basic(text):
            if tok in self.never_split:
                out.append(tok)
            else:
                out.extend(self._wordpiece(tok))
        return out

    def convert_tokens_to_ids(self, tokens) -> List[int]:
        unk = self.vocab.get(self.unk_token, 0)
        return [self.vocab.get(t, unk) for t in tokens]

    # ---------------------------------------------------------------- the call the CLIP app makes
    def __call__(self, texts, padding="max_length", truncation=True, max_length=32, return_tensors="pt") -> Dict[str, "object"]:
        import torch
        if isinstance(texts, str):
            texts = [texts]
        cls_id, sep_id, pad_id = self.vocab[self.cls_token], self.vocab[self.sep_token], self.vocab.get(self.pad_token, 0)
        ids_all, mask_all = [], []
        for t in texts:
            ids = self.convert_tokens_to_ids(self.tokenize(t))
            if truncation and len(ids) > max_length - 2:
                ids = ids[: max_length - 2]
            ids = [cls_id] + ids + [sep_id]
            mask = [1] * len(ids)
            if padding == "max_length":
                pad = max_length - len(ids)
                ids = ids + [pad_id] * pad; mask = mask + [0] * pad
            ids_all.append(ids); mask_all.append(mask)
        if return_tensors == "pt":
            return {"input_ids": torch.tensor(ids_all, dtype=torch.long), "token_type_ids": torch.zeros(len(ids_all), len(ids_all[0]), dtype=torch.long),
                    "attention_mask": torch.tensor(mask_all, dtype=torch.long)}
        return {"input_ids": ids_all, "token_type_ids": [[0] * len(x) for x in ids_all], "attention_mask": mask_all}
