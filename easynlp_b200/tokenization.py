"""WordPiece tokenizer for the text side of the CLIP path -- restates the behaviour of the reference's BertTokenizer
(easynlp/modelzoo/models/bert/tokenization_bert.py:67-504: BasicTokenizer + WordpieceTokenizer, do_lower_case=True,
tokenize_chinese_chars=True) for the one call pattern the CLIP app uses (appzoo/clip/data.py:262-264,
appzoo/clip/predictor.py): `tokenizer([text], padding='max_length', truncation=True, max_length=L)`.
Pinned against the reference tokenizer by tests/golden/tokenizer.json.

Two implementations with identical results: this Python restatement and the NATIVE one behind the C ABI (clipk_wp_*,
easynlp_b200/csrc/wordpiece.cu: multi-threaded batches, per-code-point tables generated from the same Python unicodedata).  `__call__`
uses the native encoder when the library is present; a text holding a code point outside the native tables (supplementary planes other
than the CJK extensions, capital sigma) is tokenized here."""
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
        self.vocab_file = vocab_file
        self._native = None          # handle of the native encoder (created lazily; False = unavailable)
        self.native_threads = 8
        self.native_stats = {"native": 0, "fallback": 0}

    def _native_handle(self):
        if self._native is None:
            self._native = False
            defaults = (self.unk_token, self.sep_token, self.pad_token, self.cls_token, self.mask_token) == ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")
            try:
                import os
                from . import _lib as L
                table = os.path.join(os.path.dirname(L.LIB_PATH), "unicode_bmp.bin")
                if defaults and os.path.exists(L.LIB_PATH) and os.path.exists(table):
                    h = L.lib().clipk_wp_create(self.vocab_file.encode(), table.encode(), int(self.do_lower_case))
                    if h:
                        self._native = h
            except Exception:
                self._native = False
        return self._native

    def encode_native(self, texts, max_length):
        """(ids [n, L] int64, mask [n, L] int64, status [n]) from the native encoder, or None when it is unavailable"""
        h = self._native_handle()
        if not h:
            return None
        import ctypes as C
        import numpy as np
        from . import _lib as L
        n = len(texts)
        ids = np.zeros((n, max_length), dtype=np.int64); mask = np.zeros((n, max_length), dtype=np.int64); status = np.zeros(n, dtype=np.int32)
        arr = (C.c_char_p * n)(*[t.encode("utf-8", "surrogatepass") if isinstance(t, str) else t for t in texts])
        rc = L.lib().clipk_wp_encode_batch(C.c_void_p(h), arr, n, int(max_length), ids.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p),
                                           status.ctypes.data_as(C.c_void_p), int(self.native_threads))
        if rc != 0:
            return None
        return ids, mask, status

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
        if padding == "max_length" and truncation and max_length >= 2 and all(isinstance(t, str) and "\x00" not in t for t in texts):
            nat = self.encode_native(texts, max_length)
            if nat is not None:
                ids_np, mask_np, status = nat
                for i, t in enumerate(texts):
                    if status[i] < 0:                      # code point outside the native tables: the Python restatement takes this text
                        one = self._encode_python([t], padding, truncation, max_length, cls_id, sep_id, pad_id)
                        ids_np[i] = one[0][0]; mask_np[i] = one[1][0]
                        self.native_stats["fallback"] += 1
                    else:
                        self.native_stats["native"] += 1
                if return_tensors == "pt":
                    return {"input_ids": torch.from_numpy(ids_np.copy()), "token_type_ids": torch.zeros(len(texts), max_length, dtype=torch.long),
                            "attention_mask": torch.from_numpy(mask_np.copy())}
                return {"input_ids": ids_np.tolist(), "token_type_ids": [[0] * max_length for _ in texts], "attention_mask": mask_np.tolist()}
        ids_all, mask_all = self._encode_python(texts, padding, truncation, max_length, cls_id, sep_id, pad_id)
        if return_tensors == "pt":
            return {"input_ids": torch.tensor(ids_all, dtype=torch.long), "token_type_ids": torch.zeros(len(ids_all), len(ids_all[0]), dtype=torch.long),
                    "attention_mask": torch.tensor(mask_all, dtype=torch.long)}
        return {"input_ids": ids_all, "token_type_ids": [[0] * len(x) for x in ids_all], "attention_mask": mask_all}

    def _encode_python(self, texts, padding, truncation, max_length, cls_id, sep_id, pad_id):
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
        return ids_all, mask_all
