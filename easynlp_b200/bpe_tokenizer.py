"""Byte-level BPE tokenizer of the open_clip branch -- host code, like the reference's (easynlp/modelzoo/models/clip/openclip_tokenizer.py:
27-157, used by CLIPDataset / CLIPPredictor when model_type == open_clip, appzoo/clip/data.py:137-163,225-226,259-261).

Same vocabulary construction (256 byte symbols, their `</w>` forms, one entry per merge rule in file order, then `<start_of_text>` /
`<end_of_text>` and any extra special tokens), same pre-tokenisation pattern, same lowest-rank-first merging, same id layout, so ids
are interchangeable with the reference's for the same `vocab.txt` (a gzip'd merges file).

Deviation: the reference cleans text with `ftfy.fix_text` first (:59-62).  ftfy is used when it is importable; where it is not (this
image) the clean-up is Unicode NFC normalisation -- ftfy's final step and the only one that touches well-formed text.  The golden
vectors (oracle/make_golden_bpe.py) were produced by the unmodified reference class with ftfy stubbed the same way."""
import gzip
import html
import unicodedata
from typing import Dict, List, Tuple, Union

import regex as re

try:                                    # pragma: no cover - not present in this image
    from ftfy import fix_text as _fix_text
except Exception:                       # noqa: BLE001
    def _fix_text(text: str) -> str:
        return unicodedata.normalize("NFC", text)

N_MERGES = 49152 - 256 - 2              # merge rules read from the file (openclip_tokenizer.py:75)


def byte_symbols() -> Dict[int, str]:
    """byte -> printable stand-in character: printable Latin-1 bytes map to themselves, the 68 others to U+0100.. in byte order
    (openclip_tokenizer.py:27-47)"""
    keep = list(range(0x21, 0x7F)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # the reference's vocabulary order is: kept bytes in the order above, then the remapped ones in byte order
    return {b: table[b] for b in keep + [b for b in range(256) if b not in keep]}


class SimpleTokenizer:
    def __init__(self, bpe_path: str, special_tokens=None):
        self.byte_encoder = byte_symbols()
        self.byte_decoder = {c: b for b, c in self.byte_encoder.items()}
        with gzip.open(bpe_path) as f:
            lines = f.read().decode("utf-8").split("\n")
        merges: List[Tuple[str, ...]] = [tuple(ln.split()) for ln in lines[1:N_MERGES + 1]]       # line 0 is the file's header
        symbols = list(self.byte_encoder.values())
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges]
        specials = ["<start_of_text>", "<end_of_text>"] + list(special_tokens or [])
        vocab += specials
        self.encoder = {tok: i for i, tok in enumerate(vocab)}        # later duplicates win, as in dict(zip(...))
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.bpe_ranks = {m: r for r, m in enumerate(merges)}
        self.cache = {t: t for t in specials}
        self.pat = re.compile("|".join(specials) + r"""|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)
        self.vocab_size = len(self.encoder)
        self.all_special_ids = [self.encoder[t] for t in specials]

    # -- one pre-token -> space-joined BPE symbols
    def bpe(self, token: str) -> str:
        hit = self.cache.get(token)
        if hit is not None:
            return hit
        word = list(token[:-1]) + [token[-1] + "</w>"]
        if len(word) == 1:
            return token + "</w>"
        inf = float("inf")
        while len(word) > 1:
            # the adjacent pair with the lowest merge rank; all of its occurrences are merged left to right in this round
            best = min(zip(word, word[1:]), key=lambda pr: self.bpe_ranks.get(pr, inf))
            if best not in self.bpe_ranks:
                break
            a, b = best
            merged, i, n = [], 0, len(word)
            while i < n:
                if i + 1 < n and word[i] == a and word[i + 1] == b:
                    merged.append(a + b); i += 2
                else:
                    merged.append(word[i]); i += 1
            word = merged
        out = " ".join(word)
        self.cache[token] = out
        return out

    @staticmethod
    def clean(text: str) -> str:
        text = html.unescape(html.unescape(_fix_text(text))).strip()
        return re.sub(r"\s+", " ", text).strip()

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for tok in re.findall(self.pat, self.clean(text).lower()):
            sym = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self.bpe(sym).split(" "))
        return ids

    def decode(self, tokens) -> str:
        text = "".join(self.decoder[int(t)] for t in tokens)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")


def openclip_tokenize(texts: Union[str, List[str]], context_length: int = 77, _tokenizer: SimpleTokenizer = None):
    """[n, context_length] int64: <start_of_text> ids <end_of_text>, cut to context_length (the cut may drop <end_of_text>, as in the
    reference), zero padded (appzoo/clip/data.py:137-163)"""
    import torch
    if isinstance(texts, str):
        texts = [texts]
    sot, eot = _tokenizer.encoder["<start_of_text>"], _tokenizer.encoder["<end_of_text>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        row = ([sot] + _tokenizer.encode(text) + [eot])[:context_length]
        result[i, :len(row)] = torch.tensor(row)
    return result
