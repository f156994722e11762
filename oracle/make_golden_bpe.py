"""Golden vectors for the open_clip BPE tokenizer from the UNMODIFIED reference class (modelzoo/models/clip/openclip_tokenizer.py:71-157 and
openclip_tokenize, appzoo/clip/data.py:137-163).  Run in the build container:  python oracle/make_golden_bpe.py

The reference imports `ftfy`, which this image lacks: the generator installs a stub module whose fix_text is Unicode NFC normalisation
(what easynlp_b200/bpe_tokenizer.py does in the same situation).  The merges file is synthetic: BPE trained right here on a small
corpus (pretrained vocabularies are not reachable), written as tests/golden/bpe_merges.txt.gz in the reference's file format."""
import collections
import gzip
import json
import os
import sys
import types
import unicodedata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CORPUS = ("a photo of a cat sitting on the mat . a photo of the red bike , isn't it ? the dog's ball rolled under the table ! "
          "photos of cats and dogs : 2023 was a great year for photography . she'll say they've seen it ; i'm sure we're done . "
          "naïve café résumé über straße 猫 和 狗 在 沙发 上 the quick brown fox jumps over the lazy dog " * 3)
TEXTS = ["a photo of a cat", "A Photo of THE red bike, isn't it?", "  multiple   spaces\tand\nnewlines ", "the dog's ball (2023) rolled!!!",
         "naïve café résumé", "café with a combining accent", "猫和狗在沙发上", "&amp;lt;b&amp;gt; html &quot;entities&quot; &amp;amp; more",
         "<start_of_text> literal specials <end_of_text>", "", "unseenwordzzzqqq 12345 x", "emoji 🙂 and symbols ©®™ ...", "word " * 100]


def train_merges(corpus, n):
    from collections import Counter
    import regex as re
    pat = re.compile(r"""'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)
    sys.path.insert(0, ROOT)
    from easynlp_b200.bpe_tokenizer import byte_symbols
    enc = byte_symbols()
    words = Counter()
    for tok in re.findall(pat, corpus.lower()):
        s = [enc[b] for b in tok.encode("utf-8")]
        s[-1] += "</w>"
        words[tuple(s)] += 1
    merges = []
    for _ in range(n):
        pairs = Counter()
        for w, c in words.items():
            for p in zip(w, w[1:]):
                pairs[p] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1]); i += 2
                else:
                    out.append(w[i]); i += 1
            new[tuple(out)] += c
        words = new
    return merges


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    merges = train_merges(CORPUS, 300)
    path = os.path.join(gold, "bpe_merges.txt.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(("#version: synthetic\n" + "\n".join(" ".join(m) for m in merges) + "\n").encode("utf-8"))
    stub = types.ModuleType("ftfy"); stub.fix_text = lambda t: unicodedata.normalize("NFC", t)
    sys.modules["ftfy"] = stub
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_openclip_tokenizer", os.path.join(REF, "easynlp/modelzoo/models/clip/openclip_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    tok = mod.SimpleTokenizer(bpe_path=path)
    # openclip_tokenize lives in appzoo/clip/data.py whose imports need the whole package: execute just that function's source
    src = open(os.path.join(REF, "easynlp/appzoo/clip/data.py")).read()
    a = src.index("def openclip_tokenize"); b = src.index("class CLIPDataset")
    ns = {}
    exec("import torch\nfrom typing import Union, List\n" + src[a:b], ns)
    cases = []
    for t in TEXTS:
        ids = tok.encode(t)
        cases.append({"text": t, "ids": ids, "decoded": tok.decode(ids), "row24": ns["openclip_tokenize"]([t], 24, tok)[0].tolist(),
                      "row77": ns["openclip_tokenize"](t, 77, tok)[0].tolist()})
    with open(os.path.join(gold, "bpe_tokenizer.json"), "w", encoding="utf-8") as f:
        json.dump({"vocab_size": tok.vocab_size, "sot": tok.encoder["<start_of_text>"], "eot": tok.encoder["<end_of_text>"], "cases": cases},
                  f, ensure_ascii=False, indent=0)
    print("merges", len(merges), "vocab", tok.vocab_size, "cases", len(cases), "max len", max(len(c["ids"]) for c in cases))


if __name__ == "__main__":
    main()
