"""Golden tokenisations from the UNMODIFIED reference BertTokenizer (easynlp/modelzoo/models/bert/tokenization_bert.py) and
image-preprocessing vectors from the reference's _resize/_center_crop/_normalize (appzoo/clip/data.py:29-135).
Run in the build container:  python oracle/make_tokenizer_golden.py   ->  tests/golden/tokenizer.json, preprocess.npz"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
os.environ.setdefault("HOME", "/root")

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "a", "the", "cat", "dog", "sit", "##s", "##ting", "on", "mat", "red", "bike", "##r",
         "一", "只", "猫", "狗", "在", "沙", "发", "上", "红", "色", "的", "自", "行", "车", ",", ".", "!", "?", "-", "hello", "world", "##ly", "cafe",
         "2023", "20", "##23", "photo", "of", "##a", "un", "##believ", "##able", "(", ")", "/", "resume", "naive", "mm", "##m"]
TEXTS = ["The cat sits on the mat.", "一只猫在沙发上", "红色的自行车, hello WORLD!", "A dog sitting on a RED biker", "unbelievable café photo-of 2023?",
         "  multiple   spaces\tand\nnewlines ", "猫dog狗cat", "[CLS] the [MASK] cat [SEP]", "résumé naïve (mmm) 20/23", "",
         "x" * 120, "the " * 40]


def main():
    out = os.path.join(ROOT, "tests", "golden")
    from easynlp.modelzoo.models.bert.tokenization_bert import BertTokenizer
    vocab_path = os.path.join(out, "tokenizer_vocab.txt")
    with open(vocab_path, "w", encoding="utf-8") as f:
        f.write("\n".join(VOCAB) + "\n")
    tok = BertTokenizer.from_pretrained(vocab_path)
    cases = []
    for L in (16, 32):
        for t in TEXTS:
            r = tok([t], padding="max_length", truncation=True, max_length=L, return_tensors="pt")
            cases.append({"text": t, "max_length": L, "input_ids": r["input_ids"][0].tolist(),
                          "attention_mask": r["attention_mask"][0].tolist(), "tokens": tok.tokenize(t)})
    with open(os.path.join(out, "tokenizer.json"), "w", encoding="utf-8") as f:
        json.dump({"vocab": VOCAB, "cases": cases}, f, ensure_ascii=False, indent=0)
    print("tokenizer cases:", len(cases))

    # image preprocessing: import the reference module with the shims SURVEY.md 8c lists
    for name, sub in (("easynlp.appzoo", "easynlp/appzoo"), ("easynlp.appzoo.clip", "easynlp/appzoo/clip")):
        m = types.ModuleType(name); m.__path__ = [os.path.join(REF, sub)]; sys.modules[name] = m
    try:
        import importlib.util
        src = open(os.path.join(REF, "easynlp/appzoo/clip/data.py")).read()
        # execute only the three pure functions (the module's imports need datasets/ftfy stubs)
        start = src.index("def _center_crop"); end = src.index("def openclip_tokenize")
        ns = {}
        exec("import numpy as np\nfrom PIL import Image\ndef is_torch_tensor(x):\n    return False\n" + src[start:end], ns)
        from PIL import Image
        rng = np.random.RandomState(0)
        blob = {}
        for i, (w, h) in enumerate([(300, 200), (200, 300), (224, 224), (640, 481), (100, 90)]):
            arr = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
            img = Image.fromarray(arr)
            x = ns["_resize"](image=img, size=224, resample=Image.BICUBIC)
            x = ns["_center_crop"](x, 224)
            x = ns["_normalize"](image=x)
            blob[f"in{i}"] = arr; blob[f"out{i}"] = np.asarray(x, dtype=np.float32)
        np.savez_compressed(os.path.join(out, "preprocess.npz"), **blob)
        print("preprocess cases written")
    except Exception as e:  # pragma: no cover
        print("preprocess golden skipped:", repr(e))


if __name__ == "__main__":
    main()
