"""CPU: the data formats either side of the hot path -- WordPiece tokenizer and image preprocessing against vectors
produced by the UNMODIFIED reference (oracle/make_tokenizer_golden.py), and the dataset collate contract."""
import base64
import io
import json
import os

import numpy as np
import torch
from PIL import Image

from easynlp_b200.tokenization import BertTokenizer
from easynlp_b200.appzoo.clip import data as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_wordpiece_matches_reference_tokenizer():
    g = json.load(open(os.path.join(GOLD, "tokenizer.json"), encoding="utf-8"))
    tok = BertTokenizer.from_pretrained(os.path.join(GOLD, "tokenizer_vocab.txt"))
    assert len(g["cases"]) >= 20
    for c in g["cases"]:
        r = tok([c["text"]], padding="max_length", truncation=True, max_length=c["max_length"], return_tensors="pt")
        assert tok.tokenize(c["text"]) == c["tokens"], c["text"]
        assert r["input_ids"][0].tolist() == c["input_ids"], c["text"]
        assert r["attention_mask"][0].tolist() == c["attention_mask"], c["text"]
        assert r["input_ids"].shape == (1, c["max_length"]) and r["input_ids"].dtype == torch.int64


def test_image_preprocessing_matches_reference():
    z = np.load(os.path.join(GOLD, "preprocess.npz"))
    n = len([k for k in z.files if k.startswith("in")])
    assert n >= 5
    for i in range(n):
        x = D.preprocess_image(Image.fromarray(z[f"in{i}"]))
        assert x.shape == (1, 3, 224, 224) and x.dtype == torch.float32
        assert np.allclose(x[0].numpy(), z[f"out{i}"], atol=1e-6), i


def test_dataset_rows_and_collate(tmp_path):
    model_dir = tmp_path / "m"; model_dir.mkdir()
    (model_dir / "config.json").write_text(json.dumps({"model_type": "chinese_clip"}))
    vocab = open(os.path.join(GOLD, "tokenizer_vocab.txt"), encoding="utf-8").read()
    (model_dir / "vocab.txt").write_text(vocab, encoding="utf-8")
    rows = []
    rng = np.random.RandomState(1)
    for t in ("the cat", "一只猫", "red bike"):
        buf = io.BytesIO(); Image.fromarray(rng.randint(0, 255, (50, 70, 3)).astype(np.uint8)).save(buf, format="PNG")
        rows.append(t + "\t" + base64.urlsafe_b64encode(buf.getvalue()).decode())
    tsv = tmp_path / "d.tsv"; tsv.write_text("\n".join(rows) + "\n", encoding="utf-8")
    ds = D.CLIPDataset(str(model_dir), str(tsv), 16, input_schema="text:str:1,image:str:1", first_sequence="text", second_sequence="image")
    assert len(ds) == 3
    b = ds.batch_fn([ds[0], ds[1], ds[2]])
    assert b["pixel_values"].shape == (3, 3, 224, 224) and b["input_ids"].shape == (3, 16) and b["label_ids"] == []
    assert set(b) == {"pixel_values", "input_ids", "token_type_ids", "attention_mask", "label_ids"}


def test_bpe_tokenizer_matches_reference_vectors():
    """open_clip's byte-level BPE (modelzoo/models/clip/openclip_tokenizer.py:71-157) and openclip_tokenize (appzoo/clip/data.py:137-163)
    against vectors produced by the unmodified reference class (oracle/make_golden_bpe.py): ids, decode round trip, padded / cut rows"""
    from easynlp_b200.bpe_tokenizer import SimpleTokenizer, openclip_tokenize
    g = json.load(open(os.path.join(GOLD, "bpe_tokenizer.json"), encoding="utf-8"))
    t = SimpleTokenizer(os.path.join(GOLD, "bpe_merges.txt.gz"))
    assert t.vocab_size == g["vocab_size"] and t.encoder["<start_of_text>"] == g["sot"] and t.encoder["<end_of_text>"] == g["eot"]
    assert t.all_special_ids == [g["sot"], g["eot"]] and len(g["cases"]) == 13
    for c in g["cases"]:
        ids = t.encode(c["text"])
        assert ids == c["ids"], c["text"]
        assert t.decode(ids) == c["decoded"]
        assert openclip_tokenize([c["text"]], 24, t)[0].tolist() == c["row24"]
        assert openclip_tokenize(c["text"], 77, t)[0].tolist() == c["row77"]
    rows = openclip_tokenize([c["text"] for c in g["cases"]], 77, t)
    assert rows.shape == (13, 77) and rows.dtype == torch.int64 and rows.tolist() == [c["row77"] for c in g["cases"]]
    # the EOT token carries the highest id, so the tower's argmax pooling finds it whenever the row was not cut
    short = [i for i, c in enumerate(g["cases"]) if len(c["ids"]) + 2 <= 77 and "<end_of_text>" not in c["text"]]
    assert all(rows[i].argmax().item() == len(g["cases"][i]["ids"]) + 1 for i in short)


def test_openclip_dataset_rows(tmp_path):
    model_dir = tmp_path / "oc"; model_dir.mkdir()
    (model_dir / "config.json").write_text(json.dumps({"model_type": "open_clip"}))
    import shutil
    shutil.copy(os.path.join(GOLD, "bpe_merges.txt.gz"), model_dir / "vocab.txt")
    rng = np.random.RandomState(2)
    rows = []
    for t in ("a photo of a cat", "the red bike"):
        buf = io.BytesIO(); Image.fromarray(rng.randint(0, 255, (60, 40, 3)).astype(np.uint8)).save(buf, format="PNG")
        rows.append(t + "\t" + base64.urlsafe_b64encode(buf.getvalue()).decode())
    tsv = tmp_path / "d.tsv"; tsv.write_text("\n".join(rows) + "\n", encoding="utf-8")
    ds = D.CLIPDataset(str(model_dir), str(tsv), 16, input_schema="text:str:1,image:str:1", first_sequence="text", second_sequence="image")
    b = ds.batch_fn([ds[0], ds[1]])
    assert b["pixel_values"].shape == (2, 3, 224, 224) and b["input_ids"].shape == (2, 77)
    assert b["token_type_ids"] == [] and b["attention_mask"] == [] and b["label_ids"] == []
    g = json.load(open(os.path.join(GOLD, "bpe_tokenizer.json"), encoding="utf-8"))
    assert b["input_ids"][0].tolist() == g["cases"][0]["row77"]
