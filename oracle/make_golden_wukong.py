"""Generate tests/golden/wukong_tiny.npz and tests/golden/wukong_tokenizer.json by running the UNMODIFIED reference's Wukong application
(appzoo/wukong_clip/model.py:22-88 -> WukongModel, modelzoo/models/wukong/modeling_wukong.py:234-413) and its FullTokenizer
(appzoo/wukong_clip/bert_tokenizer.py:166-396) in the build container.  Test infrastructure only.

    python oracle/make_golden_wukong.py

Pins inputs (one [SEP] = 102 per text at a different position per row), every weight, the reference's image / text features, loss and
every gradient; the oracle restatement (oracle.clip_oracle.wukong_forward) is checked against the reference before writing."""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import clip_oracle as O  # noqa: E402
from oracle.ref_loader import import_reference  # noqa: E402

TEXTS = ["The cat sits on the mat.", "一只猫在沙发上", "红色的自行车, hello WORLD!", "[CLS] the [MASK] cat [SEP]", "résumé naïve (mmm) 20/23", "",
         "x" * 120, "m" * 150, "the " * 40, "猫dog狗cat unbelievable"]


def main():
    R = import_reference()
    root = R["root"]
    m = types.ModuleType("easynlp.appzoo.wukong_clip"); m.__path__ = [os.path.join(root, "easynlp/appzoo/wukong_clip")]
    sys.modules["easynlp.appzoo.wukong_clip"] = m
    from easynlp.appzoo.wukong_clip.model import WukongCLIP
    from easynlp.appzoo.wukong_clip.bert_tokenizer import FullTokenizer

    raw = O.wukong_tiny_config()
    sd = O.wukong_init_state_dict(raw, seed=23, scale_boost=2.0)
    g = torch.Generator().manual_seed(23)
    B, L = 6, raw["model"]["text"]["context_length"]
    pixels = torch.randn(B, 3, 64, 64, generator=g)
    lens = torch.tensor([32, 9, 12, 3, 17, 5])
    ids = torch.randint(103, 500, (B, L), generator=g)
    ids[:, 0] = 101
    ids = torch.where(torch.arange(L)[None, :] < lens[:, None], ids, torch.zeros_like(ids))
    ids[torch.arange(B), lens - 1] = 102
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(raw, f)
        torch.save({"model." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
        app = WukongCLIP(d)
    app.train()
    out, _ = app({"pixel_values": pixels.clone(), "input_ids": ids.clone()})
    loss = app.compute_loss(out, [])["loss"]
    app.zero_grad()
    loss.backward()
    grads = {n[len("model."):]: p.grad.detach().clone() for n, p in app.named_parameters() if p.grad is not None}
    o = O.wukong_forward(sd, raw, pixels, ids)
    for k in ("image_features", "text_features"):
        err = (o[k] - out[k].detach()).abs().max().item()
        print(f"oracle vs reference {k}: max err {err:.2e}")
        assert err < 2e-5, k
    o_loss = O.clip_loss(o["logits_per_text"])
    assert abs(o_loss.item() - loss.item()) < 1e-5 * max(1.0, abs(loss.item())), (o_loss.item(), loss.item())
    blob = {"cfg_json": np.frombuffer(json.dumps(raw).encode(), dtype=np.uint8), "pixels": pixels.numpy(), "ids": ids.numpy(),
            "out.loss": loss.detach().numpy(), "out.logit_scale": out["logit_scale"].detach().numpy()}
    for k in ("image_features", "text_features"):
        blob["out." + k] = out[k].detach().numpy()
    for k, v in sd.items():
        blob["w." + k] = v.numpy()
    for k, v in grads.items():
        blob["g." + k] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "wukong_tiny.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB;", len(grads), "gradient tensors of", len(sd))

    # FullTokenizer + the dataset's tokenize rule (data.py:166-187) on the vocabulary of the CLIP tokenizer fixture
    vocab = os.path.join(ROOT, "tests", "golden", "tokenizer_vocab.txt")
    tok = FullTokenizer(vocab_file=vocab)
    cases = []
    for t in TEXTS:
        toks = tok.tokenize(t)
        row = [tok.vocab["[CLS]"]] + tok.convert_tokens_to_ids(toks)[:30] + [tok.vocab["[SEP]"]]
        cases.append({"text": t, "tokens": toks, "input_ids": row + [0] * (32 - len(row))})
    with open(os.path.join(ROOT, "tests", "golden", "wukong_tokenizer.json"), "w", encoding="utf-8") as f:
        json.dump({"cases": cases}, f, ensure_ascii=False, indent=0)
    print("tokenizer cases:", len(cases))


if __name__ == "__main__":
    main()
