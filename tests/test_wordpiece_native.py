"""CPU: the native WordPiece tokenizer (clipk_wp_*, host C++) against (a) the vectors produced by the UNMODIFIED reference tokenizer
(tests/golden/tokenizer.json) and (b) the Python restatement on randomly generated text mixing ASCII, CJK, Latin accents, full-width forms,
punctuation, controls and unsupported code points (fallback rule)."""
import json
import os
import random

import pytest

from easynlp_b200.tokenization import BertTokenizer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def tok():
    from easynlp_b200 import build as B
    B.build()
    t = BertTokenizer.from_pretrained(os.path.join(GOLD, "tokenizer_vocab.txt"))
    assert t._native_handle(), "native encoder must load (library + unicode table are built in-tree)"
    return t


def test_native_matches_reference_vectors(tok):
    cases = json.load(open(os.path.join(GOLD, "tokenizer.json"), encoding="utf-8"))["cases"]
    for c in cases:
        ids, mask, status = tok.encode_native([c["text"]], c["max_length"])
        assert status[0] >= 0, c["text"]
        assert ids[0].tolist() == c["input_ids"] and mask[0].tolist() == c["attention_mask"], c["text"]
        assert status[0] == sum(c["attention_mask"])


def test_native_equals_python_restatement_on_random_text(tok):
    rnd = random.Random(7)
    vocab = [v for v in tok.vocab if not v.startswith("[")]
    pools = [
        lambda: rnd.choice(vocab).replace("##", ""),
        lambda: "".join(rnd.choice("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789") for _ in range(rnd.randint(1, 12))),
        lambda: "".join(chr(rnd.randint(0x4E00, 0x4E80)) for _ in range(rnd.randint(1, 5))),
        lambda: rnd.choice(["é", "Ü", "ñ", "Çà", "Å", "ø", "ß", "İ", "ǅ", "ﬁ", "ｈｅｌｌｏ", "ＷＯＲＬＤ", "café", "naïve", "Ωmega", "Привет", "ÀÉÎÕÜ"]),
        lambda: rnd.choice([",", ".", "!", "?", "、", "。", "《", "》", "—", "…", "$", "^", "`", "~", "＄", "￥", "·", "«", "»", "'", "\""]),
        lambda: rnd.choice([" ", "  ", "\t", "\n", "　", " ", " ", "​", "\x07", "�", "́"]),
        lambda: rnd.choice(["[UNK]", "[SEP]", "[MASK]", "[CLS]", "x" * 120]),
    ]
    texts = []
    for _ in range(600):
        n = rnd.randint(0, 14)
        texts.append("".join(rnd.choice(pools)() + rnd.choice(["", " ", " "]) for _ in range(n)))
    for L in (8, 32):
        ids, mask, status = tok.encode_native(texts, L)
        assert (status >= 0).all()                       # everything above is inside the native tables
        cls_id, sep_id, pad_id = tok.vocab["[CLS]"], tok.vocab["[SEP]"], tok.vocab.get("[PAD]", 0)
        ref_ids, ref_mask = tok._encode_python(texts, "max_length", True, L, cls_id, sep_id, pad_id)
        bad = [i for i in range(len(texts)) if ids[i].tolist() != ref_ids[i] or mask[i].tolist() != ref_mask[i]]
        assert not bad, (texts[bad[0]], ids[bad[0]].tolist(), ref_ids[bad[0]])


def test_fallback_rule_and_call_path(tok):
    texts = ["a cat \U0001F600 smiles", "ΣΑΣ final sigma", "plain text", "\U00020000 CJK extension B is native"]
    ids, mask, status = tok.encode_native(texts, 16)
    assert status[0] == -100 and status[1] == -100 and status[2] >= 0 and status[3] >= 0
    before = dict(tok.native_stats)
    out = tok(texts, padding="max_length", truncation=True, max_length=16)
    assert tok.native_stats["fallback"] - before["fallback"] == 2 and tok.native_stats["native"] - before["native"] == 2
    cls_id, sep_id, pad_id = tok.vocab["[CLS]"], tok.vocab["[SEP]"], tok.vocab.get("[PAD]", 0)
    ref_ids, ref_mask = tok._encode_python(texts, "max_length", True, 16, cls_id, sep_id, pad_id)
    assert out["input_ids"].tolist() == ref_ids and out["attention_mask"].tolist() == ref_mask
