"""CPU: host-side logic of the product package (parameter schema / flat layout / registry), no kernels."""
import json
import os

import numpy as np
import pytest
import torch

from easynlp_b200.params import ParamStore, param_schema, uses_weight_decay, NO_GRAD
from oracle import clip_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_schema_equals_reference_checkpoint_keys():
    z = np.load(os.path.join(GOLD, "tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    ref = {k[2:]: z[k].shape for k in z.files if k.startswith("w.")}
    ref.pop("bert.embeddings.position_ids")
    sch = param_schema(cfg)
    assert set(sch) == set(ref)
    for k, shp in sch.items():
        assert tuple(shp) == tuple(ref[k]), k
    b16 = param_schema(O.vit_b16_bert_base_config())
    assert len(b16) == 353                              # + position_ids buffer = the 354 entries of SURVEY.md A.3
    n = sum(int(np.prod(s)) if len(s) else 1 for s in b16.values())
    assert n == 188_847_361 or abs(n - 188.85e6) < 0.01e6, n


def test_flat_layout_groups_and_alignment():
    cfg = O.tiny_config()
    st = ParamStore(cfg, device="cpu")
    offs = st.offsets
    for n, o in offs.items():
        assert o % 64 == 0, n
    for n in st.schema:
        if n in NO_GRAD:
            assert offs[n] >= st.n_trainable
        elif uses_weight_decay(n):
            assert offs[n] < st.n_decay
        else:
            assert st.n_decay <= offs[n] < st.n_trainable
    # BERT q/k/v are adjacent so that [3H, H] views exist
    H = cfg["text_hidden_size"]
    p = "bert.encoder.layer.1.attention.self."
    assert offs[p + "key.weight"] - offs[p + "query.weight"] == H * H
    assert offs[p + "value.weight"] - offs[p + "key.weight"] == H * H
    assert offs[p + "key.bias"] - offs[p + "query.bias"] == H
    fused = st.p(p + "query.weight", (3 * H, H))
    st.p(p + "value.weight").fill_(3.0)
    assert float(fused[2 * H:].min()) == 3.0 and float(fused[:2 * H].abs().max()) == 0.0
    # decay rule mirrors the oracle's restatement of optimizers.py:519-523
    for n in st.schema:
        assert uses_weight_decay("chinese_clip." + n) == O.uses_weight_decay("chinese_clip." + n)


def test_registry_prefix_match_and_unknown_app():
    from easynlp_b200.appzoo import api
    with pytest.raises(NotImplementedError):
        api.get_application_model("sequence_classification", "/tmp/x")
    assert api._classes("clip")[0].__name__ == "CLIPApp" and api._classes("clip_finetune")[0].__name__ == "CLIPApp"
    assert api._classes("wukong_clip")[0].__name__ == "WukongCLIP" and api._model_cls("clip4clip").__name__ == "Text2VideoRetrieval"


def test_synthetic_batch_layout():
    from easynlp_b200.synthetic import synthetic_batch, random_state_dict
    cfg = O.tiny_config()
    px, ids = synthetic_batch(cfg, 5, 16, seed=3)
    assert px.shape == (5, 3, 64, 64) and ids.shape == (5, 16) and ids.dtype == torch.int64
    assert (ids[:, 0] == 101).all()
    for r in ids:
        nz = (r != 0).nonzero().flatten()
        assert nz[-1].item() == len(nz) - 1           # contiguous tokens then 0-padding
    sd = random_state_dict(cfg, seed=1)
    assert set(sd) == set(param_schema(cfg))
    assert float(sd["bert.embeddings.word_embeddings.weight"][0].abs().max()) == 0.0


def test_predictor_feature_formats_and_npy_sink(tmp_path):
    """postprocess contract (reference predictor.py:140-153) + the binary sink behind it; no GPU: the encoder is stubbed"""
    from easynlp_b200.appzoo.clip.predictor import CLIPPredictor
    from easynlp_b200.core.predictor import Predictor, SimplePredictorManager
    emb = torch.arange(12, dtype=torch.float32).view(3, 4) / 7.0
    p = CLIPPredictor.__new__(CLIPPredictor)                       # no checkpoint / GPU: only postprocess is exercised
    p.feature_format = "text"
    out = p.postprocess({"image_embeds": None, "text_embeds": emb})
    assert [list(o) for o in out] == [["text_feat"]] * 3
    assert out[1]["text_feat"] == "\t".join(str(x) for x in emb[1].numpy())         # byte-identical to the reference formatting
    p.feature_format = "numpy"
    out = p.postprocess({"image_embeds": emb, "text_embeds": emb * 2})
    assert list(out[0]) == ["image_feat"] and out[2]["image_feat"].dtype == np.float32 and np.array_equal(out[2]["image_feat"], emb[2].numpy())

    class Stub(Predictor):                                          # rows -> vectors that depend on the row, in batches of 2
        def __init__(self, fmt): self.fmt = fmt
        def run(self, rows):
            vs = [np.full(4, float(r["idx"]), np.float32) for r in rows]
            return [{"text_feat": v if self.fmt == "numpy" else "\t".join(str(x) for x in v)} for v in vs]

    src = tmp_path / "in.tsv"
    src.write_text("".join(f"{i}\ttext {i}\n" for i in range(5)))
    dst = str(tmp_path / "feat.npy")
    SimplePredictorManager(Stub("numpy"), str(src), "idx:str:1,first_sequence:str:1", dst, "text_feat", "idx", batch_size=2).run()
    arr = np.load(dst)
    assert arr.shape == (5, 4) and arr.dtype == np.float32 and np.array_equal(arr[:, 0], np.arange(5, dtype=np.float32))
    assert open(dst + ".tsv").read().split() == [str(i) for i in range(5)]
    with pytest.raises(TypeError):
        SimplePredictorManager(Stub("text"), str(src), "idx:str:1,first_sequence:str:1", dst, "text_feat", "", batch_size=2).run()
    txt = str(tmp_path / "feat.tsv")
    SimplePredictorManager(Stub("text"), str(src), "idx:str:1,first_sequence:str:1", txt, "text_feat", "idx", batch_size=2).run()
    lines = open(txt).read().splitlines()
    assert len(lines) == 5 and lines[3].split("\t") == ["3.0"] * 4 + ["3"]


def test_split_score_error_bound():
    """Arithmetic of clipk_retrieval_rank_tc / the contrastive logits restated on the host: fp32 unit vectors split into
    bf16 hi + lo, score = hi.hi + hi.lo + lo.hi accumulated in fp32.  The error against fp64 stays below 3e-6, the hit@K decisions of a
    retrieval run equal the fp64 ones, and every rank lies inside the band that near-ties of that size allow."""
    g = torch.Generator().manual_seed(0)
    N, E, nq = 16384, 512, 512
    img = torch.nn.functional.normalize(torch.randn(N, E, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(img + (7.0 / E ** 0.5) * torch.randn(N, E, generator=g), dim=-1)
    sub = torch.arange(0, N, N // nq)[:nq]
    q = txt[sub]
    qh = q.bfloat16().float(); ql = (q - qh).bfloat16().float()
    kh = img.bfloat16().float(); kl = (img - kh).bfloat16().float()
    s = qh @ kh.t() + qh @ kl.t() + ql @ kh.t()
    s64 = q.double() @ img.double().t()
    assert (s.double() - s64).abs().max().item() < 3e-6
    other = sub[:, None] != torch.arange(N)[None, :]
    thr32 = (q * img[sub]).sum(-1, keepdim=True); thr64 = (q.double() * img[sub].double()).sum(-1, keepdim=True)
    got = ((s > thr32) & other).sum(1)
    exact = ((s64 > thr64) & other).sum(1)
    for k in (1, 5, 10):
        assert torch.equal(got < k, exact < k)
    lo = ((s64 > thr64 + 4e-6) & other).sum(1); hi = ((s64 > thr64 - 4e-6) & other).sum(1)
    assert bool(((lo <= got) & (got <= hi)).all()) and 0.05 < (exact < 1).float().mean().item() < 0.95


def test_recall_report_lines(capsys):
    """the evaluators' shared report: return value and the reference's three printed lines (appzoo/clip/evaluator.py:62-70)"""
    from easynlp_b200.appzoo.clip.evaluator import summarize_recall
    out = summarize_recall({1: 3, 5: 4, 10: 8}, 8, 0.5)
    assert out == [("mean_recall", (3 / 8 + 4 / 8 + 1.0) / 3.0)]
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "r1_num:3 r5_num:4 r10_num:8 query_num:8"
    assert lines[1] == "r1(%):" + str(3 / 8 * 100) + " r5(%):" + str(50.0) + " r10(%):" + str(100.0) + " mean_recall(%):" + str((3 / 8 + 4 / 8 + 1.0) / 3.0 * 100)
    assert lines[2] == "Inference time = 0.50s, [62.5000 ms / sample] "


def test_bench_traffic_is_tied_to_the_library_build(tmp_path, monkeypatch):
    """bench.py reports roofline.traffic from the committed ncu launch list ONLY when that profile was captured with the library that is
    loaded now (sha256 recorded by tools/launch_summary.py); the committed profile matches the current sources' build stamp."""
    import json
    import bench
    from easynlp_b200 import build as B
    val, why = bench._ncu_gemm_traffic()
    stamp = open(os.path.join(B.LIBDIR, "libclipk.sha256")).read().strip()
    prof = json.load(open(os.path.join(bench.ROOT, bench.PROFILE_JSON)))
    if prof["libclipk_sha256"] == stamp:
        assert val == prof["gemm"]["dram_bytes_per_launch"] and 1e8 < val < 1e9 and "same libclipk.so" in why
    else:
        assert val is None and "refused" in why
    # a profile of another build is refused
    fake = dict(prof, libclipk_sha256="0" * 64)
    p = tmp_path / "fake.json"; p.write_text(json.dumps(fake))
    monkeypatch.setattr(bench, "PROFILE_JSON", os.path.relpath(str(p), bench.ROOT))
    val, why = bench._ncu_gemm_traffic()
    assert val is None and "refused" in why
    # FLOP constants of the roofline
    assert bench.FLOPS_FWD_PER_PAIR == 35_126_906_880 + 13_300_469_760 and bench.FLOPS_TRAIN_PER_PAIR == 145_282_129_920
