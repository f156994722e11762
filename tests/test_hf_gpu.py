"""GPU parity of the huggingface_clip branch (appzoo/clip/model.py:73-104,128-144; BASELINE configs[3] family): RobertaModel text tower
with pad-aware position ids / batch token types / batch attention mask / tanh pooler, frozen CLIPVisionModel image tower, biased
projections -- against the fixture written by the unmodified reference (oracle/make_golden_hf.py) and the oracle on fresh inputs."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from easynlp_b200.engine import ClipEngine, hf_engine_config  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def max_err(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def load():
    z = np.load(os.path.join(GOLD, "hf_tiny_fwd_bwd.npz"))
    raw = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    return z, raw, sd


def yardstick(sd, raw, pixels, ids, tt, am):
    names = [k for k, v in sd.items() if v.is_floating_point()]
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(params)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = O.hf_clip_forward(full, raw, pixels, ids, tt, am)
    out = {k: v.float() for k, v in out.items()}
    loss = O.clip_loss(out["logits_per_text"])
    g = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    return loss.item(), out, {k: v for k, v in zip(names, g) if v is not None}


def test_hf_tiny_forward_backward_vs_reference_golden():
    z, raw, sd = load()
    cfg = hf_engine_config(raw, sd["text_projection.weight"].shape[0])
    eng = ClipEngine(cfg)
    missing = eng.params.load_state_dict(sd)
    assert not missing
    pixels = torch.from_numpy(z["pixels"]); ids = torch.from_numpy(z["ids"]); tt = torch.from_numpy(z["token_type_ids"]); am = torch.from_numpy(z["attention_mask"])
    out = eng.forward(pixels.cuda(), ids.cuda(), token_type_ids=tt.cuda(), attention_mask=am.cuda())
    torch.cuda.synchronize()
    ref = {k: torch.from_numpy(z["out." + k]) for k in ("image_embeds", "text_embeds", "logits_per_text")}
    yl, yo, yg = yardstick(sd, raw, pixels, ids, tt, am)
    e = {k: max_err(out[k], ref[k]) for k in ref}; y = {k: max_err(yo[k], ref[k]) for k in ref}
    loss_ref = float(z["out.loss"]); loss = out["loss"].item()
    print(f"PARITY hf tiny fwd: err {e} (PyTorch bf16 yardstick {y}); loss {loss:.6f} vs {loss_ref:.6f}")
    # same bound as the chinese_clip tests: no worse than 1.5x PyTorch's own bf16 autocast on the same inputs
    assert e["image_embeds"] < 1.5 * y["image_embeds"] + 1e-4 and e["text_embeds"] < 1.5 * y["text_embeds"] + 1e-4
    assert e["logits_per_text"] < 1.5 * y["logits_per_text"] + 1e-3
    # 6-pair batch: the loss moves by at most max |d logit| (cross-entropy is 1-Lipschitz in the sup norm); bf16 logits are off by ~0.02 of
    # 14.3 here, PyTorch's own bf16 autocast by 0.04 -> rtol 2e-3 on this fixture (the B = 8 / B = 256 ViT-B/16 tests hold rtol 1e-3)
    assert abs(loss - loss_ref) < max(2e-3 * abs(loss_ref), 2.0 * abs(yl - loss_ref)) and abs(loss - loss_ref) <= e["logits_per_text"]
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    refg = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    gnorm = math.sqrt(sum(float(v.double().norm()) ** 2 for v in refg.values()))
    worst = (0.0, None)
    for k, r in refg.items():
        got = eng.params.g(k).detach().float().cpu().view_as(r)
        err = (got - r).norm().item(); yerr = (yg[k].float() - r).norm().item()
        tol = max(0.03 * r.norm().item(), 1.5 * yerr) + 1e-4 * gnorm
        worst = max(worst, (err / (r.norm().item() + 1e-4 * gnorm), k))
        assert err <= tol, f"grad {k}: err {err:.3e} > tol {tol:.3e} (|ref| {r.norm().item():.3e}, yardstick {yerr:.3e})"
    print(f"PARITY hf tiny bwd: worst per-tensor relative gradient error {worst[0]:.3e} ({worst[1]})")
    # the frozen tower's parameters sit outside the updated range and carry no gradient slots
    assert all(n.startswith("vision_encoder.") for n in eng.params.no_grad)
    assert "vision_encoder.vision_model.post_layernorm.weight" not in eng.params.trainable_names()
    before = eng.params.p("vision_encoder.vision_model.embeddings.class_embedding").clone()
    tb = eng.params.p("text_encoder.pooler.dense.weight").clone()
    eng.optimizer_step(lr=1e-3)
    torch.cuda.synchronize()
    assert torch.equal(before, eng.params.p("vision_encoder.vision_model.embeddings.class_embedding"))
    assert (eng.params.p("text_encoder.pooler.dense.weight") - tb).abs().max().item() > 1e-5      # the pooler IS trained on this branch


def test_hf_plugin_surface_and_defaults(tmp_path):
    """get_application_model on a huggingface_clip checkpoint directory: key names without a prefix, feat=True paths, missing
    token_type_ids / attention_mask (type 0, mask from ids != pad), Trainer step with dropout on."""
    from easynlp_b200.appzoo import get_application_model
    z, raw, sd = load()
    raw = json.loads(json.dumps(raw)); raw["text_config"]["hidden_dropout_prob"] = 0.1; raw["text_config"]["attention_probs_dropout_prob"] = 0.1
    d = str(tmp_path / "hf")
    os.makedirs(d)
    json.dump(raw, open(os.path.join(d, "config.json"), "w"))
    torch.save(sd, os.path.join(d, "pytorch_model.bin"))
    open(os.path.join(d, "vocab.txt"), "w").write("[PAD]\n[UNK]\n[CLS]\n[SEP]\n")
    model = get_application_model("clip", d, user_defined_parameters={"app_parameters": {}})
    assert model.model_type == "huggingface_clip"
    names = [n for n, _ in model.named_parameters()]
    assert "text_projection.bias" in names and "vision_encoder.vision_model.pre_layrnorm.weight" in names and "logit_scale" in names
    sd2 = model.state_dict()
    assert "text_encoder.embeddings.position_ids" in sd2 and torch.equal(sd2["vision_projection.weight"].cpu(), sd["vision_projection.weight"])
    ids = torch.from_numpy(z["ids"]); pixels = torch.from_numpy(z["pixels"])
    model.eval()
    with torch.no_grad():
        f = model({"input_ids": ids.clone()}, feat=True)               # no token types / mask given
        g = model({"pixel_values": pixels.clone()}, feat=True)
    ref = O.hf_clip_forward(sd, json.loads(bytes(z["cfg_json"]).decode()), pixels, ids, None, None)
    assert f["image_embeds"] is None and max_err(f["text_embeds"], ref["text_embeds"]) < 5e-3
    assert g["text_embeds"] is None and max_err(g["image_embeds"], ref["image_embeds"]) < 5e-3
    # one fused training step through the engine (dropout on): finite, moves only trainable tensors
    model.train()
    vis = model.engine.params.p("vision_encoder.vision_model.encoder.layers.0.mlp.fc1.weight").clone()
    out = model.engine.train_step(pixels, ids, lr=1e-3, use_graph=False, token_type_ids=torch.from_numpy(z["token_type_ids"]),
                                  attention_mask=torch.from_numpy(z["attention_mask"]))
    assert math.isfinite(out["loss"].item())
    assert torch.equal(vis, model.engine.params.p("vision_encoder.vision_model.encoder.layers.0.mlp.fc1.weight"))


def test_hf_patch14_257_tokens_forward_vs_oracle():
    """ViT-*/14 geometry of BASELINE configs[3] at a small width: 224 x 224 / patch 14 -> 257 tokens (the single-buffer 272-key attention
    path) and patch dim 3*14*14 = 588 padded to 592 for the TMA rows of the patch GEMM; forward only (the tower is frozen on this branch)."""
    raw = O.hf_tiny_config()
    raw = json.loads(json.dumps(raw)); raw["vision_config"].update(image_size=224, patch_size=14)
    sd = O.hf_init_state_dict(raw, seed=3, scale_boost=2.0)
    cfg = hf_engine_config(raw, 128)
    eng = ClipEngine(cfg, with_optimizer_state=False)
    assert eng.Lv == 257 and eng.kdim == 588 and eng.kdim_pad == 592
    eng.params.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    pixels = torch.randn(5, 3, 224, 224, generator=g)
    out = eng.encode(pixels.cuda(), None)["image_embeds"]
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.hf_clip_forward(sd, raw, pixels, None)["image_embeds"]
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yard = O.hf_clip_forward(sd, raw, pixels, None)["image_embeds"].float()
    e, y = max_err(out, ref), max_err(yard, ref)
    print(f"PARITY hf patch14 L=257 fwd: image embeds max err {e:.2e} (PyTorch bf16 yardstick {y:.2e})")
    assert e < 1.5 * y + 1e-4


# ------------------------------------------------------------------------------------------------ open_clip branch
def test_openclip_tiny_forward_backward_vs_reference_golden(tmp_path):
    """model_type == open_clip (appzoo/clip/model.py:56-63): ViT + causal pre-LN text transformer + EOT-argmax pooling -- the causal variants of
    the attention kernels, the EOT row gather / scatter, and every gradient against the fixture written by the unmodified reference."""
    from easynlp_b200.appzoo import get_application_model
    z = np.load(os.path.join(GOLD, "openclip_tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    d = str(tmp_path / "oc"); os.makedirs(d)
    json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    torch.save({"open_clip." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    model = get_application_model("clip", d, user_defined_parameters={"app_parameters": {}})
    assert model.model_type == "open_clip" and all(n.startswith("open_clip.") for n, _ in model.named_parameters())
    eng = model.engine
    pixels = torch.from_numpy(z["pixels"]); ids = torch.from_numpy(z["ids"])
    out = eng.forward(pixels.cuda(), ids.cuda())
    torch.cuda.synchronize()
    ref = {k: torch.from_numpy(z["out." + k]) for k in ("image_embeds", "text_embeds", "logits_per_text")}
    names = list(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yo = O.openclip_forward(params, cfg, pixels, ids)
    yo = {k: v.float() for k, v in yo.items()}
    yl = O.clip_loss(yo["logits_per_text"])
    yg = dict(zip(names, torch.autograd.grad(yl, [params[k] for k in names], allow_unused=True)))
    e = {k: max_err(out[k], ref[k]) for k in ref}; y = {k: max_err(yo[k], ref[k]) for k in ref}
    loss_ref = float(z["out.loss"]); loss = out["loss"].item()
    print(f"PARITY open_clip tiny fwd: err {e} (PyTorch bf16 yardstick {y}); loss {loss:.6f} vs {loss_ref:.6f}")
    assert e["image_embeds"] < 1.5 * y["image_embeds"] + 1e-4 and e["text_embeds"] < 1.5 * y["text_embeds"] + 1e-4
    assert e["logits_per_text"] < 1.5 * y["logits_per_text"] + 1e-3
    assert abs(loss - loss_ref) < max(2e-3 * abs(loss_ref), 2.0 * abs(yl.item() - loss_ref))
    eng.zero_grad(); eng.backward()
    torch.cuda.synchronize()
    refg = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    gnorm = math.sqrt(sum(float(v.double().norm()) ** 2 for v in refg.values()))
    worst = (0.0, None)
    for k, r in refg.items():
        got = eng.params.g(k).detach().float().cpu().view_as(r)
        err = (got - r).norm().item(); yerr = (yg[k].float() - r).norm().item()
        tol = max(0.03 * r.norm().item(), 1.5 * yerr) + 1e-4 * gnorm
        worst = max(worst, (err / (r.norm().item() + 1e-4 * gnorm), k))
        assert err <= tol, f"grad {k}: err {err:.3e} > tol {tol:.3e} (|ref| {r.norm().item():.3e}, yardstick {yerr:.3e})"
    print(f"PARITY open_clip tiny bwd: worst per-tensor relative gradient error {worst[0]:.3e} ({worst[1]})")
    eng.optimizer_step(lr=1e-3)
    torch.cuda.synchronize()
    assert torch.isfinite(eng.params.master).all()


def test_text2video_retrieval_vs_reference_golden(tmp_path):
    """'clip4clip' sibling application (appzoo/text2video_retrieval/model.py:38-120): T frames per video through the open_clip image tower,
    masked mean over the frames, InfoNCE against the texts -- forward and a sample of gradients against the unmodified reference."""
    from easynlp_b200.appzoo import get_application_model
    zc = np.load(os.path.join(GOLD, "openclip_tiny_fwd_bwd.npz")); z = np.load(os.path.join(GOLD, "t2v_tiny.npz"))
    cfg = json.loads(bytes(zc["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(zc[k]) for k in zc.files if k.startswith("w.")}
    d = str(tmp_path / "t2v"); os.makedirs(d)
    json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    torch.save({"open_clip." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    model = get_application_model("clip4clip", d, user_defined_parameters={"app_parameters": {}})
    assert type(model).__name__ == "Text2VideoRetrieval"
    model.train()
    batch = {"pixel_values": torch.from_numpy(z["pixels"]), "video_masks": torch.from_numpy(z["video_masks"]), "input_ids": torch.from_numpy(z["ids"])}
    out = model(batch)
    assert set(out) == {"logits_per_text", "logits_per_video", "video_embeds", "text_embeds"} and batch["pixel_values"].shape[0] == 15
    assert max_err(out["video_embeds"], torch.from_numpy(z["out.video_embeds"])) < 6e-3
    assert max_err(out["text_embeds"], torch.from_numpy(z["out.text_embeds"])) < 6e-3
    loss = model.compute_loss(out, [])["loss"]
    assert abs(loss.item() - float(z["out.loss"])) < 5e-3 * float(z["out.loss"]) + 2e-3
    model.zero_grad(); loss.backward()
    for k in [f[2:] for f in z.files if f.startswith("g.")]:
        r = torch.from_numpy(z["g." + k]); got = model.engine.params.g(k).detach().float().cpu().view_as(r)
        assert (got - r).norm().item() < 0.08 * r.norm().item() + 1e-4, (k, (got - r).norm().item(), r.norm().item())
    model.eval()
    with torch.no_grad():
        f = model({"pixel_values": torch.from_numpy(z["pixels"]), "video_masks": torch.from_numpy(z["video_masks"])}, feat=True)
    assert f["text_embeds"] is None and max_err(f["video_embeds"], torch.from_numpy(z["out.video_embeds"])) < 6e-3


# ------------------------------------------------------------------------------------------------ wukong_clip sibling application
def _wukong_dir(tmp_path, z, with_vocab=False):
    raw = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    d = str(tmp_path / "wk"); os.makedirs(d)
    json.dump(raw, open(os.path.join(d, "config.json"), "w"))
    torch.save({"model." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    if with_vocab:     # BERT layout: [PAD] 0, [UNK] 100, [CLS] 101, [SEP] 102 -- the text tower pools the position of id 102
        words = ["[PAD]"] + [f"[unused{i}]" for i in range(1, 100)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]", "a", "red", "cat", "dog", "on", "the", "mat", "##s", "猫", "狗"]
        open(os.path.join(d, "vocab.txt"), "w", encoding="utf-8").write("\n".join(words) + "\n")
    return d, raw, sd


def test_wukong_tiny_forward_backward_vs_reference_golden(tmp_path):
    """WukongCLIP (appzoo/wukong_clip/model.py:22-88): ViT + causal TextTransformer pooled at [SEP], LayerNorm eps 1e-7 -- tuple output,
    features / loss / every gradient against the fixture written by the unmodified reference (oracle/make_golden_wukong.py)."""
    from easynlp_b200.appzoo import get_application_model
    z = np.load(os.path.join(GOLD, "wukong_tiny.npz"))
    d, raw, sd = _wukong_dir(tmp_path, z)
    model = get_application_model("wukong_clip", d, user_defined_parameters={})
    assert type(model).__name__ == "WukongCLIP" and all(n.startswith("model.") for n, _ in model.named_parameters())
    assert json.loads(model.config.to_json_string()) == raw
    assert set(model.state_dict()) == {"model." + k for k in sd}
    model.train()
    pixels = torch.from_numpy(z["pixels"]); ids = torch.from_numpy(z["ids"])
    out, extra = model({"pixel_values": pixels.clone(), "input_ids": ids.clone()})
    assert extra == [] and set(out) == {"image_features", "text_features", "logit_scale"}
    ref = {k: torch.from_numpy(z["out." + k]) for k in ("image_features", "text_features")}
    names = list(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yo = O.wukong_forward(params, raw, pixels, ids)
    yl = O.clip_loss(yo["logits_per_text"].float())
    yg = dict(zip(names, torch.autograd.grad(yl, [params[k] for k in names], allow_unused=True)))
    e = {k: max_err(out[k], ref[k]) for k in ref}; y = {k: max_err(yo[k].float(), ref[k]) for k in ref}
    loss = model.compute_loss((out, extra), [])["loss"]
    loss_ref = float(z["out.loss"])
    print(f"PARITY wukong tiny fwd: err {e} (PyTorch bf16 yardstick {y}); loss {loss.item():.6f} vs {loss_ref:.6f}")
    assert e["image_features"] < 1.5 * y["image_features"] + 1e-4 and e["text_features"] < 1.5 * y["text_features"] + 1e-4
    assert abs(out["logit_scale"].item() - float(z["out.logit_scale"])) < 1e-4
    assert abs(loss.item() - loss_ref) < max(2e-3 * abs(loss_ref), 2.0 * abs(yl.item() - loss_ref))
    model.zero_grad(); loss.backward()
    torch.cuda.synchronize()
    refg = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    gnorm = math.sqrt(sum(float(v.double().norm()) ** 2 for v in refg.values()))
    worst = (0.0, None)
    for k, r in refg.items():
        got = model.engine.params.g(k).detach().float().cpu().view_as(r)
        err = (got - r).norm().item(); yerr = (yg[k].float() - r).norm().item()
        tol = max(0.03 * r.norm().item(), 1.5 * yerr) + 1e-4 * gnorm
        worst = max(worst, (err / (r.norm().item() + 1e-4 * gnorm), k))
        assert err <= tol, f"grad {k}: err {err:.3e} > tol {tol:.3e} (|ref| {r.norm().item():.3e}, yardstick {yerr:.3e})"
    print(f"PARITY wukong tiny bwd: worst per-tensor relative gradient error {worst[0]:.3e} ({worst[1]})")
    # single-modality calls (model.py:58-69) and the [SEP] rule; 1e-2 = 1.4 x the PyTorch-bf16 yardstick printed above on unit vectors
    model.eval()
    with torch.no_grad():
        o1, _ = model({"input_ids": ids.clone()})
        o2, _ = model({"pixel_values": pixels.clone()})
    assert o1["image_features"] is None and max_err(o1["text_features"], ref["text_features"]) < 1e-2
    assert o2["text_features"] is None and max_err(o2["image_features"], ref["image_features"]) < 1e-2
    bad = ids.clone(); bad[0, 1] = 102
    with pytest.raises(ValueError):
        model({"input_ids": bad})
    # kernel: position and count of the pooled token
    from easynlp_b200 import ops
    idx = torch.empty(ids.shape[0], dtype=torch.int32, device="cuda"); cnt = torch.empty_like(idx)
    ops.find_token_rows(bad.cuda(), 102, idx, cnt)
    assert idx.tolist() == [1] + ((ids == 102).int().argmax(1)[1:]).tolist() and cnt.tolist() == [2, 1, 1, 1, 1, 1]


def test_wukong_dataset_evaluator_predictor(tmp_path):
    """the application's data format end to end (appzoo/wukong_clip/data.py, evaluator.py, predictor.py): TSV rows of text + base64 image ->
    WukongCLIPDataset.batch_fn -> WukongCLIPEvaluator (blocked ranking) and WukongCLIPPredictor rows, checked against the oracle on the
    same preprocessed tensors"""
    import base64
    from io import BytesIO
    from PIL import Image
    from easynlp_b200.appzoo import get_application_dataset, get_application_evaluator, get_application_model_for_evaluation, get_application_predictor
    z = np.load(os.path.join(GOLD, "wukong_tiny.npz"))
    d, raw, sd = _wukong_dir(tmp_path, z, with_vocab=True)
    raw["model"]["visual"]["input_resolution"] = 224      # the dataset crops to 224 x 224 (data.py:155-160): positional table for 14 x 14 + 1 tokens
    g = torch.Generator().manual_seed(5)
    sd["visual_encoder.positional_embedding"] = torch.randn(197, 128, generator=g) * 128 ** -0.5
    json.dump(raw, open(os.path.join(d, "config.json"), "w"))
    torch.save({"model." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    rng = np.random.RandomState(3)
    texts = ["a red cat", "the dog on the mat", "猫 on mats", "dogs", "a cat a dog", "红 unknown words", "the the the", "狗"]
    rows = []
    for t in texts:
        buf = BytesIO(); Image.fromarray(rng.randint(0, 256, (240, 300, 3)).astype(np.uint8)).save(buf, format="PNG")
        rows.append(t + "\t" + base64.urlsafe_b64encode(buf.getvalue()).decode())
    tsv = str(tmp_path / "valid.tsv"); open(tsv, "w", encoding="utf-8").write("\n".join(rows) + "\n")
    ds = get_application_dataset("wukong_clip", d, tsv, 32, input_schema="text:str:1,image:str:1", first_sequence="text", second_sequence="image")
    batch = ds.batch_fn([ds[i] for i in range(len(ds))])
    assert batch["pixel_values"].shape == (8, 3, 224, 224) and batch["input_ids"].shape == (8, 32)
    assert batch["input_ids"][0].tolist()[:5] == [101, 104, 105, 106, 102] and (batch["input_ids"] == 102).sum(1).tolist() == [1] * 8
    model = get_application_model_for_evaluation("wukong_clip", d)
    ev = get_application_evaluator("wukong_clip", ds, user_defined_parameters={}, eval_batch_size=4)
    res = ev.evaluate(model)
    o = O.wukong_forward(sd, raw, batch["pixel_values"], batch["input_ids"])
    with torch.no_grad():
        f, _ = model({"pixel_values": batch["pixel_values"].clone(), "input_ids": batch["input_ids"].clone()})
    assert max_err(f["text_features"], o["text_features"]) < 1e-2 and max_err(f["image_features"], o["image_features"]) < 1e-2
    r = O.rank_of_match(f["text_features"].double().cpu(), f["image_features"].double().cpu())     # the ranking the evaluator must reproduce
    want = sum(float((r < k).sum()) / 8 for k in (1, 5, 10)) / 3
    assert res[0][0] == "mean_recall" and abs(res[0][1] - want) < 1e-9
    assert get_application_evaluator("wukong_clip", ds, user_defined_parameters={"cosine_similarity": "True"}, eval_batch_size=8).evaluate(model) is None
    pred = get_application_predictor("wukong_clip", d, first_sequence="text", second_sequence="image")
    recs = pred.run([{"text": texts[1]}, {"text": texts[2]}])
    got = np.array([[float(x) for x in rec["text_feat"].split("\t")] for rec in recs])
    assert np.abs(got - o["text_features"][1:3].numpy()).max() < 1e-2
    recs = pred.run([{"image": rows[0].split("\t")[1]}])
    got = np.array([float(x) for x in recs[0]["image_feat"].split("\t")])
    assert np.abs(got - o["image_features"][0].numpy()).max() < 1e-2


def test_patch14_tower_trains(tmp_path):
    """ViT-*/14 image towers that DO train (chinese_clip / open_clip / Wukong ViT-L/14): the 588-wide patch rows are padded to 592 for TMA;
    the patch-embedding weight gradient and everything upstream against the oracle's autograd."""
    from easynlp_b200.appzoo import get_application_model
    raw = O.wukong_tiny_config()
    raw["model"]["visual"].update(patch_size=14, input_resolution=56)
    sd = O.wukong_init_state_dict(raw, seed=31, scale_boost=2.0)
    assert sd["visual_encoder.conv1.weight"].shape == (128, 3, 14, 14) and sd["visual_encoder.positional_embedding"].shape == (17, 128)
    d = str(tmp_path / "wk14"); os.makedirs(d)
    json.dump(raw, open(os.path.join(d, "config.json"), "w"))
    torch.save({"model." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    model = get_application_model("wukong_clip", d)
    assert model.engine.kdim == 588 and model.engine.kdim_pad == 592
    model.train()
    g = torch.Generator().manual_seed(31)
    B = 6
    pixels = torch.randn(B, 3, 56, 56, generator=g)
    ids = torch.randint(103, 500, (B, 32), generator=g); ids[:, 0] = 101
    lens = torch.tensor([32, 4, 11, 20, 7, 16])
    ids = torch.where(torch.arange(32)[None, :] < lens[:, None], ids, torch.zeros_like(ids)); ids[torch.arange(B), lens - 1] = 102
    names = list(sd)
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    ref = O.wukong_forward(params, raw, pixels, ids)
    ref_loss = O.clip_loss(ref["logits_per_text"])
    rg = dict(zip(names, torch.autograd.grad(ref_loss, [params[k] for k in names])))
    out, _ = model({"pixel_values": pixels.clone(), "input_ids": ids.clone()})
    loss = model.compute_loss(out, [])["loss"]
    assert max_err(out["image_features"], ref["image_features"]) < 1e-2 and abs(loss.item() - ref_loss.item()) < 5e-3 * ref_loss.item() + 2e-3
    for rep in range(2):      # twice: the padded scratch tile must not carry the first pass over
        model.zero_grad(); (loss if rep == 0 else model.compute_loss(model({"pixel_values": pixels.clone(), "input_ids": ids.clone()})[0], [])["loss"]).backward()
        for k in ("visual_encoder.conv1.weight", "visual_encoder.positional_embedding", "visual_encoder.class_embedding", "visual_encoder.proj",
                  "visual_encoder.transformer.resblocks.0.attn.in_proj_weight"):
            r = rg[k]; got = model.engine.params.g(k).detach().float().cpu().view_as(r)
            assert (got - r).norm().item() < 0.05 * r.norm().item() + 1e-5, (rep, k, (got - r).norm().item(), r.norm().item())


def test_clip4clip_dataset_evaluator_predictor(tmp_path):
    """Text2VideoRetrieval's data format end to end (appzoo/text2video_retrieval/data.py, evaluator.py, predictor.py): TSV rows of text + a
    directory of frames -> padded frame stacks + video masks -> evaluator / predictor, against the oracle on the host-preprocessed frames"""
    import shutil
    from PIL import Image
    from easynlp_b200.appzoo import get_application_dataset, get_application_evaluator, get_application_model_for_evaluation, get_application_predictor
    zc = np.load(os.path.join(GOLD, "openclip_tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(zc["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(zc[k]) for k in zc.files if k.startswith("w.")}
    g = torch.Generator().manual_seed(11)
    cfg.update(image_resolution=224, vocab_size=700, context_length=77)          # frames are cropped to 224 x 224; BPE ids reach 653
    sd["visual.positional_embedding"] = torch.randn(197, 128, generator=g) * 128 ** -0.5
    sd["token_embedding.weight"] = torch.randn(700, 128, generator=g) * 0.02
    sd["positional_embedding"] = torch.randn(77, 128, generator=g) * 0.01
    d = str(tmp_path / "t2v"); os.makedirs(d)
    json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    torch.save({"open_clip." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    shutil.copy(os.path.join(GOLD, "bpe_merges.txt.gz"), os.path.join(d, "vocab.txt"))
    rng = np.random.RandomState(6)
    texts = ["a photo of a cat", "the red bike", "the dog's ball rolled", "photos of cats and dogs"]
    rows, frames = [], []
    for i, t in enumerate(texts):
        fd = tmp_path / f"video{i}"; fd.mkdir()
        arrs = [rng.randint(0, 256, (90 + 10 * i, 120, 3)).astype(np.uint8) for _ in range((3, 12, 1, 7)[i])]
        for j, a in enumerate(arrs):
            Image.fromarray(a).save(fd / f"frame{j:02d}.png")
        frames.append(arrs); rows.append(t + "\t" + str(fd))
    tsv = str(tmp_path / "valid.tsv"); open(tsv, "w", encoding="utf-8").write("\n".join(rows) + "\n")
    kw = dict(input_schema="text:str:1,image:str:1", first_sequence="text", second_sequence="image")
    ds = get_application_dataset("clip4clip", d, tsv, 77, user_defined_parameters={"app_parameters": {"gpu_preprocess": True}}, **kw)
    ds_host = get_application_dataset("clip4clip", d, tsv, 77, **kw)
    batch = ds.batch_fn([ds[i] for i in range(4)]); bh = ds_host.batch_fn([ds_host[i] for i in range(4)])
    assert batch["pixel_values"].shape == (4, 12, 3, 224, 224) and batch["video_masks"].sum(1).tolist() == [3, 12, 1, 7] and batch["input_ids"].shape == (4, 77)
    assert torch.equal(batch["pixel_values"].cpu(), bh["pixel_values"]) and torch.equal(batch["video_masks"], bh["video_masks"])
    # oracle: per-frame image embeddings -> masked mean -> normalise (text2video_retrieval/model.py:82-88); causal text tower with EOT pooling
    px = bh["pixel_values"].view(48, 3, 224, 224)
    f = O.vit_forward(sd, cfg, px); f = f / f.norm(dim=-1, keepdim=True)
    m = bh["video_masks"].float().unsqueeze(-1)
    v = (f.view(4, 12, -1) * m).sum(1) / m.sum(1); v = v / v.norm(dim=-1, keepdim=True)
    t = O.openclip_text_forward(sd, cfg, bh["input_ids"]); t = t / t.norm(dim=-1, keepdim=True)
    model = get_application_model_for_evaluation("clip4clip", d)
    model.eval()
    with torch.no_grad():
        out = model(dict(batch))
    assert max_err(out["video_embeds"], v) < 1e-2 and max_err(out["text_embeds"], t) < 1e-2
    res = get_application_evaluator("clip4clip", ds, user_defined_parameters={}, eval_batch_size=2).evaluate(model)
    r = O.rank_of_match(out["text_embeds"].double().cpu(), out["video_embeds"].double().cpu())
    assert res[0][0] == "mean_recall" and abs(res[0][1] - sum(float((r < k).sum()) / 4 for k in (1, 5, 10)) / 3) < 1e-9
    pred = get_application_predictor("clip4clip", d, first_sequence="image")
    recs = pred.run([{"image": rows[0].split("\t")[1]}, {"image": rows[3].split("\t")[1]}])
    got = np.array([[float(x) for x in rec["video_feat"].split("\t")] for rec in recs])
    assert np.abs(got - v[[0, 3]].numpy()).max() < 1e-2
    pred = get_application_predictor("clip4clip", d, first_sequence="text")
    recs = pred.run([{"text": texts[1]}])
    assert np.abs(np.array([float(x) for x in recs[0]["text_feat"].split("\t")]) - t[1].numpy()).max() < 1e-2
