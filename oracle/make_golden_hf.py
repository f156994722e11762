"""Generate tests/golden/hf_tiny_fwd_bwd.npz by running the UNMODIFIED reference's huggingface_clip branch (appzoo/clip/model.py:73-104,
128-144: RobertaModel + CLIPVisionModel + biased projections) in the build container.  Test infrastructure only.

    python oracle/make_golden_hf.py          (needs /root/reference or oracle/_ref; CPU only)

Pins: inputs (pixels, ids with padding, token_type_ids, attention_mask), every weight, the reference's embeds / logits / loss, and every
reference gradient (the frozen image tower has none).  The oracle restatement (oracle.clip_oracle.hf_clip_forward) is checked against the
reference here, before the fixture is written."""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import clip_oracle as O  # noqa: E402
from oracle.ref_loader import import_reference  # noqa: E402


def main():
    R = import_reference()
    raw = O.hf_tiny_config()
    sd = O.hf_init_state_dict(raw, seed=11, scale_boost=2.0)
    g = torch.Generator().manual_seed(11)
    B, L = 6, 16
    pixels = torch.randn(B, 3, 64, 64, generator=g)
    lens = torch.tensor([16, 9, 12, 2, 16, 5])
    ids = torch.randint(1, 512, (B, L), generator=g)
    ids[:, 0] = 101
    ids = torch.where(torch.arange(L)[None, :] < lens[:, None], ids, torch.zeros_like(ids))
    tt = torch.zeros_like(ids); tt[:, 6:] = 1; tt = tt * (ids != 0).long()
    am = (ids != 0).long()
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(raw, f)
        torch.save(sd, os.path.join(d, "pytorch_model.bin"))
        with open(os.path.join(d, "vocab.txt"), "w") as f:
            f.write("[PAD]\n[UNK]\n[CLS]\n[SEP]\n")
        app = R["CLIPApp"](d)
    assert app.model_type == "huggingface_clip"
    app.train()
    out = app({"pixel_values": pixels.clone(), "input_ids": ids.clone(), "token_type_ids": tt.clone(), "attention_mask": am.clone()})
    loss = app.compute_loss(out, [])["loss"]
    app.zero_grad()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in app.named_parameters() if p.grad is not None}
    assert not any(n.startswith("vision_encoder.") for n in grads), "the image tower must be frozen by .detach()"
    o = O.hf_clip_forward(sd, raw, pixels, ids, tt, am)
    for k in ("image_embeds", "text_embeds", "logits_per_text"):
        err = (o[k] - out[k].detach()).abs().max().item()
        print(f"oracle vs reference {k}: max err {err:.2e}")
        assert err < 2e-5 * max(1.0, out[k].abs().max().item()), k
    assert abs(O.clip_loss(o["logits_per_text"]).item() - loss.item()) < 1e-5
    blob = {"cfg_json": np.frombuffer(json.dumps(raw).encode(), dtype=np.uint8), "pixels": pixels.numpy(), "ids": ids.numpy(),
            "token_type_ids": tt.numpy(), "attention_mask": am.numpy(), "out.loss": loss.detach().numpy()}
    for k in ("image_embeds", "text_embeds", "logits_per_text"):
        blob["out." + k] = out[k].detach().numpy()
    for k, v in sd.items():
        blob["w." + k] = v.numpy()
    for k, v in grads.items():
        blob["g." + k] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "hf_tiny_fwd_bwd.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB;", len(grads), "gradient tensors")


if __name__ == "__main__":
    main()
