"""Generate tests/golden/openclip_tiny_fwd_bwd.npz by running the UNMODIFIED reference's open_clip branch (appzoo/clip/model.py:56-63 ->
OPEN_CLIP, modelzoo/models/clip/modeling_openclip.py:255-383) in the build container.  Test infrastructure only.

    python oracle/make_golden_openclip.py

Pins inputs (ids with an EOT token = highest id at a different position per row), every weight, reference embeds / logits / loss and
every gradient; the oracle restatement (oracle.clip_oracle.openclip_forward) is checked against the reference before writing."""
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
    cfg = O.openclip_tiny_config()
    sd = O.openclip_init_state_dict(cfg, seed=17, scale_boost=2.0)
    g = torch.Generator().manual_seed(17)
    B, L = 6, cfg["context_length"]
    pixels = torch.randn(B, 3, 64, 64, generator=g)
    lens = torch.tensor([24, 9, 12, 3, 17, 5])
    ids = torch.randint(1, 500, (B, L), generator=g)
    ids[:, 0] = 510                                             # start-of-text
    ids = torch.where(torch.arange(L)[None, :] < lens[:, None], ids, torch.zeros_like(ids))
    ids[torch.arange(B), lens - 1] = 511                        # end-of-text = the highest id: argmax pooling position
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(cfg, f)
        torch.save({"open_clip." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
        with open(os.path.join(d, "vocab.txt"), "w") as f:
            f.write("[PAD]\n")
        app = R["CLIPApp"](d)
    assert app.model_type == "open_clip"
    app.train()
    out = app({"pixel_values": pixels.clone(), "input_ids": ids.clone()})
    loss = app.compute_loss(out, [])["loss"]
    app.zero_grad()
    loss.backward()
    grads = {n.replace("open_clip.", ""): p.grad.detach().clone() for n, p in app.named_parameters() if p.grad is not None}
    o = O.openclip_forward(sd, cfg, pixels, ids)
    for k in ("image_embeds", "text_embeds", "logits_per_text"):
        err = (o[k] - out[k].detach()).abs().max().item()
        print(f"oracle vs reference {k}: max err {err:.2e}")
        assert err < 2e-5 * max(1.0, out[k].abs().max().item()), k
    blob = {"cfg_json": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8), "pixels": pixels.numpy(), "ids": ids.numpy(), "out.loss": loss.detach().numpy()}
    for k in ("image_embeds", "text_embeds", "logits_per_text"):
        blob["out." + k] = out[k].detach().numpy()
    for k, v in sd.items():
        blob["w." + k] = v.numpy()
    for k, v in grads.items():
        blob["g." + k] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "openclip_tiny_fwd_bwd.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB;", len(grads), "gradient tensors")


if __name__ == "__main__":
    main()
