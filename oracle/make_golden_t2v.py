"""Generate tests/golden/t2v_tiny.npz by running the UNMODIFIED reference's Text2VideoRetrieval (appzoo/text2video_retrieval/model.py:38-120)
on the open_clip tiny checkpoint of openclip_tiny_fwd_bwd.npz.  Test infrastructure only.      python oracle/make_golden_t2v.py"""
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
from oracle.ref_loader import import_reference, reference_root  # noqa: E402


def main():
    import_reference()
    root = "/root/reference" if os.path.isdir("/root/reference/easynlp") else reference_root()
    m = types.ModuleType("easynlp.appzoo.text2video_retrieval")
    m.__path__ = [os.path.join(root, "easynlp/appzoo/text2video_retrieval")]
    sys.modules["easynlp.appzoo.text2video_retrieval"] = m
    from easynlp.appzoo.text2video_retrieval.model import Text2VideoRetrieval
    z = np.load(os.path.join(ROOT, "tests", "golden", "openclip_tiny_fwd_bwd.npz"))
    cfg = json.loads(bytes(z["cfg_json"]).decode())
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    g = torch.Generator().manual_seed(23)
    B, T = 5, 3
    pixels = torch.randn(B, T, 3, 64, 64, generator=g)
    masks = torch.tensor([[1, 1, 1], [1, 1, 0], [1, 0, 0], [1, 1, 1], [0, 1, 1]])
    ids = torch.from_numpy(z["ids"])[:B].clone()
    with tempfile.TemporaryDirectory() as d:
        json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
        torch.save({"open_clip." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
        app = Text2VideoRetrieval(d)
    app.train()
    out = app({"pixel_values": pixels.clone(), "video_masks": masks.clone(), "input_ids": ids.clone()})
    loss = app.compute_loss(out, [])["loss"]
    app.zero_grad(); loss.backward()
    grads = {n.replace("open_clip.", ""): p.grad.detach().clone() for n, p in app.named_parameters() if p.grad is not None}
    # oracle composition
    f = O.vit_forward(sd, cfg, pixels.view(B * T, 3, 64, 64)).view(B, T, -1)
    f = f / f.norm(dim=-1, keepdim=True)
    mk = masks.float().unsqueeze(-1)
    cnt = mk.sum(1); cnt[cnt == 0] = 1
    vfeat = (f * mk).sum(1) / cnt
    vemb = vfeat / vfeat.norm(dim=-1, keepdim=True)
    t = O.openclip_text_forward(sd, cfg, ids); temb = t / t.norm(dim=-1, keepdim=True)
    for a, b, nm in ((vemb, out["video_embeds"], "video_embeds"), (temb, out["text_embeds"], "text_embeds")):
        err = (a - b.detach()).abs().max().item(); print("oracle vs reference", nm, f"{err:.2e}"); assert err < 2e-5
    blob = {"pixels": pixels.numpy(), "video_masks": masks.numpy(), "ids": ids.numpy(), "out.loss": loss.detach().numpy(),
            "out.video_embeds": out["video_embeds"].detach().numpy(), "out.text_embeds": out["text_embeds"].detach().numpy(),
            "out.logits_per_text": out["logits_per_text"].detach().numpy()}
    for k in ("visual.proj", "visual.conv1.weight", "visual.transformer.resblocks.0.attn.in_proj_weight", "text_projection", "logit_scale",
              "transformer.resblocks.1.mlp.c_fc.weight", "token_embedding.weight", "visual.positional_embedding"):
        blob["g." + k] = grads[k].numpy()
    path = os.path.join(ROOT, "tests", "golden", "t2v_tiny.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
