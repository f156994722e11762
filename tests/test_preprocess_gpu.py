"""GPU image preprocessing (clipk_preprocess_images) against Pillow / the reference chain: BIT-EXACT (integer resample, IEEE
single-precision normalisation) -- vectors written by the reference's own functions, the oracle restatement, the host chain on random sizes."""
import base64
import io
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

from easynlp_b200.appzoo.clip import data as D  # noqa: E402
from easynlp_b200.image_pipeline import ImagePreprocessor, preprocess_images  # noqa: E402
from oracle import pil_resample as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_matches_reference_vectors_bit_exact():
    z = np.load(os.path.join(GOLD, "preprocess.npz"))
    n = len([k for k in z.files if k.startswith("in")])
    out = preprocess_images([z[f"in{i}"] for i in range(n)]).cpu().numpy()
    for i in range(n):
        assert np.array_equal(out[i], z[f"out{i}"]), i
    one = preprocess_images([Image.fromarray(z["in3"])]).cpu().numpy()          # a batch of one, PIL input
    assert np.array_equal(one[0], z["out3"])


@pytest.mark.parametrize("seed", range(3))
def test_random_sizes_bit_exact_vs_host_chain(seed):
    """upscale, downscale (long filters), square, short side already 224, extreme aspect ratios, saturated content"""
    rng = np.random.RandomState(seed)
    shapes = [(224, 224), (224, 500), (500, 224), (37, 61), (1200, 90), (64, 900), (1024, 768), (225, 223)]
    shapes += [(int(rng.randint(20, 1500)), int(rng.randint(20, 1500))) for _ in range(8)]
    imgs = []
    for j, (w, h) in enumerate(shapes):
        if j % 3 == 0:
            arr = (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)
        else:
            arr = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        imgs.append(arr)
    pre = ImagePreprocessor()
    out = pre(imgs, check=True).cpu().numpy()
    for j, arr in enumerate(imgs):
        want = D.preprocess_image(Image.fromarray(arr))[0].numpy()               # Pillow + numpy, as the reference does it
        assert np.array_equal(out[j], want), (j, arr.shape, np.abs(out[j] - want).max())
        if j < 6:
            assert np.array_equal(R.preprocess(arr), want)                        # and the oracle restatement agrees
    out2 = pre(imgs[:5]).cpu().numpy()                                            # staging buffers are reused
    assert np.array_equal(out2, out[:5])


def test_predictor_and_dataset_use_the_gpu_chain(tmp_path):
    from easynlp_b200 import _lib as L
    from oracle import clip_oracle as O
    from easynlp_b200.appzoo import get_application_dataset
    cfg = O.tiny_config()
    model_dir = tmp_path / "m"; model_dir.mkdir()
    (model_dir / "config.json").write_text(json.dumps(dict(cfg, model_type="chinese_clip")))
    (model_dir / "vocab.txt").write_text(open(os.path.join(GOLD, "tokenizer_vocab.txt"), encoding="utf-8").read(), encoding="utf-8")
    rng = np.random.RandomState(4)
    rows, arrs = [], []
    for i, t in enumerate(("the cat", "一只猫", "red bike", "a dog")):
        arr = rng.randint(0, 256, (80 + 30 * i, 130, 3)).astype(np.uint8); arrs.append(arr)
        img = Image.fromarray(arr) if i != 2 else Image.fromarray(arr).convert("L")        # one grey image: stays on the host chain
        buf = io.BytesIO(); img.save(buf, format="PNG")
        rows.append(t + "\t" + base64.urlsafe_b64encode(buf.getvalue()).decode())
    tsv = tmp_path / "d.tsv"; tsv.write_text("\n".join(rows) + "\n", encoding="utf-8")
    kw = dict(input_schema="text:str:1,image:str:1", first_sequence="text", second_sequence="image")
    host = get_application_dataset("clip", str(model_dir), str(tsv), 16, **kw)
    gpu = get_application_dataset("clip", str(model_dir), str(tsv), 16, user_defined_parameters={"app_parameters": {"gpu_preprocess": "True"}}, **kw)
    assert gpu.gpu_preprocess and not host.gpu_preprocess
    n0 = L.launch_count()
    bh = host.batch_fn([host[i] for i in range(4)]); bg = gpu.batch_fn([gpu[i] for i in range(4)])
    assert L.launch_count() - n0 == 3 and bg["pixel_values"].is_cuda and not bh["pixel_values"].is_cuda
    assert torch.equal(bg["pixel_values"].cpu(), bh["pixel_values"]) and torch.equal(bg["input_ids"], bh["input_ids"])
