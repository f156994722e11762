"""The image-preprocessing oracle (oracle/pil_resample.py) pinned against Pillow itself and against the vectors written by the
reference's _resize / _center_crop / _normalize (tests/golden/preprocess.npz) -- bit-exact."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import pil_resample as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("seed", range(4))
def test_resize_bit_exact_vs_pillow(seed):
    rng = np.random.RandomState(seed)
    for _ in range(6):
        w, h = int(rng.randint(8, 420)), int(rng.randint(8, 420))
        nw, nh = int(rng.randint(4, 300)), int(rng.randint(4, 300))
        kind = rng.randint(3)
        if kind == 0:
            arr = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)                        # noise: exercises the clamp on overshoot
        elif kind == 1:
            arr = (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)                  # saturated edges
        else:
            yy, xx = np.mgrid[0:h, 0:w]
            arr = np.stack([(xx * 255 // max(1, w - 1)), (yy * 255 // max(1, h - 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
        want = np.array(Image.fromarray(arr).resize((nw, nh), Image.BICUBIC))
        got = R.resize_bicubic(arr, nw, nh)
        assert got.shape == want.shape and np.array_equal(got, want), (w, h, nw, nh, kind)


def test_one_axis_and_identity():
    rng = np.random.RandomState(9)
    arr = rng.randint(0, 256, (50, 70, 3)).astype(np.uint8)
    for nw, nh in ((70, 20), (33, 50), (70, 50), (140, 50), (70, 200)):
        assert np.array_equal(R.resize_bicubic(arr, nw, nh), np.array(Image.fromarray(arr).resize((nw, nh), Image.BICUBIC)))


def test_pipeline_matches_reference_vectors():
    z = np.load(os.path.join(GOLD, "preprocess.npz"))
    n = len([k for k in z.files if k.startswith("in")])
    assert n >= 5
    for i in range(n):
        got = R.preprocess(z[f"in{i}"])
        assert got.dtype == np.float32 and got.shape == (3, 224, 224)
        assert np.array_equal(got, z[f"out{i}"]), i


def test_shape_rules():
    assert R.resized_shape(300, 200) == (336, 224) and R.resized_shape(200, 300) == (224, 336) and R.resized_shape(224, 500) == (224, 500)
    assert R.resized_shape(640, 481) == (298, 224) and R.crop_origin(298, 224) == (37, 0) and R.crop_origin(224, 337) == (0, 57)
