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


def test_cabi_host_helpers_agree_with_the_oracle():
    """clipk_preprocess_kmax / clipk_preprocess_workspace are host-side helpers of the C ABI (no CUDA call): tap-table width per image ==
    the resampling filter length Pillow allocates (ksize of the longer-filter axis), workspace formula monotone and padded."""
    from easynlp_b200 import _lib as L
    lib = L.lib()
    rng = np.random.RandomState(0)
    shapes = [(224, 224), (224, 500), (500, 224), (37, 61), (4000, 3000), (64, 900), (225, 223), (1, 1), (223, 4000)]
    shapes += [(int(rng.randint(1, 3000)), int(rng.randint(1, 3000))) for _ in range(200)]
    for w, h in shapes:
        nw, nh = R.resized_shape(w, h, 224)
        want = 1
        if nw != w:
            want = max(want, R.coefficients(w, nw)[0]) if nw > 0 else want
        if nh != h:
            want = max(want, R.coefficients(h, nh)[0]) if nh > 0 else want
        assert lib.clipk_preprocess_kmax(w, h, 224) == want, (w, h, nw, nh)
    assert lib.clipk_preprocess_kmax(0, 5, 224) == 0
    a = lib.clipk_preprocess_workspace(4, 224, 9, 1000); b = lib.clipk_preprocess_workspace(8, 224, 9, 1000)
    assert 0 < a < b and lib.clipk_preprocess_workspace(4, 224, 9, 5000) - a == 4000 and lib.clipk_preprocess_workspace(0, 224, 9, 0) == 0
    assert a >= 4 * 2 * 224 * (8 + 9 * 4) + 1000
