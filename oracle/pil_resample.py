"""CPU restatement of the image side of the input pipeline: PIL's 8-bit bicubic resample + the CLIP resize / centre-crop / normalise rule.
TEST INFRASTRUCTURE ONLY (tests/ and bench legs) -- the product path is clipk_preprocess_images (easynlp_b200/csrc/preprocess.cu).

What it restates: `CLIPDataset` / `CLIPPredictor` call `_resize(image, 224, Image.BICUBIC)` -> `_center_crop(224)` -> `_normalize`
(/root/reference/easynlp/appzoo/clip/data.py:29-135,263-272).  `_resize` is `PIL.Image.resize`, i.e. Pillow's `ImagingResample`
(third-party dependency of the reference: `pillow`, unpinned in requirements.txt; 12.2.0 installed here; src/libImaging/Resample.c).
Its published algorithm for 8-bit images: separable two-pass convolution, horizontal pass first, each pass
  * per output index: centre = (i + 0.5) * scale, support = 2 * max(scale, 1) (bicubic), taps [xmin, xmax) = round(centre -/+ support)
    clipped to the image, weights = Keys' cubic (a = -0.5) of (x + 0.5 - centre) / max(scale, 1) normalised to sum 1 in double precision;
  * weights rounded to fixed point with 22 fractional bits (round half away from zero);
  * accumulate in int32 starting from 2^21, arithmetic shift right by 22, clamp to [0, 255] -- the intermediate image is uint8.
Pinned against Pillow itself (tests/test_preprocess_oracle.py: bit-exact on random images and sizes) and against the vectors written by
the reference's own functions (tests/golden/preprocess.npz)."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int):
    """-> (ksize, bounds [out, 2] = (first tap, tap count), fixed-point weights [out, ksize] int32)"""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """one resampling pass of a uint8 [H, W, C] image along `axis` (0 = vertical, 1 = horizontal)"""
    in_size = img.shape[axis]
    _, bounds, kk = coefficients(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for i in range(out_size):
        x0, n = bounds[i]
        acc = np.tensordot(kk[i, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """PIL.Image.resize((new_w, new_h), BICUBIC) of a uint8 [H, W, C] array: horizontal pass, then vertical pass"""
    h, w = img.shape[:2]
    out = img
    if new_w != w:
        out = _pass(out, new_w, 1)
    if new_h != h:
        out = _pass(out, new_h, 0)
    return out


def resized_shape(w: int, h: int, size: int = 224):
    """data.py:54-74: shorter side -> size, the longer one int(size * long / short); unchanged when the short side already matches"""
    short, long = (w, h) if w <= h else (h, w)
    if short == size:
        return w, h
    new_long = int(size * long / short)
    return (size, new_long) if w <= h else (new_long, size)


def crop_origin(w: int, h: int, size: int = 224):
    """data.py:44-52: (left, top) of the centre crop"""
    return int((w - size + 1) * 0.5), int((h - size + 1) * 0.5)


def preprocess(img: np.ndarray, size: int = 224) -> np.ndarray:
    """uint8 RGB [H, W, 3] -> float32 [3, size, size]: resize, centre crop, /255, (x - mean) / std"""
    h, w = img.shape[:2]
    nw, nh = resized_shape(w, h, size)
    r = resize_bicubic(img, nw, nh)
    left, top = crop_origin(nw, nh, size)
    c = r[top:top + size, left:left + size]
    x = c.astype(np.float32) / 255.0
    x = x.transpose(2, 0, 1)
    return (x - CLIP_MEAN[:, None, None]) / CLIP_STD[:, None, None]
