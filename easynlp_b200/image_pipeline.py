"""Batched image preprocessing on the GPU: decoded RGB images -> the fp32 [n, 3, 224, 224] batch of the image tower.

Host side of clipk_preprocess_images (csrc/preprocess.cu): packs the decoded frames of a batch into ONE pinned blob (pixels + the
descriptor table), one host->device copy, three kernel launches.  Results are bit-identical to the per-image host chain
`_resize(224, BICUBIC) -> _center_crop(224) -> _normalize` of the reference (easynlp/appzoo/clip/data.py:29-135) for mode-"RGB" images;
other modes (palette, grey, alpha) are the caller's business (appzoo/clip/data.py keeps the host chain for them).  There is no CPU
fallback in here: without the CUDA library the call raises."""
import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib as L

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)
DESC = np.dtype([("src", "<i8"), ("w", "<i4"), ("h", "<i4"), ("tmp", "<i8")])        # = clipk_image_desc


def as_rgb_array(image) -> np.ndarray:
    """PIL image (mode RGB) or uint8 [H, W, 3] array -> contiguous uint8 [H, W, 3]"""
    if isinstance(image, np.ndarray):
        arr = image
    else:
        if getattr(image, "mode", None) != "RGB":
            raise ValueError(f"GPU preprocessing takes mode-RGB images, got mode {getattr(image, 'mode', None)!r}")
        arr = np.asarray(image)
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError(f"expected uint8 [H, W, 3], got {arr.dtype} {arr.shape}")
    return np.ascontiguousarray(arr)


class ImagePreprocessor:
    """reusable staging buffers for one device; call with a list of images"""

    def __init__(self, size: int = 224, mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("easynlp_b200.image_pipeline needs a CUDA device: there is no CPU fallback")
        self.size = int(size)
        self.mean = np.ascontiguousarray(mean, dtype=np.float32); self.std = np.ascontiguousarray(std, dtype=np.float32)
        self.dev = torch.device(device)
        self._host = None; self._dev = None; self._ws = None
        self._copied = None          # event after the last host->device copy: the pinned staging buffer is rewritten only once it fired

    def _grow(self, name, nbytes, pinned=False):
        t = getattr(self, name)
        if t is None or t.numel() < nbytes:
            cap = max(nbytes, 1 << 20)
            cap = 1 << (cap - 1).bit_length()
            t = torch.empty(cap, dtype=torch.uint8, pin_memory=True) if pinned else torch.empty(cap, dtype=torch.uint8, device=self.dev)
            setattr(self, name, t)
        return t

    def stage(self, images: List):
        """pack the decoded frames + their descriptor table into the pinned blob and start the host->device copy; -> launch plan"""
        lib = L.lib()
        n = len(images); S = self.size
        arrs = [as_rgb_array(im) for im in images]
        desc = np.zeros(n, dtype=DESC)
        desc_bytes = (n * DESC.itemsize + 255) // 256 * 256
        off, tmp, kmax, max_h = desc_bytes, 0, 1, 1
        for i, a in enumerate(arrs):
            h, w = a.shape[:2]
            desc[i] = (off - desc_bytes, w, h, tmp)
            off += (a.size + 15) // 16 * 16
            tmp += (h * S * 3 + 255) // 256 * 256
            kmax = max(kmax, lib.clipk_preprocess_kmax(w, h, S)); max_h = max(max_h, h)
        if self._copied is not None:
            self._copied.synchronize()
        host = self._grow("_host", off, pinned=True)
        hv = host.numpy()
        hv[:n * DESC.itemsize] = desc.view(np.uint8)
        for i, a in enumerate(arrs):
            o = desc_bytes + int(desc[i]["src"])
            hv[o:o + a.size] = a.reshape(-1)
        dev = self._grow("_dev", off)
        dev[:off].copy_(host[:off], non_blocking=True)
        self._copied = torch.cuda.Event(); self._copied.record()
        return {"n": n, "desc_bytes": desc_bytes, "scratch": tmp, "kmax": kmax, "max_h": max_h, "bytes": off,
                "algorithmic_bytes": sum(a.size for a in arrs) + 2 * sum(a.shape[0] for a in arrs) * S * 3 + n * S * S * 3 * 4}

    def run(self, plan, check: bool = False) -> torch.Tensor:
        """the three kernel launches over a staged batch (may be repeated: the staged blob is not modified)"""
        lib = L.lib()
        n = plan["n"]; S = self.size; kmax = plan["kmax"]
        dev = self._dev
        ws_bytes = lib.clipk_preprocess_workspace(n, S, kmax, plan["scratch"])
        ws = self._grow("_ws", ws_bytes)
        out = torch.empty(n, 3, S, S, dtype=torch.float32, device=self.dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(lib.clipk_preprocess_images(C.c_void_p(dev.data_ptr() + plan["desc_bytes"]), C.c_void_p(dev.data_ptr()), n, plan["max_h"], S, kmax,
                                            self.mean.ctypes.data_as(C.c_void_p), self.std.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr()),
                                            C.c_void_p(ws.data_ptr()), ws_bytes, plan["scratch"], stream), "preprocess_images")
        if check and lib.clipk_preprocess_status(C.c_void_p(ws.data_ptr()), n, S, kmax, stream) != 0:
            raise RuntimeError("clipk_preprocess_images: tap table overflow (kmax too small)")
        return out

    def __call__(self, images: List, check: bool = False) -> torch.Tensor:
        if len(images) == 0:
            return torch.empty(0, 3, self.size, self.size, dtype=torch.float32, device=self.dev)
        return self.run(self.stage(images), check)


_default = {}


def preprocess_images(images: List, size: int = 224, device="cuda") -> torch.Tensor:
    """module-level convenience: one cached ImagePreprocessor per (device, size)"""
    key = (str(device), int(size))
    if key not in _default:
        _default[key] = ImagePreprocessor(size=size, device=device)
    return _default[key](images)
