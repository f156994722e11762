"""Throughput of the GPU image preprocessing vs the host chain (diagnostic, not a test):  python tests/preprocess_bench.py [n] [w] [h]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
from PIL import Image

from easynlp_b200.appzoo.clip import data as D
from easynlp_b200.image_pipeline import ImagePreprocessor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
rng = np.random.RandomState(0)
imgs = [rng.randint(0, 256, (h, w, 3)).astype(np.uint8) for _ in range(n)]
pre = ImagePreprocessor()
for _ in range(3):
    out = pre(imgs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out = pre(imgs)
torch.cuda.synchronize()
e2e = (time.perf_counter() - t0) / 5
# kernels only: the three launches over the staged blob, CUDA events on the launching stream
plan = pre.stage(imgs)
for _ in range(3):
    pre.run(plan)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    pre.run(plan)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"GPU chain: {n / e2e:.0f} images/s end to end (host pack + H2D + 3 kernels), {e2e * 1e6 / n:.1f} us/image; kernels alone {ms:.3f} ms per batch of {n} "
      f"= {n / ms * 1e3:.0f} images/s, {plan['algorithmic_bytes'] / ms / 1e6:.0f} GB/s algorithmic ({plan['algorithmic_bytes'] / 1e6:.1f} MB)")
pil = [Image.fromarray(a) for a in imgs[:32]]
t0 = time.perf_counter()
for im in pil:
    D.preprocess_image(im)
host = (time.perf_counter() - t0) / len(pil)
print(f"host chain (Pillow + numpy, 1 thread): {1 / host:.0f} images/s, {host * 1e6:.0f} us/image")
