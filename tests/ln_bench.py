"""LayerNorm fwd / bwd at the step's shapes (diagnostic): effective HBM GB/s.  CLIPK_LN_BWD_PRE=0|1|2 selects the dx_add prefetch variant"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from easynlp_b200 import ops
from gemm_bench import timeit

dev = "cuda"
tag = os.environ.get("CLIPK_LN_BWD_PRE", "default")
for M in (50432, 19712):
    d = 768
    x = torch.randn(M, d, device=dev); add = torch.randn(M, d, device=dev).bfloat16(); xo = torch.empty_like(x)
    g = torch.randn(d, device=dev); b = torch.randn(d, device=dev)
    y = torch.empty(M, d, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    t = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-5, y_bf16=y, mean=mean, rstd=rstd, add=add, x_out=xo), iters=20)
    byt = M * d * (4 + 2 + 4 + 2)
    print(f"[{tag}] ln_fwd  M={M}: {t*1e3:7.1f} us  {byt/t/1e6:7.0f} GB/s", flush=True)
    dy = torch.randn(M, d, device=dev).bfloat16(); dxa = torch.randn(M, d, device=dev); dx = torch.empty_like(x); dxb = torch.empty_like(dy)
    dg = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev); dbias = torch.zeros(d, device=dev)
    t = timeit(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dx_add=dxa, dx_f32=dx, dx_bf16=dxb, dgamma=dg, dbeta=db, dbias=dbias), iters=20)
    byt = M * d * (2 + 4 + 4 + 4 + 2)
    print(f"[{tag}] ln_bwd  M={M}: {t*1e3:7.1f} us  {byt/t/1e6:7.0f} GB/s", flush=True)
