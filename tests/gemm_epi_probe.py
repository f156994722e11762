"""epilogue-variant probe for the K=768,N=768 GEMM (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import timeit
M, N, K = 50432, 768, 768
dev = "cuda"
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); bias = torch.randn(N, device=dev)
res = torch.randn(M, N, device=dev); of = torch.empty(M, N, device=dev); ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
fl = 2.0 * M * N * K
for name, f in (("bf16 out", lambda: ops.gemm(A, W, ob)), ("bf16 out + bias", lambda: ops.gemm(A, W, ob, bias=bias)),
                ("f32 out", lambda: ops.gemm(A, W, of)), ("f32 out + bias", lambda: ops.gemm(A, W, of, bias=bias)),
                ("bf16 out + residual", lambda: ops.gemm(A, W, ob, residual=res)),
                ("f32 out + residual", lambda: ops.gemm(A, W, of, residual=res)),
                ("f32 out + bias + residual", lambda: ops.gemm(A, W, of, bias=bias, residual=res)),
                ("f32 out + residual in place", lambda: ops.gemm(A, W, res, residual=res)),
                ("f32 out + bf16 shadow", lambda: ops.gemm(A, W, of, out2=ob))):
    t = timeit(f)
    print(f"{name:32s} {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TF/s")
