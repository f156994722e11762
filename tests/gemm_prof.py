"""three GEMM launches for an ncu capture (diagnostic): fwd qkv (bias, bf16 out), fwd out-proj (bias + fp32 residual), dgrad qkv"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
M = 50432
dev = "cuda"
A = torch.randn(M, 768, device=dev).bfloat16(); W = (torch.randn(2304, 768, device=dev) * 0.05).bfloat16(); bias = torch.randn(2304, device=dev)
out = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
W2 = (torch.randn(768, 768, device=dev) * 0.05).bfloat16(); b2 = torch.randn(768, device=dev); res = torch.randn(M, 768, device=dev); o2 = torch.empty(M, 768, device=dev)
dX = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
torch.cuda.synchronize()
ops.gemm(A, W, out, bias=bias)
ops.gemm(A, W2, o2, bias=b2, residual=res)
ops.gemm(out, W, dX, b_mn_major=1)
torch.cuda.synchronize()
print("done")
