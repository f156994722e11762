"""three K = 768 GEMM launches of the step for an ncu capture (diagnostic): fc1 + QuickGELU pair, fc1 dgrad x saved derivative + colsum, qkv"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
M = 50432
dev = "cuda"
A = torch.randn(M, 768, device=dev).bfloat16(); W = (torch.randn(3072, 768, device=dev) * 0.05).bfloat16(); bias = torch.randn(3072, device=dev)
g = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16); a = torch.empty_like(g)
Wt = (torch.randn(768, 3072, device=dev) * 0.05).bfloat16(); dz = torch.empty_like(g); cs = torch.zeros(3072, device=dev)
Wq = (torch.randn(2304, 768, device=dev) * 0.05).bfloat16(); bq = torch.randn(2304, device=dev); q = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
torch.cuda.synchronize()
for _ in range(2):
    ops.gemm(A, W, g, bias=bias, mode=1, out2=a)
    ops.gemm(A, Wt, dz, b_mn_major=1, mode=3, aux=g, colsum=cs)
    ops.gemm(A, Wq, q, bias=bq)
torch.cuda.synchronize()
print("done")
