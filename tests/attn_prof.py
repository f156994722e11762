"""attention fwd/bwd launches for an ncu capture (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
B, L, H = 256, 197, 12
d = H * 64
qkv = (torch.randn(B * L, 3 * d, device="cuda")).bfloat16()
ctx = torch.empty(B * L, d, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B * H * L, device="cuda")
dctx = torch.randn(B * L, d, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.attention_fwd(qkv, None, ctx, lse, B, L, H)
    ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
s.record(); ops.attention_fwd(qkv, None, ctx, lse, B, L, H); e.record(); ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H); e2.record()
torch.cuda.synchronize()
print("fwd ms", s.elapsed_time(e), "bwd ms", e.elapsed_time(e2))
