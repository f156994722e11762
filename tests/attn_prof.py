"""attention fwd/bwd launches for an ncu capture (diagnostic).  usage: attn_prof.py [L] [masked]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
B, H = 256, 12
L = int(sys.argv[1]) if len(sys.argv) > 1 else 197
masked = len(sys.argv) > 2 and sys.argv[2] == "1"
d = H * 64
qkv = (torch.randn(B * L, 3 * d, device="cuda")).bfloat16()
mask = None
if masked:
    lens = torch.randint(8, L + 1, (B,), device="cuda")
    mask = ((torch.arange(L, device="cuda")[None, :] >= lens[:, None]).float() * -10000.0).contiguous()
ctx = torch.empty(B * L, d, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B * H * L, device="cuda")
dctx = torch.randn(B * L, d, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.attention_fwd(qkv, mask, ctx, lse, B, L, H)
    ops.attention_bwd(qkv, mask, ctx, lse, dctx, dqkv, B, L, H)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
s.record(); ops.attention_fwd(qkv, mask, ctx, lse, B, L, H); e.record(); ops.attention_bwd(qkv, mask, ctx, lse, dctx, dqkv, B, L, H); e2.record()
torch.cuda.synchronize()
print("fwd ms", s.elapsed_time(e), "bwd ms", e.elapsed_time(e2))
