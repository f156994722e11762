"""per-role clock64 timeline of CTA 0 of the persistent attention backward (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
B, H, L = 256, 12, 197
d = H * 64
qkv = torch.randn(B * L, 3 * d, device="cuda").bfloat16()
ctx = torch.empty(B * L, d, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B * H * L, device="cuda")
dctx = torch.randn(B * L, d, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv); dbias = torch.zeros(3 * d, device="cuda")
ops.attention_fwd(qkv, None, ctx, lse, B, L, H)
for _ in range(2):
    ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H, dqkv_colsum=dbias)
dbg = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
os.environ["CLIPK_ATTN_DBG_PTR"] = str(dbg.data_ptr())
ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H, dqkv_colsum=dbias)
torch.cuda.synchronize()
del os.environ["CLIPK_ATTN_DBG_PTR"]
t = dbg.cpu().view(64, 16)
t0 = int(t[0, 0])
names = {0: "mma:sdp-enter", 1: "mma:sdp-loads-ok", 2: "mma:sdp-issue", 3: "mma:gr-enter", 4: "mma:pds-full", 5: "mma:gr-issue", 6: "mma:gr-issued",
         8: "sw:enter", 9: "sw:sdp-ready", 10: "sw:computed", 11: "sw:pds-empty", 12: "sw:stored", 13: "sw:stored-last"}
for i in range(8, 18):
    print("step", i, " ".join(f"{names[e]}={int(t[i, e]) - t0}" for e in sorted(names) if t[i, e] > 0))
