"""per-role clock64 timeline of CTA 0 of the persistent attention forward (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
B, H = 256, 12
L = int(sys.argv[1]) if len(sys.argv) > 1 else 197
d = H * 64
qkv = torch.randn(B * L, 3 * d, device="cuda").bfloat16()
ctx = torch.empty(B * L, d, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B * H * L, device="cuda")
for _ in range(2):
    ops.attention_fwd(qkv, None, ctx, lse, B, L, H)
dbg = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
os.environ["CLIPK_ATTN_DBG_PTR"] = str(dbg.data_ptr())
ops.attention_fwd(qkv, None, ctx, lse, B, L, H)
torch.cuda.synchronize()
del os.environ["CLIPK_ATTN_DBG_PTR"]
t = dbg.cpu().view(64, 16)
t0 = int(t[0, 0])
names = {0: "mma:pre-wait-buf", 1: "mma:S-issue", 2: "mma:PV-issue", 3: "mma:PV-issued", 9: "sm:P-done-last", 4: "sm:pre-wait-S", 5: "sm:S-ready", 6: "sm:max-done", 7: "sm:bar-done", 8: "sm:P-done",
         10: "ep:O-ready", 11: "ep:release", 12: "ep:done"}
for i in range(10, 20):
    print("tile", i, " ".join(f"{names[e]}={int(t[i, e]) - t0}" for e in sorted(names) if t[i, e] > 0))
