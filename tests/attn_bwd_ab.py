"""attention backward: correctness against torch autograd (fp32) + same-box A/B timing v1 (CLIPK_ATTN_V1=1) vs v2 (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops
DEV = "cuda"


def run(B, L, H, which, check):
    os.environ["CLIPK_ATTN_V1"] = "1" if which == 1 else "0"
    d = H * 64
    g = torch.Generator(device=DEV).manual_seed(5)
    qkv = (torch.randn(B * L, 3 * d, generator=g, device=DEV) * 1.5).bfloat16()
    dctx = torch.randn(B * L, d, generator=g, device=DEV).bfloat16()
    ctx = torch.zeros(B * L, d, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, H, L, device=DEV)
    ops.attention_fwd(qkv, None, ctx, lse, B, L, H)
    dqkv = torch.zeros(B * L, 3 * d, device=DEV, dtype=torch.bfloat16); dbias = torch.zeros(3 * d, device=DEV)
    ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H, dqkv_colsum=dbias)
    torch.cuda.synchronize()
    out = {}
    if check:
        qf = qkv.float().requires_grad_(True)
        q, k, v = qf.view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
        o = ((q @ k.transpose(-1, -2) / 8.0).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * L, d)
        o.backward(dctx.float())
        ref = qf.grad
        sc = ref.abs().max().item()
        for nm, a in (("dq", 0), ("dk", 1), ("dv", 2)):
            out[nm + "_err/scale"] = (dqkv.float()[:, a * d:(a + 1) * d] - ref[:, a * d:(a + 1) * d]).abs().max().item() / sc
        out["dbias_err/scale"] = (dbias - ref.sum(0)).abs().max().item() / ref.sum(0).abs().max().item()
    for _ in range(3):
        ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H, dqkv_colsum=dbias)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    n = 10
    s.record()
    for _ in range(n):
        ops.attention_bwd(qkv, None, ctx, lse, dctx, dqkv, B, L, H, dqkv_colsum=dbias)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    out["bwd_us"] = ms * 1e3; out["bwd_tflops"] = 10.0 * B * H * L * L * 64 / (ms * 1e-3) / 1e12
    return out


if __name__ == "__main__":
    shapes = [(2, 197, 12, True), (1, 256, 2, True), (3, 130, 1, True), (2, 208, 3, True), (2, 209, 2, True), (40, 197, 12, True), (256, 197, 12, False)]
    only = sys.argv[1:] and sys.argv[1]
    for sh in shapes:
        for which in ((2,) if only == "v2" else (1, 2)):
            try:
                r = run(*sh[:3], which, sh[3])
            except Exception as ex:
                r = {"error": repr(ex)[:200]}
            print(sh[:3], "v%d" % which, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
