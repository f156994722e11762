"""attention kernels: correctness against a torch fp32 reference + same-box A/B timing of the first-generation (CLIPK_ATTN_V1=1) and
the persistent warp-specialised kernels (diagnostic, not a test)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import ops

DEV = "cuda"


def ref(qkv, mask, B, L, H):
    q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + mask[:, None, None, :]
    p = s.softmax(-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * L, H * 64), torch.logsumexp(s, -1)


def run(B, L, H, masked, which):
    os.environ["CLIPK_ATTN_V1"] = "1" if which == 1 else "0"
    d = H * 64
    g = torch.Generator(device=DEV).manual_seed(5)
    qkv = (torch.randn(B * L, 3 * d, generator=g, device=DEV) * 1.5).bfloat16()
    mask = None
    if masked:
        lens = torch.randint(max(1, L // 4), L + 1, (B,), device=DEV, generator=g)
        mask = ((torch.arange(L, device=DEV)[None, :] >= lens[:, None]).float() * -10000.0).contiguous()
    ctx = torch.zeros(B * L, d, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, H, L, device=DEV)
    ops.attention_fwd(qkv, mask, ctx, lse, B, L, H)
    torch.cuda.synchronize()
    out = {}
    if B * L * L * H < 3e8:
        o, l = ref(qkv, mask, B, L, H)
        out["ctx_err"] = (ctx.float() - o).abs().max().item(); out["lse_err"] = (lse - l).abs().max().item()
    for _ in range(3):
        ops.attention_fwd(qkv, mask, ctx, lse, B, L, H)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    n = 10
    s.record()
    for _ in range(n):
        ops.attention_fwd(qkv, mask, ctx, lse, B, L, H)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    out["fwd_us"] = ms * 1e3; out["fwd_tflops"] = 4.0 * B * H * L * L * 64 / (ms * 1e-3) / 1e12
    return out


if __name__ == "__main__":
    shapes = [(2, 197, 12, False), (3, 77, 12, True), (2, 17, 2, False), (2, 16, 2, True), (1, 256, 2, False), (2, 128, 1, True), (5, 50, 3, True),
              (40, 197, 12, False), (256, 197, 12, False), (256, 77, 12, True), (2, 257, 3, False), (3, 270, 2, False), (64, 257, 16, False)]
    only = sys.argv[1:] and sys.argv[1]
    if len(sys.argv) > 2:
        shapes = [shapes[int(x)] for x in sys.argv[2].split(",")]
    for sh in shapes:
        for which in ((2,) if only == "v2" else (1, 2)):
            try:
                r = run(*sh, which)
            except Exception as ex:
                r = {"error": repr(ex)[:200]}
            print(sh, "v%d" % which, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
