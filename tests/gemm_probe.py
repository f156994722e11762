"""GPU probe for the tcgen05 GEMM: runs every operand-major / epilogue variant, never stops at the first failure,
and prints an error map so a wrong descriptor field can be identified from one run.  (diagnostic tool, not a test)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import _lib as L
from easynlp_b200.ops import gemm


def run(name, M, N, K, a_mn, b_mn, splits=1, ints=False):
    g = torch.Generator(device="cuda").manual_seed(1)
    if ints:
        A = torch.randint(-3, 4, (M, K), generator=g, device="cuda").float()
        B = torch.randint(-3, 4, (N, K), generator=g, device="cuda").float()
    else:
        A = torch.randn(M, K, generator=g, device="cuda")
        B = torch.randn(N, K, generator=g, device="cuda")
    Ab, Bb = A.bfloat16(), B.bfloat16()
    ref = Ab.float() @ Bb.float().t()
    a_store = Ab.t().contiguous() if a_mn else Ab
    b_store = Bb.t().contiguous() if b_mn else Bb
    try:
        if splits > 1:
            out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
            gemm(a_store, b_store, out, a_mn_major=a_mn, b_mn_major=b_mn, mode=L.EPI_ATOMIC_ADD, splits=splits)
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.float32)
            gemm(a_store, b_store, out, a_mn_major=a_mn, b_mn_major=b_mn)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[{name}] EXCEPTION {e}")
        return False
    err = (out - ref).abs()
    tol = 1e-3 * K ** 0.5 + 1e-2
    bad = err > tol
    ok = not bool(bad.any())
    print(f"[{name}] M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} splits={splits}: max_err={err.max().item():.4g} "
          f"ref_max={ref.abs().max().item():.4g} bad={int(bad.sum())}/{bad.numel()} {'OK' if ok else 'FAIL'}")
    if not ok:
        # error map over 16x16 blocks (first 128x128 region)
        r = min(M, 128); c = min(N, 128)
        blk = bad[:r, :c].float()
        rb, cb = (r + 15) // 16, (c + 15) // 16
        print("   bad-fraction map (16x16 blocks, first 128x128):")
        for i in range(rb):
            print("   " + " ".join(f"{blk[i*16:(i+1)*16, j*16:(j+1)*16].mean().item():.1f}" for j in range(cb)))
        rows_bad = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols_bad = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows_bad, "first bad cols", cols_bad)
        print("   out[0,:8]", out[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())
        # is the result a permutation / partial-K sum?  test hypotheses
        for kk in (16, 32, 48, 64):
            if kk < K:
                part = Ab[:, :kk].float() @ Bb[:, :kk].float().t()
                print(f"   ||out - ref(K[:{kk}])||max = {(out - part).abs().max().item():.4g}")
    return ok


def main():
    print(torch.cuda.get_device_name(0))
    res = []
    res.append(run("nt-1tile-int", 128, 128, 64, 0, 0, ints=True))
    res.append(run("nt-k256-int", 128, 128, 256, 0, 0, ints=True))
    res.append(run("nt-bn256", 256, 256, 128, 0, 0))
    res.append(run("nt-tails", 300, 384, 192, 0, 0))
    res.append(run("nt-vit-qkv", 1576, 2304, 768, 0, 0))
    res.append(run("nn-1tile-int", 128, 128, 64, 0, 1, ints=True))
    res.append(run("nn-bn256", 256, 256, 128, 0, 1))
    res.append(run("nn-dgrad", 1576, 768, 3072, 0, 1))
    res.append(run("tn-1tile-int", 128, 128, 64, 1, 0, ints=True))
    res.append(run("tt-1tile-int", 128, 128, 64, 1, 1, ints=True))
    res.append(run("tt-bn256", 256, 256, 128, 1, 1))
    res.append(run("tt-wgrad-splitk", 768, 3072, 1576, 1, 1, splits=6))
    res.append(run("tt-wgrad-nosplit", 768, 768, 1576, 1, 1))
    print("SUMMARY", res)


if __name__ == "__main__":
    main()
