"""GPU micro-benchmark of the tcgen05 GEMM on the model's shapes (diagnostic tool, not a test / not the bench line)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import _lib as L
from easynlp_b200 import ops
from easynlp_b200.engine import _splits_for


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = "cuda"
    torch.manual_seed(0)
    rows = []
    for M in (50432, 19712):
        for name, N, K, kind in (("qkv", 2304, 768, "bf16"), ("out", 768, 768, "res"), ("fc1", 3072, 768, "gelu"), ("fc2", 768, 3072, "res")):
            A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16() * 0.05
            bias = torch.randn(N, device=dev)
            flops = 2.0 * M * N * K
            if kind == "bf16":
                out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                f = lambda: ops.gemm(A, W, out, bias=bias)
            elif kind == "res":
                out = torch.empty(M, N, device=dev); res = torch.randn(M, N, device=dev)
                f = lambda: ops.gemm(A, W, out, bias=bias, residual=res)
            else:
                out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); out2 = torch.empty_like(out)
                f = lambda: ops.gemm(A, W, out, bias=bias, mode=L.EPI_QUICK_GELU, out2=out2)
            t = timeit(f)
            tc = timeit(lambda: torch.matmul(A, W.t()))
            rows.append((f"fwd {name} M={M} N={N} K={K} [{kind}]", flops / t / 1e9, flops / tc / 1e9, t))
            # dgrad: dX[M,K] = dY[M,N] W[N,K]
            dY = torch.randn(M, N, device=dev).bfloat16()
            dX = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
            t = timeit(lambda: ops.gemm(dY, W, dX, b_mn_major=1))
            tc = timeit(lambda: torch.matmul(dY, W))
            rows.append((f"dgrad {name} M={M} N={K} K={N}", flops / t / 1e9, flops / tc / 1e9, t))
            # wgrad: dW[N,K] += dY^T A
            dW = torch.zeros(N, K, device=dev)
            sp = _splits_for(N, K, M)
            t = timeit(lambda: ops.gemm(dY, A, dW, a_mn_major=1, b_mn_major=1, mode=L.EPI_ATOMIC_ADD, splits=sp))
            tc = timeit(lambda: torch.matmul(dY.t(), A))
            rows.append((f"wgrad {name} M={N} N={K} K={M} splits={sp}", flops / t / 1e9, flops / tc / 1e9, t))
            del A, W, dY, dX, dW, out
    print(f"{'shape':60s} {'clipk TF/s':>11s} {'cuBLAS TF/s':>12s} {'ms':>8s}")
    for r in rows:
        print(f"{r[0]:60s} {r[1]:11.1f} {r[2]:12.1f} {r[3]:8.3f}")


if __name__ == "__main__":
    main()
