"""A/B of the bf16 GEMM epilogue store path on the K = 768 shapes of the step (diagnostic, not a test).
   run with CLIPK_GEMM_EPI16=1 for the 16-epilogue-warp variants, CLIPK_GEMM_NO_TMA_OUT=1 for the transpose-through-smem path"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import _lib as L
from easynlp_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import timeit


def main():
    dev = "cuda"
    torch.manual_seed(0)
    tag = ("epi16" if os.environ.get("CLIPK_GEMM_EPI16") == "1" else "epi8") + ("/old" if os.environ.get("CLIPK_GEMM_NO_TMA_OUT") == "1" else "")
    for M in (50432, 19712):
        for N, K, b_mn, mode in ((2304, 768, 0, 0), (768, 768, 0, 0), (768, 768, 1, 0), (3072, 768, 0, 1), (3072, 768, 0, 2), (3072, 768, 1, 3), (768, 3072, 0, 0)):
            A = torch.randn(M, K, device=dev).bfloat16()
            W = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).bfloat16() * 0.05
            bias = torch.randn(N, device=dev)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            kw = dict(b_mn_major=b_mn)
            if mode in (1, 2):
                kw.update(mode=mode, out2=torch.empty_like(out), bias=bias)
            elif mode == 3:
                kw.update(mode=3, aux=torch.randn(M, N, device=dev).bfloat16(), colsum=torch.zeros(N, device=dev))
            else:
                kw.update(bias=bias)
            t = timeit(lambda: ops.gemm(A, W, out, **kw), iters=20)
            print(f"[{tag:8s}] M={M} N={N} K={K} b_mn={b_mn} mode={mode}: {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
