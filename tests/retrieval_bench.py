"""blocked retrieval on the tensor cores (diagnostic): ranks of N queries against an N-item gallery, no N x N matrix.
   prints queries/s and the GEMM rate (2 * N * N * 3E flops: three bf16 products per fp32-level dot)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from easynlp_b200 import ops
from gemm_bench import timeit

E = 512
for N in (8192, 65536, 262144):
    g = torch.Generator(device="cuda").manual_seed(N)
    Q = torch.nn.functional.normalize(torch.randn(N, E, device="cuda", generator=g), dim=-1)
    K = torch.nn.functional.normalize(torch.nn.functional.normalize(torch.randn(N, E, device="cuda", generator=g), dim=-1) + 0.25 * Q, dim=-1)
    ranks = torch.empty(N, dtype=torch.int32, device="cuda")
    t = timeit(lambda: ops.retrieval_rank_tc(Q, K, ranks), iters=3) / 1e3
    r = ranks.cpu()
    print(f"N={N}: {t*1e3:9.2f} ms  {N/t/1e6:8.2f} M queries/s  {2.0*N*N*3*E/t/1e12:7.1f} TF/s   R@1 {(r < 1).float().mean():.4f} R@10 {(r < 10).float().mean():.4f}", flush=True)
    if N <= 8192:
        ref = torch.empty_like(ranks)
        t2 = timeit(lambda: ops.retrieval_rank(Q, K, ref), iters=2) / 1e3
        d = (ref - ranks).abs()
        print(f"   fp32 CUDA-core kernel: {t2*1e3:9.2f} ms; ranks differing: {int((d > 0).sum())} of {N} (max |diff| {int(d.max())}: near-ties at the 1e-7 level)", flush=True)
