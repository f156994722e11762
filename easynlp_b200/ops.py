"""Thin Python wrappers over the clipk C ABI: argument checking + pointer extraction only.
torch is used for device memory and streams; no computation happens here."""
import ctypes as C

import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a, b, out, *, a_mn_major=0, b_mn_major=0, mode=L.EPI_LINEAR, bias=None, residual=None, out2=None, aux=None,
         alpha=1.0, splits=1):
    """out[M,N] = epilogue(op(a) @ op(b)^T); see include/clipk.h clipk_gemm_bf16."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn_major, b_mn_major)
    assert out.shape == (M, N), (out.shape, M, N)
    e = L.Epilogue()
    e.mode = mode
    e.out_dtype = L.F32 if out.dtype == torch.float32 else L.BF16
    assert out.dtype in (torch.float32, torch.bfloat16)
    e.out = out.data_ptr(); e.ldo = out.stride(0)
    if out2 is not None:
        assert out2.dtype == torch.bfloat16 and out2.shape == (M, N) and out2.stride(1) == 1
        e.out2 = out2.data_ptr(); e.ldo2 = out2.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
        e.bias = bias.data_ptr()
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N) and residual.stride(1) == 1
        e.residual = residual.data_ptr(); e.ldr = residual.stride(0)
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and aux.shape == (M, N) and aux.stride(1) == 1
        e.aux = aux.data_ptr(); e.ldaux = aux.stride(0)
    e.alpha = alpha
    L.check(L.lib().clipk_gemm_bf16(_ptr(a), a.stride(0), int(a_mn_major), _ptr(b), b.stride(0), int(b_mn_major),
                                    M, N, K, C.byref(e), int(splits), _stream()), "clipk_gemm_bf16")
    return out
