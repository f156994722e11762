"""Thin Python wrappers over the clipk C ABI: argument checking + pointer extraction only.
torch is used for device memory and streams; no computation happens here."""
import ctypes as C

import torch

from . import _lib as L
L_ = L


# optional launch trace: bench.py sets TRACE = [] to time every launch with CUDA events on the launching stream
TRACE = None


def _traced(label, flops=0.0, bytes_=0.0):
    class _Ctx:
        def __enter__(self_):
            if TRACE is not None:
                self_.s = torch.cuda.Event(enable_timing=True); self_.e = torch.cuda.Event(enable_timing=True)
                self_.s.record()
            return self_

        def __exit__(self_, *a):
            if TRACE is not None:
                self_.e.record()
                TRACE.append((label, flops, bytes_, self_.s, self_.e))
            return False
    return _Ctx()


def _op(fn):
    """when TRACE is a list, bracket the launch with CUDA events (label = op name) unless the op records a richer entry itself"""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        if TRACE is None or fn.__name__ in ("gemm", "attention_fwd", "attention_bwd"):
            return fn(*a, **k)
        s_ = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
        s_.record()
        r = fn(*a, **k)
        e_.record()
        TRACE.append((fn.__name__, 0.0, 0.0, s_, e_))
        return r
    return wrapped


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@_op
def gemm(a, b, out, *, a_mn_major=0, b_mn_major=0, mode=L.EPI_LINEAR, bias=None, residual=None, out2=None, aux=None,
         alpha=1.0, splits=1, colsum=None):
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
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.numel() == N and colsum.is_contiguous()
        e.colsum = colsum.data_ptr()
    with _traced(f"gemm|{M}x{N}x{K}|{int(a_mn_major)}{int(b_mn_major)}|m{mode}{'r' if residual is not None else ''}{'f' if out.dtype == torch.float32 else 'b'}", 2.0 * M * N * K):
        L.check(L.lib().clipk_gemm_bf16(_ptr(a), a.stride(0), int(a_mn_major), _ptr(b), b.stride(0), int(b_mn_major),
                                        M, N, K, C.byref(e), int(splits), _stream()), "clipk_gemm_bf16")
    return out


def make_dropout(p, seed, site, dev_offset=None):
    """clipk_dropout_t for one call site (None when p == 0).  dev_offset: optional uint32/int32 device tensor xor-ed into the key."""
    if not p or p <= 0.0:
        return None
    d = L_.Dropout()
    d.p = float(p); d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF; d.site = int(site)
    d.dev_offset = dev_offset.data_ptr() if dev_offset is not None else None
    return d


def _dp(d):
    return C.byref(d) if d is not None else None


def _f32(t):
    assert t is None or (t.dtype == torch.float32 and t.is_cuda), "expected a CUDA float32 tensor"
    return _ptr(t)


def _b16(t):
    assert t is None or (t.dtype == torch.bfloat16 and t.is_cuda), "expected a CUDA bfloat16 tensor"
    return _ptr(t)


@_op
def attention_fwd(qkv, key_mask, ctx, lse, B, L, H, drop=None, causal=False):
    d = H * 64
    assert qkv.shape == (B * L, 3 * d) and qkv.is_contiguous() and ctx.shape == (B * L, d) and ctx.is_contiguous()
    assert lse.numel() == B * H * L
    with _traced(f"attention_fwd|L{L}", 4.0 * B * H * L * L * 64):
        if causal:
            assert key_mask is None and drop is None
            L_.check(L_.lib().clipk_attention_causal_fwd(_b16(qkv), _b16(ctx), _f32(lse), B, L, H, d, _stream()), "attention_causal_fwd")
        else:
            L_.check(L_.lib().clipk_attention_fwd(_b16(qkv), _f32(key_mask), _b16(ctx), _f32(lse), B, L, H, d, _dp(drop), _stream()), "attention_fwd")


@_op
def attention_bwd(qkv, key_mask, ctx, lse, dctx, dqkv, B, L, H, drop=None, dqkv_colsum=None, causal=False):
    """dqkv_colsum: optional f32 [3d], += column sums of dqkv (the QKV projection's bias gradient), fused into the kernel"""
    d = H * 64
    assert dqkv.shape == (B * L, 3 * d) and dqkv.is_contiguous() and dctx.shape == (B * L, d) and dctx.is_contiguous()
    assert dqkv_colsum is None or (dqkv_colsum.numel() == 3 * d and dqkv_colsum.is_contiguous())
    with _traced(f"attention_bwd|L{L}", 10.0 * B * H * L * L * 64):
        if causal:
            assert key_mask is None and drop is None
            L_.check(L_.lib().clipk_attention_causal_bwd(_b16(qkv), _b16(ctx), _f32(lse), _b16(dctx), _b16(dqkv), _f32(dqkv_colsum), B, L, H, d, _stream()),
                     "attention_causal_bwd")
            return
        L_.check(L_.lib().clipk_attention_bwd(_b16(qkv), _f32(key_mask), _b16(ctx), _f32(lse), _b16(dctx), _b16(dqkv), _f32(dqkv_colsum),
                                              B, L, H, d, _dp(drop), _stream()), "attention_bwd")


@_op
def layernorm_fwd(x, gamma, beta, eps, y_bf16=None, y_f32=None, mean=None, rstd=None, rows=None, ldx=None, add=None, ldadd=None,
                  x_out=None, drop=None, drop_mode=0):
    d = gamma.numel()
    if rows is None:
        rows = x.numel() // d
    if ldx is None:
        ldx = d
    if add is not None and ldadd is None:
        ldadd = d
    L_.check(L_.lib().clipk_layernorm_fwd(_f32(x), ldx, _b16(add), ldadd or 0, _f32(x_out), _f32(gamma), _f32(beta), eps, _b16(y_bf16),
                                          _f32(y_f32), _f32(mean), _f32(rstd), rows, d, _dp(drop), int(drop_mode), _stream()), "layernorm_fwd")


@_op
def layernorm_bwd(dy, x, gamma, mean, rstd, *, dy_add=None, dx_add=None, dx_f32=None, dx_bf16=None, dgamma=None, dbeta=None,
                  dbias=None, rows=None, ldx=None, lddx=None, drop=None, drop_mode=0):
    d = gamma.numel()
    if rows is None:
        rows = mean.numel()
    ldx = d if ldx is None else ldx
    lddx = d if lddx is None else lddx
    is_f32 = int(dy.dtype == torch.float32)
    assert dy.dtype in (torch.float32, torch.bfloat16)
    L_.check(L_.lib().clipk_layernorm_bwd(_ptr(dy), is_f32, _f32(dy_add), _f32(x), ldx, _f32(gamma), _f32(mean), _f32(rstd),
                                          _f32(dx_add), _f32(dx_f32), lddx, _b16(dx_bf16), _f32(dgamma), _f32(dbeta),
                                          _f32(dbias), rows, d, _dp(drop), int(drop_mode), _stream()), "layernorm_bwd")


@_op
def colsum(x, out, rows, n, ldx=None):
    ldx = n if ldx is None else ldx
    L_.check(L_.lib().clipk_colsum(_ptr(x), int(x.dtype == torch.float32), ldx, _f32(out), rows, n, _stream()), "colsum")


@_op
def im2col_patches(pixels, patches, B, R, P):
    L_.check(L_.lib().clipk_im2col_patches(_f32(pixels), _b16(patches), B, R, P, patches.stride(0), _stream()), "im2col")


@_op
def vit_assemble(patch, cls, pos, x0, B, L, W):
    L_.check(L_.lib().clipk_vit_assemble(_f32(patch), _f32(cls), _f32(pos), _f32(x0), B, L, W, _stream()), "vit_assemble")


@_op
def vit_assemble_bwd(dx0, dpatch, B, L, W):
    L_.check(L_.lib().clipk_vit_assemble_bwd(_f32(dx0), _b16(dpatch), B, L, W, _stream()), "vit_assemble_bwd")


@_op
def bert_embed(ids, word, pos, type0, e, rows, L, H, vocab, key_mask=None):
    assert ids.dtype == torch.int64 and ids.is_cuda and ids.is_contiguous()
    L_.check(L_.lib().clipk_bert_embed(_ptr(ids), _f32(word), _f32(pos), _f32(type0), _f32(e), _f32(key_mask), rows, L, H, vocab,
                                       _stream()), "bert_embed")


@_op
def bert_embed_bwd(ids, de, dword, rows, H, vocab):
    L_.check(L_.lib().clipk_bert_embed_bwd(_ptr(ids), _f32(de), _f32(dword), rows, H, vocab, _stream()), "bert_embed_bwd")


def _i64(t):
    assert t is None or (t.dtype == torch.int64 and t.is_cuda and t.is_contiguous()), "expected a contiguous CUDA int64 tensor"
    return _ptr(t)


@_op
def position_ids(ids, pos_ids, pad_id):
    B, Lt = ids.shape
    assert pos_ids.dtype == torch.int32 and pos_ids.shape == (B, Lt) and pos_ids.is_contiguous()
    L_.check(L_.lib().clipk_position_ids(_i64(ids), _ptr(pos_ids), B, Lt, int(pad_id), _stream()), "position_ids")


@_op
def embed_gather(ids, pos_ids, type_ids, attn_mask, word, pos, type_, e, key_mask, pad_id):
    rows = ids.numel(); H = word.shape[1]
    L_.check(L_.lib().clipk_embed_gather(_i64(ids), _ptr(pos_ids), _i64(type_ids), _i64(attn_mask), _f32(word), _f32(pos), _f32(type_), _f32(e),
                                         _f32(key_mask), rows, H, word.shape[0], pos.shape[0], type_.shape[0], int(pad_id), _stream()), "embed_gather")


@_op
def embed_gather_bwd(ids, pos_ids, type_ids, de, dword, dpos, dtype_, pad_id):
    rows = ids.numel(); H = dword.shape[1]
    L_.check(L_.lib().clipk_embed_gather_bwd(_i64(ids), _ptr(pos_ids), _i64(type_ids), _f32(de), _f32(dword), _f32(dpos), _f32(dtype_), rows, H,
                                             dword.shape[0], dpos.shape[0], dtype_.shape[0], int(pad_id), _stream()), "embed_gather_bwd")


@_op
def argmax_rows(ids, idx):
    B, Lt = ids.shape
    assert idx.dtype == torch.int32 and idx.numel() == B
    L_.check(L_.lib().clipk_argmax_rows(_i64(ids), _ptr(idx), B, Lt, _stream()), "argmax_rows")


@_op
def find_token_rows(ids, token, idx, count=None):
    B, Lt = ids.shape
    assert idx.dtype == torch.int32 and idx.numel() == B and (count is None or (count.dtype == torch.int32 and count.numel() == B))
    L_.check(L_.lib().clipk_find_token_rows(_i64(ids), int(token), _ptr(idx), _ptr(count), B, Lt, _stream()), "find_token_rows")


@_op
def gather_rows_bf16(x, idx, out, B, Lt, W):
    L_.check(L_.lib().clipk_gather_rows_bf16(_b16(x), _ptr(idx), _b16(out), B, Lt, W, _stream()), "gather_rows_bf16")


@_op
def scatter_rows_f32(src, idx, dst, B, Lt, W):
    L_.check(L_.lib().clipk_scatter_rows_f32(_f32(src), _ptr(idx), _f32(dst), B, Lt, W, _stream()), "scatter_rows_f32")


@_op
def frame_pool_fwd(x, mask, out, B, T, E):
    L_.check(L_.lib().clipk_frame_pool_fwd(_f32(x), _i64(mask), _f32(out), B, T, E, _stream()), "frame_pool_fwd")


@_op
def frame_pool_bwd(dout, mask, dx, B, T, E):
    L_.check(L_.lib().clipk_frame_pool_bwd(_f32(dout), _i64(mask), _f32(dx), B, T, E, _stream()), "frame_pool_bwd")


@_op
def tanh_fwd(x, y, y_bf16=None):
    L_.check(L_.lib().clipk_tanh_fwd(_f32(x), _f32(y), _b16(y_bf16), x.numel(), _stream()), "tanh_fwd")


@_op
def tanh_bwd(dy, y, dx=None, dx_bf16=None):
    L_.check(L_.lib().clipk_tanh_bwd(_f32(dy), _f32(y), _f32(dx), _b16(dx_bf16), dy.numel(), _stream()), "tanh_bwd")


@_op
def l2norm_fwd(x, y, norm, rows, d):
    L_.check(L_.lib().clipk_l2norm_fwd(_f32(x), _f32(y), _f32(norm), rows, d, _stream()), "l2norm_fwd")


@_op
def l2norm_bwd(dy, y, norm, dx_f32, dx_bf16, rows, d):
    L_.check(L_.lib().clipk_l2norm_bwd(_f32(dy), _f32(y), _f32(norm), _f32(dx_f32), _b16(dx_bf16), rows, d, _stream()), "l2norm_bwd")


@_op
def cast_bf16(x, y):
    assert x.numel() == y.numel()
    L_.check(L_.lib().clipk_cast_bf16(_f32(x), _b16(y), x.numel(), _stream()), "cast_bf16")


@_op
def ce_strip_fwd(Q, K, logit_scale_log, label_offset, lse, loss_rows, S_out=None, lds=0, transpose_out=False):
    nq, E = Q.shape
    nk = K.shape[0]
    L_.check(L_.lib().clipk_ce_strip_fwd(_f32(Q), _f32(K), _f32(logit_scale_log), label_offset, _f32(S_out), lds,
                                         int(transpose_out), _f32(lse), _f32(loss_rows), nq, nk, E, _stream()), "ce_strip_fwd")


@_op
def ce_strip_bwd(own, streamed, logit_scale_log, lse, label_offset, coef, own_is_query, out, accumulate, dscale_log=None):
    n_own, E = own.shape
    L_.check(L_.lib().clipk_ce_strip_bwd(_f32(own), _f32(streamed), _f32(logit_scale_log), _f32(lse), label_offset, coef,
                                         int(own_is_query), _f32(out), int(accumulate), _f32(dscale_log), n_own,
                                         streamed.shape[0], E, _stream()), "ce_strip_bwd")


@_op
def split_bf16x3(x, out, pattern):
    """x f32 [rows, E] -> out bf16 [rows(+pad), 3E]: [hi|hi|lo] (pattern 0, query side) or [hi|lo|hi] (pattern 1, gallery side)"""
    rows, E = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.bfloat16 and out.shape[1] == 3 * E and out.shape[0] >= rows
    L_.check(L_.lib().clipk_split_bf16x3(_f32(x), _b16(out), rows, E, int(pattern), out.stride(0), _stream()), "split_bf16x3")


@_op
def ce_rows_fwd(S, logit_scale_log, label_offset, lse, loss_rows, nq, nk):
    assert S.dtype == torch.float32 and S.stride(1) == 1
    L_.check(L_.lib().clipk_ce_rows_fwd(_f32(S), S.stride(0), _f32(logit_scale_log), label_offset, _f32(lse), _f32(loss_rows), nq, nk,
                                        _stream()), "ce_rows_fwd")


@_op
def ce_rows_bwd(S, logit_scale_log, lse, label_offset, coef, dS, nq, nk, dscale_log=None):
    assert S.dtype == torch.float32 and S.stride(1) == 1 and dS.dtype == torch.bfloat16 and dS.stride(1) == 1
    L_.check(L_.lib().clipk_ce_rows_bwd(_f32(S), S.stride(0), _f32(logit_scale_log), _f32(lse), label_offset, coef, _b16(dS), dS.stride(0),
                                        _f32(dscale_log), nq, nk, _stream()), "ce_rows_bwd")


@_op
def reduce_sum(x, n, scale, out, accumulate=False):
    L_.check(L_.lib().clipk_reduce_sum(_f32(x), n, scale, _f32(out), int(accumulate), _stream()), "reduce_sum")


@_op
def grad_norm(g, n, max_norm, workspace, norm_and_coef):
    assert workspace.dtype == torch.float64
    L_.check(L_.lib().clipk_grad_norm(_f32(g), n, max_norm, _ptr(workspace), workspace.numel(), _f32(norm_and_coef), _stream()),
             "grad_norm")


@_op
def adamw_step(p, g, m, v, w_bf16, n, lr, weight_decay, step, clip_coef=None, beta1=0.9, beta2=0.999, eps=1e-6, dev_hyper=None):
    L_.check(L_.lib().clipk_adamw_step(_f32(p), _f32(g), _f32(m), _f32(v), _b16(w_bf16), n, lr, beta1, beta2, eps, weight_decay,
                                       step, _f32(clip_coef), _f32(dev_hyper), _stream()), "adamw_step")


@_op
def adam_schedule(step_dev, hyper_dev, base_lr, warmup_steps, t_total, beta1=0.9, beta2=0.999):
    assert step_dev.dtype == torch.int32 and hyper_dev.dtype == torch.float32 and hyper_dev.numel() >= 2
    L_.check(L_.lib().clipk_adam_schedule(_ptr(step_dev), _ptr(hyper_dev), base_lr, warmup_steps, t_total, beta1, beta2, _stream()),
             "adam_schedule")


@_op
def counter_add(counter, value=1):
    assert counter.dtype == torch.int32 and counter.is_cuda
    L_.check(L_.lib().clipk_counter_add(_ptr(counter), int(value), _stream()), "counter_add")


@_op
def axpy(x, y, alpha=1.0):
    assert x.numel() == y.numel() and x.is_contiguous() and y.is_contiguous()
    L_.check(L_.lib().clipk_axpy(_f32(x), _f32(y), float(alpha), x.numel(), _stream()), "axpy")


@_op
def retrieval_rank(Q, K, rank_out, label_offset=0):
    nq, E = Q.shape
    assert rank_out.dtype == torch.int32 and rank_out.numel() == nq and Q.is_contiguous() and K.is_contiguous()
    L_.check(L_.lib().clipk_retrieval_rank(_f32(Q), _f32(K), label_offset, _ptr(rank_out), nq, K.shape[0], E, _stream()), "retrieval_rank")


@_op
def retrieval_rank_tc(Q, K, rank_out, label_offset=0):
    """ranks on the tensor cores: hi/lo-split K = 3E GEMM with the compare-and-count epilogue (no N x N matrix); see clipk.h"""
    nq, E = Q.shape
    nk = K.shape[0]
    assert rank_out.dtype == torch.int32 and rank_out.numel() == nq and Q.is_contiguous() and K.is_contiguous()
    nbytes = L_.lib().clipk_retrieval_rank_tc_workspace(nq, nk, E)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=Q.device)
    with _traced(f"retrieval_rank_tc|{nq}x{nk}x{E}", 2.0 * nq * nk * 3 * E):
        L_.check(L_.lib().clipk_retrieval_rank_tc(_f32(Q), _f32(K), label_offset, _ptr(rank_out), nq, nk, E, _ptr(ws), nbytes, _stream()),
                 "retrieval_rank_tc")


@_op
def dropout_mask(out, rows, cols, drop):
    L_.check(L_.lib().clipk_dropout_mask(_f32(out), rows, cols, _dp(drop), _stream()), "dropout_mask")
