"""ctypes binding of the clipk C ABI (include/clipk.h).  The product path has NO fallback: if the shared
library is missing or a call fails, an exception is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libclipk.so")

EPI_LINEAR, EPI_QUICK_GELU, EPI_ERF_GELU, EPI_MUL_AUX, EPI_RANK_COUNT, EPI_ATOMIC_ADD = range(6)
BF16, F32 = 0, 1


class ClipkError(RuntimeError):
    pass


class Dropout(C.Structure):
    _fields_ = [("p", C.c_float), ("seed", C.c_ulonglong), ("dev_offset", C.c_void_p), ("site", C.c_uint)]


class Epilogue(C.Structure):
    _fields_ = [("mode", C.c_int), ("out_dtype", C.c_int), ("out", C.c_void_p), ("ldo", C.c_int),
                ("out2", C.c_void_p), ("ldo2", C.c_int), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("ldr", C.c_int), ("aux", C.c_void_p), ("ldaux", C.c_int), ("alpha", C.c_float), ("colsum", C.c_void_p),
                ("label_offset", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClipkError(f"{LIB_PATH} not found: build it with `python -m easynlp_b200.build` "
                             "(there is no CPU / PyTorch fallback on the product path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.clipk_last_error.restype = C.c_char_p
        _lib.clipk_launch_count.restype = C.c_int64
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    ll = C.c_longlong
    L.clipk_gemm_bf16.argtypes = [vp, i, i, vp, i, i, i, i, i, C.POINTER(Epilogue), i, vp]
    dp = C.POINTER(Dropout)
    L.clipk_attention_fwd.argtypes = [vp, vp, vp, vp, i, i, i, i, dp, vp]
    L.clipk_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, dp, vp]
    L.clipk_dropout_mask.argtypes = [vp, i, i, dp, vp]
    L.clipk_layernorm_fwd.argtypes = [vp, ll, vp, ll, vp, vp, vp, f, vp, vp, vp, vp, i, i, dp, i, vp]
    L.clipk_layernorm_bwd.argtypes = [vp, i, vp, vp, ll, vp, vp, vp, vp, vp, ll, vp, vp, vp, vp, i, i, dp, i, vp]
    L.clipk_colsum.argtypes = [vp, i, ll, vp, i, i, vp]
    L.clipk_im2col_patches.argtypes = [vp, vp, i, i, i, i, vp]
    L.clipk_vit_assemble.argtypes = [vp, vp, vp, vp, i, i, i, vp]
    L.clipk_vit_assemble_bwd.argtypes = [vp, vp, i, i, i, vp]
    L.clipk_bert_embed.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
    L.clipk_bert_embed_bwd.argtypes = [vp, vp, vp, i, i, i, vp]
    L.clipk_l2norm_fwd.argtypes = [vp, vp, vp, i, i, vp]
    L.clipk_l2norm_bwd.argtypes = [vp, vp, vp, vp, vp, i, i, vp]
    L.clipk_cast_bf16.argtypes = [vp, vp, ll, vp]
    L.clipk_axpy.argtypes = [vp, vp, f, ll, vp]
    L.clipk_retrieval_rank_tc_workspace.argtypes = [i, i, i]
    L.clipk_retrieval_rank_tc_workspace.restype = C.c_size_t
    L.clipk_retrieval_rank_tc.argtypes = [vp, vp, i, vp, i, i, i, vp, C.c_size_t, vp]
    L.clipk_split_bf16x3.argtypes = [vp, vp, i, i, i, ll, vp]
    L.clipk_ce_rows_fwd.argtypes = [vp, ll, vp, i, vp, vp, i, i, vp]
    L.clipk_ce_rows_bwd.argtypes = [vp, ll, vp, vp, i, f, vp, ll, vp, i, i, vp]
    L.clipk_ce_strip_fwd.argtypes = [vp, vp, vp, i, vp, ll, i, vp, vp, i, i, i, vp]
    L.clipk_ce_strip_bwd.argtypes = [vp, vp, vp, vp, i, f, i, vp, i, vp, i, i, i, vp]
    L.clipk_reduce_sum.argtypes = [vp, i, f, vp, i, vp]
    L.clipk_retrieval_rank.argtypes = [vp, vp, i, vp, i, i, i, vp]
    L.clipk_grad_norm.argtypes = [vp, ll, f, vp, i, vp, vp]
    L.clipk_adamw_step.argtypes = [vp, vp, vp, vp, vp, ll, f, f, f, f, f, i, vp, vp, vp]
    L.clipk_adam_schedule.argtypes = [vp, vp, f, i, i, f, f, vp]
    L.clipk_counter_add.argtypes = [vp, i, vp]
    L.clipk_peer_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.clipk_peer_free.argtypes = [vp]
    L.clipk_peer_export.argtypes = [vp, C.c_char_p]
    L.clipk_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.clipk_peer_close.argtypes = [vp]
    L.clipk_l2norm_allgather.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    L.clipk_peer_signal.argtypes = [vp, i, i, i, C.c_uint, vp]
    L.clipk_peer_wait.argtypes = [vp, i, i, C.c_uint, vp]
    L.clipk_peer_reduce_rows.argtypes = [vp, i, i, vp, i, i, i, vp]
    L.clipk_attention_causal_fwd.argtypes = [vp, vp, vp, i, i, i, i, vp]
    L.clipk_attention_causal_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
    L.clipk_argmax_rows.argtypes = [vp, vp, i, i, vp]
    L.clipk_find_token_rows.argtypes = [vp, ll, vp, vp, i, i, vp]
    L.clipk_gather_rows_bf16.argtypes = [vp, vp, vp, i, i, i, vp]
    L.clipk_scatter_rows_f32.argtypes = [vp, vp, vp, i, i, i, vp]
    L.clipk_frame_pool_fwd.argtypes = [vp, vp, vp, i, i, i, vp]
    L.clipk_frame_pool_bwd.argtypes = [vp, vp, vp, i, i, i, vp]
    L.clipk_preprocess_kmax.argtypes = [i, i, i]
    L.clipk_preprocess_workspace.argtypes = [i, i, i, ll]
    L.clipk_preprocess_workspace.restype = C.c_size_t
    L.clipk_preprocess_images.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, C.c_size_t, ll, vp]
    L.clipk_preprocess_status.argtypes = [vp, i, i, i, vp]
    L.clipk_wp_create.argtypes = [C.c_char_p, C.c_char_p, i]
    L.clipk_wp_create.restype = vp
    L.clipk_wp_destroy.argtypes = [vp]
    L.clipk_wp_destroy.restype = None
    L.clipk_wp_encode.argtypes = [vp, C.c_char_p, i, vp, vp]
    L.clipk_wp_encode_batch.argtypes = [vp, C.POINTER(C.c_char_p), i, i, vp, vp, vp, i]
    L.clipk_position_ids.argtypes = [vp, vp, i, i, i, vp]
    L.clipk_embed_gather.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    L.clipk_embed_gather_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    L.clipk_tanh_fwd.argtypes = [vp, vp, vp, ll, vp]
    L.clipk_tanh_bwd.argtypes = [vp, vp, vp, vp, ll, vp]


def check(rc: int, what: str = ""):
    if rc != 0:
        raise ClipkError(f"{what} failed ({rc}): {lib().clipk_last_error().decode()}")


def launch_count() -> int:
    return int(lib().clipk_launch_count())
