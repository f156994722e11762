"""Retrieval / inference at scale (BASELINE configs[4]; SURVEY.md 8e "Retrieval: shard queries; replicate (all-gather) the gallery
embeddings; no further exchange; recall counts summed with one tiny all-reduce").

Replaces CLIPEvaluator's single-process N x N agreement matrix + per-row torch.sort (easynlp/appzoo/clip/evaluator.py:47-61) and the
predictor's one-batch-at-a-time encode loop (appzoo/clip/predictor.py:118-138) for corpora that do not fit one process:
  * encode_stream: forward-only encode of an iterable of host batches (no activations kept), embeddings accumulated on the device;
  * sharded_recall: every rank holds n_local (text, image) pairs; the image gallery is all-gathered (NCCL over NVLink), each rank ranks
    ITS queries against the whole gallery on the tensor cores (clipk_retrieval_rank_tc: rank-count GEMM epilogue, the N x N matrix never
    exists) with label offset rank * n_local, and the hit counts are summed with one all-reduce of three integers."""
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import distributed as D
from . import ops


def encode_stream(engine, batches: Iterable[dict], max_rows: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """batches yield {'pixel_values': [b,3,R,R] f32, 'input_ids': [b,L] i64, (token_type_ids, attention_mask)} on the host or the device;
    returns (text_embeds, image_embeds), fp32 [n, E] on the engine's device."""
    T, I = [], []
    n = 0
    for b in batches:
        pix = b["pixel_values"].to(engine.dev, non_blocking=True).float().contiguous()
        ids = b["input_ids"].to(engine.dev, non_blocking=True).long().contiguous()
        out = engine.encode(pix, ids, token_type_ids=b.get("token_type_ids"), attention_mask=b.get("attention_mask"))
        T.append(out["text_embeds"].clone()); I.append(out["image_embeds"].clone())
        n += ids.shape[0]
        if max_rows is not None and n >= max_rows:
            break
    return torch.cat(T), torch.cat(I)


def sharded_recall(text_local: torch.Tensor, image_local: torch.Tensor, ks=(1, 5, 10)) -> Tuple[Dict[int, int], int]:
    """text -> image recall over the GLOBAL corpus; every rank passes its own n_local pairs (same n_local on every rank).
    -> ({k: hits summed over all ranks}, total number of queries)"""
    w, r = D.world_size(), D.get_rank()
    n = text_local.shape[0]
    if image_local.shape[0] != n:
        raise ValueError("sharded_recall: each rank must hold as many images as texts (pairs)")
    q = text_local.float().contiguous()
    gallery = D.gather_rows(image_local.float().contiguous()) if w > 1 else image_local.float().contiguous()
    ranks = torch.empty(n, dtype=torch.int32, device=q.device)
    ops.retrieval_rank_tc(q, gallery, ranks, label_offset=r * n)
    hits = torch.stack([(ranks < k).sum() for k in ks]).to(torch.int64)
    if w > 1:
        torch.distributed.all_reduce(hits)
    h = hits.tolist()
    return {k: int(v) for k, v in zip(ks, h)}, w * n
