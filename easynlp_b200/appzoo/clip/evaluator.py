"""CLIPEvaluator -- drop-in for easynlp/appzoo/clip/evaluator.py:27-72.

Same protocol (encode the whole validation set, text->image recall@1/5/10, return [("mean_recall", fraction)]) but the
N x N agreement matrix and the per-row torch.sort loop are replaced by clipk_retrieval_rank_tc: a tcgen05 GEMM over bf16 hi/lo
splits of the embeddings (fp32-level scores) whose epilogue counts, per query, the gallery items scoring above the match --
hit@K <=> rank < K, and the matrix is never materialised (the fp32 CUDA-core kernel clipk_retrieval_rank remains as
the exact-arithmetic cross-check, `exact=True`)."""
import time

import torch

from ...core.evaluator import Evaluator
from ... import ops


def recall_from_embeddings(text_embeds: torch.Tensor, image_embeds: torch.Tensor, ks=(1, 5, 10), exact: bool = False):
    n = text_embeds.shape[0]
    ranks = torch.empty(n, dtype=torch.int32, device=text_embeds.device)
    fn = ops.retrieval_rank if (exact or text_embeds.shape[1] % 8) else ops.retrieval_rank_tc
    fn(text_embeds.float().contiguous(), image_embeds.float().contiguous(), ranks)
    r = ranks.cpu()
    return {k: int((r < k).sum()) for k in ks}


def recall_sharded(text_embeds_local: torch.Tensor, image_embeds_local: torch.Tensor, ks=(1, 5, 10)):
    """multi-GPU form: queries sharded over the ranks, gallery all-gathered (easynlp_b200/retrieval.py)"""
    from ...retrieval import sharded_recall
    return sharded_recall(text_embeds_local, image_embeds_local, ks)


def summarize_recall(hits, query_len: int, seconds: float):
    """hit counts {1: n1, 5: n5, 10: n10} -> the evaluators' return value [("mean_recall", fraction)]; prints the three report lines in
    the format of the reference evaluators (appzoo/clip/evaluator.py:62-70, wukong_clip/evaluator.py:69-77, text2video_retrieval/
    evaluator.py:63-71) -- shared by CLIPEvaluator, WukongCLIPEvaluator and Text2VideoRetrievalEvaluator"""
    fractions = [hits[k] / query_len for k in (1, 5, 10)]
    mean_recall = sum(fractions) / 3.0
    print(" ".join(f"r{k}_num:{hits[k]}" for k in (1, 5, 10)), "query_num:" + str(query_len))
    print(" ".join(f"r{k}(%):{f * 100}" for k, f in zip((1, 5, 10), fractions)), "mean_recall(%):" + str(mean_recall * 100))
    print("Inference time = {:.2f}s, [{:.4f} ms / sample] ".format(seconds, seconds * 1000 / max(1, query_len)))
    return [("mean_recall", mean_recall)]


class CLIPEvaluator(Evaluator):

    def __init__(self, valid_dataset, **kwargs):
        super().__init__(valid_dataset, **kwargs)
        self.metrics = ["accuracy", "f1"]
        self.before = 0.0

    def evaluate(self, model):
        model.eval()
        total_spent_time = 0.0
        image_embeds_all, text_embeds_all = [], []
        for _step, batch in enumerate(self.valid_loader):
            t0 = time.time()
            with torch.no_grad():
                outputs = model(batch, feat=True)
            torch.cuda.synchronize()
            total_spent_time += time.time() - t0
            image_embeds_all.append(outputs["image_embeds"].clone())
            text_embeds_all.append(outputs["text_embeds"].clone())
        image_embeds = torch.cat(image_embeds_all, dim=0)
        text_embeds = torch.cat(text_embeds_all, dim=0)
        return summarize_recall(recall_from_embeddings(text_embeds, image_embeds), text_embeds.shape[0], total_spent_time)
