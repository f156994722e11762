"""Text2VideoRetrievalEvaluator -- drop-in for easynlp/appzoo/text2video_retrieval/evaluator.py:27-73: text->video recall@1/5/10 over the
validation set (`model(batch)` -> video_embeds / text_embeds); the N x N matrix and per-row sorts are replaced by clipk_retrieval_rank_tc."""
import time

import torch

from ...core.evaluator import Evaluator
from ..clip.evaluator import recall_from_embeddings, summarize_recall


class Text2VideoRetrievalEvaluator(Evaluator):

    def __init__(self, valid_dataset, **kwargs):
        kwargs.pop("user_defined_parameters", None)
        super().__init__(valid_dataset, **kwargs)
        self.metrics = ["accuracy", "f1"]
        self.before = 0.0

    def evaluate(self, model):
        model.eval()
        total_spent_time = 0.0
        video_all, text_all = [], []
        for _step, batch in enumerate(self.valid_loader):
            t0 = time.time()
            with torch.no_grad():
                outputs = model(batch)
            torch.cuda.synchronize()
            total_spent_time += time.time() - t0
            video_all.append(outputs["video_embeds"]); text_all.append(outputs["text_embeds"])
        video_embeds = torch.cat(video_all, dim=0); text_embeds = torch.cat(text_all, dim=0)
        query_len = text_embeds.shape[0]
        return summarize_recall(recall_from_embeddings(text_embeds, video_embeds), query_len, total_spent_time)
