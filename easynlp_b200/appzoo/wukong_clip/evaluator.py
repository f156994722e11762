"""WukongCLIPEvaluator -- drop-in for easynlp/appzoo/wukong_clip/evaluator.py:29-80: text->image recall@1/5/10 over the validation set
(`outputs, _ = model(batch)`, keys image_features / text_features), or with user_defined_parameters['cosine_similarity'] == 'True' the
mean matched-pair cosine similarity (:50-55, returns None like the reference).  The N x N matrix and the per-row sorts are replaced by
the blocked tcgen05 ranking kernel (clipk_retrieval_rank_tc)."""
import time

import torch

from ...core.evaluator import Evaluator
from ..clip.evaluator import recall_from_embeddings, summarize_recall


class WukongCLIPEvaluator(Evaluator):

    def __init__(self, valid_dataset, user_defined_parameters=None, **kwargs):
        super().__init__(valid_dataset, **kwargs)
        self.metrics = ["accuracy", "f1"]
        self.before = 0.0
        self.cal_sim = bool(user_defined_parameters) and user_defined_parameters.get("cosine_similarity") == "True"

    def evaluate(self, model):
        model.eval()
        total_spent_time = 0.0
        image_all, text_all = [], []
        for _step, batch in enumerate(self.valid_loader):
            t0 = time.time()
            with torch.no_grad():
                outputs, _ = model(batch)
            torch.cuda.synchronize()
            total_spent_time += time.time() - t0
            image_all.append(outputs["image_features"]); text_all.append(outputs["text_features"])
        image_embeds = torch.cat(image_all, dim=0); text_embeds = torch.cat(text_all, dim=0)
        query_len = text_embeds.shape[0]
        if self.cal_sim:
            similarity = (text_embeds * image_embeds).sum(1)          # the diagonal of the agreement matrix
            print("pair number: ", similarity.shape)
            print(similarity)
            print("averaged consine similarity ", similarity.mean())
            return
        return summarize_recall(recall_from_embeddings(text_embeds, image_embeds), query_len, total_spent_time)
