"""Base class of the evaluators the Trainer drives (contract of easynlp/core/evaluator.py:19-34).

What the training loop relies on (easynlp/core/trainer.py:242,371-381):
  * `valid_loader`      -- batches of the validation set in dataset order, collated by the dataset's own `batch_fn`;
  * `best_valid_score`  -- the best primary score seen so far (starts at -inf; the Trainer saves a checkpoint when it is beaten);
  * `evaluate(model)`   -- returns `[(metric_name, value), ...]`, primary metric first.
"""
from typing import List, Tuple

from torch.utils.data import DataLoader

DEFAULT_EVAL_BATCH = 32


class Evaluator:
    def __init__(self, valid_dataset, eval_batch_size: int = DEFAULT_EVAL_BATCH, **unused):
        self.best_valid_score = float("-inf")
        self.valid_loader = self._make_loader(valid_dataset, int(eval_batch_size))

    @staticmethod
    def _make_loader(dataset, batch_size: int) -> DataLoader:
        # sequential order matters: retrieval metrics pair the i-th text with the i-th image
        return DataLoader(dataset, batch_size=batch_size, shuffle=False, collate_fn=dataset.batch_fn)

    def evaluate(self, model) -> List[Tuple[str, float]]:
        raise NotImplementedError(f"{type(self).__name__} must implement evaluate(model)")

    @property
    def eval_metrics(self):
        raise NotImplementedError(f"{type(self).__name__} does not define eval_metrics")
