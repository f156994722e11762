"""Application base class -- same duck-typed contract as easynlp/appzoo/application.py:26-99."""
import torch


class Application(torch.nn.Module):
    def __init__(self):
        super().__init__()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, user_defined_parameters={}, **kwargs):
        return cls(pretrained_model_name_or_path, user_defined_parameters=user_defined_parameters, **kwargs)

    def forward(self, inputs):
        raise NotImplementedError

    def compute_loss(self, forward_outputs, label_ids, **kwargs):
        raise NotImplementedError
