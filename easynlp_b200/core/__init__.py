from .trainer import Trainer  # noqa: F401
from .evaluator import Evaluator  # noqa: F401
