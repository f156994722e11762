from .model import WukongCLIP  # noqa: F401
from .evaluator import WukongCLIPEvaluator  # noqa: F401
from .predictor import WukongCLIPPredictor  # noqa: F401
from .data import WukongCLIPDataset, FullTokenizer  # noqa: F401
