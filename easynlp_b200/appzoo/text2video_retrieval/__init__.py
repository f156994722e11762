from .model import Text2VideoRetrieval  # noqa: F401
