from .model import CLIPApp  # noqa: F401
