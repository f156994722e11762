from .arguments import get_args, parse_args, set_args, parse_user_defined_parameters  # noqa: F401
from .schedule import warmup_linear_lambda  # noqa: F401
