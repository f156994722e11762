from .api import (get_application_model, get_application_model_for_evaluation, get_application_evaluator,  # noqa: F401
                  get_application_predictor, get_application_dataset)
