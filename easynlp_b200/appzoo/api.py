"""AppZoo registry for the applications on the hot path -- same call signatures as easynlp/appzoo/api.py:281-468 (prefix match on
app_name in the reference's dictionary order: 'clip4clip' before 'clip'; NotImplementedError for unknown apps, extra kwargs tolerated).
'clip', 'wukong_clip' and 'clip4clip' (Text2VideoRetrieval) are registered with model, dataset, evaluator and predictor."""


def _clip_classes():
    from .clip.model import CLIPApp
    from .clip.evaluator import CLIPEvaluator
    from .clip.predictor import CLIPPredictor
    from .clip.data import CLIPDataset
    return CLIPApp, CLIPEvaluator, CLIPPredictor, CLIPDataset


def _wukong_classes():
    from .wukong_clip.model import WukongCLIP
    from .wukong_clip.evaluator import WukongCLIPEvaluator
    from .wukong_clip.predictor import WukongCLIPPredictor
    from .wukong_clip.data import WukongCLIPDataset
    return WukongCLIP, WukongCLIPEvaluator, WukongCLIPPredictor, WukongCLIPDataset


def _t2v_classes():
    from .text2video_retrieval.model import Text2VideoRetrieval
    from .text2video_retrieval.evaluator import Text2VideoRetrievalEvaluator
    from .text2video_retrieval.predictor import Text2VideoRetrievalPredictor
    from .text2video_retrieval.data import Text2VideoRetrievalDataset
    return Text2VideoRetrieval, Text2VideoRetrievalEvaluator, Text2VideoRetrievalPredictor, Text2VideoRetrievalDataset


def _classes(app_name):
    """(model, evaluator, predictor, dataset) classes of a fully registered application"""
    if app_name is not None and app_name.startswith("wukong_clip"):
        return _wukong_classes()
    if app_name is not None and app_name.startswith("clip4clip"):       # before 'clip': the reference's dictionary order (api.py:141,162)
        return _t2v_classes()
    if app_name is not None and app_name.startswith("clip"):
        return _clip_classes()
    raise NotImplementedError(f"application {app_name!r} is outside the B200 hot path (registered: 'clip', 'clip4clip', 'wukong_clip')")


def _model_cls(app_name):
    return _classes(app_name)[0]


def get_application_model(app_name, pretrained_model_name_or_path, user_defined_parameters=None, **kwargs):
    return _model_cls(app_name)(pretrained_model_name_or_path, user_defined_parameters=user_defined_parameters, **kwargs)


def get_application_model_for_evaluation(app_name, pretrained_model_name_or_path, user_defined_parameters=None, **kwargs):
    return _model_cls(app_name).from_pretrained(pretrained_model_name_or_path, user_defined_parameters=user_defined_parameters or {}, **kwargs)


def get_application_evaluator(app_name, valid_dataset, user_defined_parameters=None, **kwargs):
    return _classes(app_name)[1](valid_dataset=valid_dataset, user_defined_parameters=user_defined_parameters, **kwargs)


def get_application_predictor(app_name, model_dir, user_defined_parameters=None, **kwargs):
    cls = _classes(app_name)
    return cls[2](model_dir=model_dir, model_cls=cls[0], user_defined_parameters=user_defined_parameters, **kwargs)


def get_application_dataset(app_name, pretrained_model_name_or_path, data_file, max_seq_length, user_defined_parameters=None, **kwargs):
    return _classes(app_name)[3](pretrained_model_name_or_path=pretrained_model_name_or_path, data_file=data_file,
                                 max_seq_length=max_seq_length, user_defined_parameters=user_defined_parameters, **kwargs)
