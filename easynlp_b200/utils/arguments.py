"""The slice of EasyNLP's global argument namespace the CLIP path reads (easynlp/utils/arguments.py:190-440,
easynlp/utils/global_vars.py:111-200): same flag names and defaults, one process-global namespace behind get_args()."""
import argparse
import os

_ARGS = None


def build_parser():
    p = argparse.ArgumentParser("easynlp_b200", allow_abbrev=False)
    a = p.add_argument
    a("--random_seed", type=int, default=1234)
    a("--mode", default="train", choices=["train", "evaluate", "predict"])
    a("--tables", default=None, type=str)
    a("--user_defined_parameters", default=None, type=str)
    a("--skip_first_line", action="store_true")
    a("--outputs", default=None, type=str)
    a("--app_name", default="clip", type=str)
    a("--distributed_backend", default="nccl", type=str)
    a("--sequence_length", default=16, type=int)
    a("--micro_batch_size", "--train_batch_size", default=2, type=int)
    a("--local_rank", default=None, type=int)
    a("--checkpoint_dir", "--checkpoint_path", default=None, type=str)
    a("--epoch_num", default=3.0, type=float)
    a("--save_checkpoint_steps", type=int, default=None)
    a("--save_all_checkpoints", action="store_true")
    a("--learning_rate", default=5e-5, type=float)
    a("--weight_decay", "--wd", default=1e-4, type=float)
    a("--max_grad_norm", "--mn", default=1.0, type=float)
    a("--optimizer_type", "--optimizer", default="AdamW", type=str)
    a("--warmup_proportion", "--wp", default=0.1, type=float)
    a("--logging_steps", default=100, type=int)
    a("--gradient_accumulation_steps", default=1, type=int)
    a("--resume_from_checkpoint", default=None, type=str)
    a("--input_schema", default=None, type=str)
    a("--first_sequence", default=None, type=str)
    a("--second_sequence", default=None, type=str)
    a("--label_name", default=None, type=str)
    a("--label_enumerate_values", default=None, type=str)
    a("--output_schema", default="", type=str)
    a("--append_cols", default=None, type=str)
    a("--worker_gpu", default=1, type=int)
    a("--use_amp", action="store_true")
    a("--data_threads", default=10, type=int)
    a("--pretrained_model_name_or_path", default=None, type=str)
    return p


def parse_args(argv=None):
    args, _ = build_parser().parse_known_args(argv)
    args.rank = int(os.environ.get("RANK", "0"))
    args.world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.local_rank is None:
        args.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.is_master_node = args.rank == 0
    args.n_gpu = args.world_size
    return args


def set_args(args):
    global _ARGS
    _ARGS = args
    return args


def get_args():
    global _ARGS
    if _ARGS is None:
        _ARGS = parse_args([])
    return _ARGS


# app-level keys routed into ret['app_parameters'] with their types (the role of _GLOBAL_APP_PARAMETER_NAMES, global_vars.py:170-200);
# `global_contrastive` is this path's own switch: InfoNCE over the all-gathered global batch instead of each rank's local batch
APP_PARAMETER_NAMES = {"global_contrastive": "bool", "feature_format": "str"}


def parse_user_defined_parameters(s):
    """'k=v k=v' -> {plain keys..., 'app_parameters': {typed app-level keys}}  (global_vars.py:170-200)."""
    ret, app = {}, {}
    if s:
        for ele in s.split():
            k, v = ele.split("=", 1)
            if k in APP_PARAMETER_NAMES:
                t = APP_PARAMETER_NAMES[k]
                app[k] = int(v) if t == "int" else float(v) if t == "float" else (v == "True") if t == "bool" else v
            else:
                ret[k] = v
    ret["app_parameters"] = app
    return ret
