"""Trainer -- the hot loop of easynlp/core/trainer.py:39-677 for the CLIP application, same constructor and call
sequence (model(batch) -> compute_loss -> backward -> every gradient_accumulation_steps: clip, step, schedule, zero_grad;
loss.item() per micro-step; rank-0 logging, periodic evaluate + checkpoint), with the optimizer side replaced by ONE fused
multi-tensor clip+AdamW launch per decay group (reference: ~2k tiny kernels from the Python loop of optimizers.py:405-464)
and DDP's gradient mean replaced by a SUM all-reduce of the flat gradient (the loss is already global-batch normalised
when `global_contrastive=True`; see DESIGN.md "multi-GPU").

Checkpoint FILES match the reference (trainer.py:421-580): config.json, pytorch_model[.step_N].bin with the
`chinese_clip.`-prefixed keys, *.meta.bin {epoch, global_step, optimizer}, train_config.json, label_mapping.json, vocab.txt.
`meta['optimizer']` is NOT torch's per-parameter `optimizer.state_dict()`: the moments live in one flat buffer
(`layout: flat:easynlp_b200.params`, easynlp_b200/params.py).  A reference-format meta (per-parameter `state` / `param_groups`) is
converted on resume through the parameter order of `named_parameters()`; anything else raises instead of silently restarting Adam.

Counters follow the reference: `_global_step` counts OPTIMIZER steps (trainer.py:345-347), the accumulation boundary is the
per-epoch `(_step + 1) % gradient_accumulation_steps` (trainer.py:345), logging / checkpoint cadences are in optimizer steps.
"""
import warnings
import json
import math
import os
import shutil
import time

import torch
from torch.utils.data import DataLoader, RandomSampler
from torch.utils.data.distributed import DistributedSampler

from ..utils.arguments import get_args
from ..utils.schedule import warmup_linear_lambda
from .. import distributed as D


class Trainer(object):
    def __init__(self, model, train_dataset, evaluator=None, **kwargs):
        self.args = kwargs.get("args", None) or get_args()
        self._model = model
        self.model_module = model
        self.evaluator = evaluator
        self._global_step = 0
        self._current_epoch = 0
        self._start_epoch = 0
        self._user_defined_parameters = kwargs.get("user_defined_parameters", None)
        if not hasattr(model, "engine") or model.engine is None:
            raise TypeError("easynlp_b200.Trainer drives models backed by the clipk engine (easynlp_b200 CLIPApp)")
        self.engine = model.engine
        # one GPU: the whole step replays as ONE CUDA graph.  N > 1: eager launches with the per-layer gradient all-reduce overlapped with
        # the backward pass (DESIGN.md "multi-GPU"); capturing NCCL inside the graph is opt-in (use_cuda_graph=True)
        self.use_graph = bool(kwargs.get("use_cuda_graph", D.world_size() == 1))
        self._micro_step = 0          # forward/backward passes since the start (the dropout stream advances per pass)
        self.set_train_loader(train_dataset, self.args)
        self.set_model_and_optimizer(model, self.args)
        self.resume_from_ckpt(self.args)
        self._log = []

    # ------------------------------------------------------------------ setup
    def set_train_loader(self, train_dataset, args):
        if D.world_size() > 1:
            sampler = DistributedSampler(train_dataset)
        else:
            sampler = RandomSampler(train_dataset)
        workers = max(0, min(int(getattr(args, "data_threads", 0) or 0), os.cpu_count() or 1))
        self._train_loader = DataLoader(train_dataset, sampler=sampler, batch_size=args.micro_batch_size, collate_fn=train_dataset.batch_fn,
                                        num_workers=workers, drop_last=D.world_size() > 1)

    def set_model_and_optimizer(self, model, args):
        steps_per_epoch = len(self._train_loader)
        self._t_total = int(math.ceil(steps_per_epoch / args.gradient_accumulation_steps * args.epoch_num))
        self._warmup_steps = int(args.warmup_proportion * self._t_total)      # optimizers.py:529-533
        self._sched_step = 0
        if args.optimizer_type != "AdamW":
            raise NotImplementedError("the fused optimizer implements the reference's default AdamW (optimizers.py:381-470)")

    @property
    def learning_rate(self):
        return self.args.learning_rate * warmup_linear_lambda(self._sched_step, self._warmup_steps, self._t_total)

    def resume_from_ckpt(self, args):
        prefix = getattr(args, "resume_from_checkpoint", None)
        if not prefix:
            return
        meta = torch.load(prefix + ".meta.bin", map_location="cpu")
        sd = torch.load(prefix + ".bin", map_location="cpu")
        self.model_module.load_state_dict(sd)
        self._load_optimizer_state(meta.get("optimizer"))
        self._start_epoch = int(meta.get("epoch", 0))
        self._global_step = int(meta.get("global_step", 0)) + 1          # optimizer steps already taken (trainer.py:146)
        self._sched_step = self._global_step                               # unlike the reference (quirk A.4-6) the schedule survives a resume
        self._micro_step = self._global_step * max(1, args.gradient_accumulation_steps)
        self.engine.set_micro_step(self._micro_step)

    def _load_optimizer_state(self, opt):
        """flat layout written by save_checkpoint, or the reference's `optimizer.state_dict()` ({'state': {i: {step, exp_avg,
        exp_avg_sq}}, 'param_groups': [...]}, core/optimizers.py:405-464) mapped through the reference's parameter order
        (decay group first, then no-decay, each in named_parameters() order: optimizers.py:519-523).  Unknown layouts raise."""
        P = self.engine.params
        if not opt:
            warnings.warn("resume: the checkpoint meta holds no optimizer state -- Adam moments and the step counter restart from zero")
            return
        if "exp_avg" in opt and opt.get("layout", "").startswith("flat:"):
            if opt["exp_avg"].numel() != P.exp_avg.numel():
                raise ValueError("resume: flat optimizer state has %d elements, this model %d" % (opt["exp_avg"].numel(), P.exp_avg.numel()))
            P.exp_avg.copy_(opt["exp_avg"].to(P.exp_avg.device)); P.exp_avg_sq.copy_(opt["exp_avg_sq"].to(P.exp_avg.device))
            self.engine.set_step(int(opt.get("step", 0)))      # host AND device step counters (Adam bias correction, schedule)
            return
        if "state" in opt and "param_groups" in opt:
            order = [n for g in opt["param_groups"] for n in g.get("names", [])]
            if not order:      # torch state_dicts carry indices only: rebuild the reference's grouping order
                names = list(P.trainable_names())
                from ..params import uses_weight_decay
                order = [n for n in names if uses_weight_decay(n)] + [n for n in names if not uses_weight_decay(n)]
            step = 0
            for idx, name in enumerate(order):
                stt = opt["state"].get(idx)
                if stt is None:
                    continue
                name = name.replace("chinese_clip.", "")
                P.m(name).copy_(stt["exp_avg"].to(P.exp_avg.device).view_as(P.m(name)))
                P.v(name).copy_(stt["exp_avg_sq"].to(P.exp_avg.device).view_as(P.v(name)))
                step = max(step, int(stt.get("step", 0)))
            self.engine.set_step(step)
            return
        raise ValueError("resume: unrecognised optimizer state layout in the checkpoint meta (keys: %s)" % sorted(opt.keys()))

    # ------------------------------------------------------------------ loop
    def train(self):
        args = self.args
        t_start = time.time()
        ga = max(1, int(args.gradient_accumulation_steps))
        self.engine.zero_grad()
        start_global = self._global_step
        if start_global:                       # resume: recount from the start epoch while skipping (trainer.py:606-610)
            self._global_step = self._start_epoch * (len(self._train_loader) // ga)
        for epoch in range(self._start_epoch, int(math.ceil(args.epoch_num))):
            self._current_epoch = epoch
            if D.world_size() > 1:
                D.barrier()
                if hasattr(self._train_loader.sampler, "set_epoch"):
                    self._train_loader.sampler.set_epoch(epoch)
            self._model.train()
            for _step, batch in enumerate(self._train_loader):
                if self._global_step < start_global:
                    if (_step + 1) % ga == 0:
                        self._global_step += 1
                    continue
                loss_val = self.train_step(batch, _step)
                self.after_iter(_step, epoch, loss_val)
        self.after_train()
        print("Training Time: {}".format(time.time() - t_start))

    def _global_loss(self, loss_t):
        """the engine's 'loss' is this rank's SHARE of the global-batch InfoNCE in distributed-loss mode (sum over ranks = the loss);
        local-loss data parallelism logs the rank's own loss like the reference does."""
        if D.world_size() > 1 and getattr(self.model_module, "distributed_loss", False):
            t = loss_t.detach().clone()
            D.allreduce_sum_(t)
            return t.item()
        return loss_t.item()

    def train_step(self, batch, _step=None):
        """One micro-step on a collated batch (host or device tensors): forward, loss, backward and -- every
        gradient_accumulation_steps -- clip + AdamW + schedule + zero_grad; returns the loss as a Python float (the reference
        reads loss.item() every micro-step too, trainer.py:342).  Without gradient accumulation the whole step runs fused
        (ClipEngine.train_step: one CUDA graph on one GPU, eager + overlapped gradient all-reduce on several); otherwise through
        model(batch) / compute_loss / loss.backward()."""
        args = self.args
        ga = max(1, int(args.gradient_accumulation_steps))
        label_ids = batch.pop("label_ids", None)
        dist_loss = bool(getattr(self.model_module, "distributed_loss", False))
        self._micro_step += 1
        # video batches (Text2VideoRetrieval: [B, T, 3, R, R] frames + video_masks) go through the application's own forward
        if ga == 1 and batch.get("video_masks") is None:
            out = self.engine.train_step(batch["pixel_values"], batch["input_ids"], lr=args.learning_rate, weight_decay=args.weight_decay,
                                         max_grad_norm=args.max_grad_norm, warmup_steps=self._warmup_steps, t_total=self._t_total,
                                         distributed=dist_loss, use_graph=self.use_graph,
                                         token_type_ids=batch.get("token_type_ids"), attention_mask=batch.get("attention_mask"))
            self._sched_step += 1
            return self._global_loss(out["loss"])
        forward_outputs = self._model(batch)
        loss_dict = self.model_module.compute_loss(forward_outputs, label_ids)
        _loss = loss_dict["loss"] / ga
        if D.world_size() > 1 and not dist_loss:
            # local-loss data parallelism: DDP averages the ranks' gradients (trainer.py:103-108); the flat gradient is SUM-all-reduced
            # in optimizer_step, so each rank contributes 1/world of its local gradient
            _loss = _loss / D.world_size()
        _loss.backward()
        loss_val = self._global_loss(loss_dict["loss"])
        step_in_epoch = self._micro_step - 1 if _step is None else _step
        if (step_in_epoch + 1) % ga == 0:
            self.optimizer_step()
        return loss_val

    def optimizer_step(self):
        if D.world_size() > 1:
            self.engine.allreduce_grads()
        self.engine.optimizer_step(lr=self.learning_rate, weight_decay=self.args.weight_decay, max_grad_norm=self.args.max_grad_norm)
        self._sched_step += 1
        self.engine.zero_grad()

    def after_iter(self, _step, _epoch, loss_val):
        args = self.args
        if (_step + 1) % max(1, int(args.gradient_accumulation_steps)) != 0:
            return
        # an optimizer step has just been taken: cadences below are in optimizer steps (trainer.py:345-366)
        if args.is_master_node and (self._global_step + 1) % args.logging_steps == 0:
            rec = {"epoch": _epoch, "global_step": self._global_step + 1, "loss": loss_val, "lr": self.learning_rate}
            self._log.append(rec)
            print("Epoch [{}/{}], step [{}/{}], lr {:.8f}, loss {:.4f}".format(_epoch, int(math.ceil(args.epoch_num)), _step + 1,
                                                                              len(self._train_loader), rec["lr"], loss_val))
        if args.save_checkpoint_steps and (self._global_step + 1) % args.save_checkpoint_steps == 0:
            if args.save_all_checkpoints:
                self.save_checkpoint()
            if self.evaluator is not None and args.is_master_node:
                self._eval_scores = self.evaluator.evaluate(model=self.model_module)
                if self._eval_scores[0][1] > self.evaluator.best_valid_score:
                    self.evaluator.best_valid_score = self._eval_scores[0][1]
                    self.save_checkpoint(save_best=True)
                self._model.train()
        self._global_step += 1

    def after_train(self):
        args = self.args
        if not args.is_master_node:
            return
        if args.save_checkpoint_steps is None:
            self.save_checkpoint(save_best=True)
        elif self.evaluator is not None:
            self._eval_scores = self.evaluator.evaluate(model=self.model_module)
            if self._eval_scores[0][1] > self.evaluator.best_valid_score:
                self.evaluator.best_valid_score = self._eval_scores[0][1]
                self.save_checkpoint(save_best=True)

    # ------------------------------------------------------------------ checkpoint (trainer.py:421-580)
    def save_checkpoint(self, save_best=False):
        args = self.args
        if not args.is_master_node or not args.checkpoint_dir:
            return
        os.makedirs(args.checkpoint_dir, exist_ok=True)
        ds = self._train_loader.dataset
        with open(os.path.join(args.checkpoint_dir, "train_config.json"), "w") as f:
            cfg = {k: v for k, v in vars(args).items() if isinstance(v, (int, float, str, bool, type(None)))}
            json.dump({"_config_json": cfg, "model_config": self.model_module.config.json_data}, f, ensure_ascii=False, indent=1)
        with open(os.path.join(args.checkpoint_dir, "label_mapping.json"), "w") as f:
            lev = getattr(ds, "label_enumerate_values", None)
            json.dump({} if not lev else {str(v): i for i, v in enumerate(lev)}, f)
        with open(os.path.join(args.checkpoint_dir, "config.json"), "w") as f:
            f.write(self.model_module.config.to_json_string())
        src = getattr(args, "pretrained_model_name_or_path", None)
        if src:
            for name in ("vocab.txt", "vocab.json", "tokenizer.json"):
                if os.path.exists(os.path.join(src, name)) and os.path.abspath(src) != os.path.abspath(args.checkpoint_dir):
                    shutil.copy(os.path.join(src, name), os.path.join(args.checkpoint_dir, name))
        prefix = "pytorch_model" if save_best else "pytorch_model_step_%d" % (self._global_step + 1)
        torch.save({k: v.cpu() for k, v in self.model_module.state_dict().items()}, os.path.join(args.checkpoint_dir, prefix + ".bin"))
        P = self.engine.params
        meta = {"epoch": self._current_epoch, "global_step": self._global_step,
                "optimizer": {"step": P.step, "exp_avg": P.exp_avg.cpu(), "exp_avg_sq": P.exp_avg_sq.cpu(), "layout": "flat:easynlp_b200.params"}}
        torch.save(meta, os.path.join(args.checkpoint_dir, prefix + ".meta.bin"))
