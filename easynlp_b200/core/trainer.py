"""Trainer -- the hot loop of easynlp/core/trainer.py:39-677 for the CLIP application, same constructor and call
sequence (model(batch) -> compute_loss -> backward -> every gradient_accumulation_steps: clip, step, schedule, zero_grad;
loss.item() per micro-step; rank-0 logging, periodic evaluate + checkpoint), with the optimizer side replaced by ONE fused
multi-tensor clip+AdamW launch per decay group (reference: ~2k tiny kernels from the Python loop of optimizers.py:405-464)
and DDP's gradient mean replaced by a SUM all-reduce of the flat gradient (the loss is already global-batch normalised
when `global_contrastive=True`; see DESIGN.md "multi-GPU").

Checkpoint files match the reference (trainer.py:421-580): config.json, pytorch_model[.step_N].bin with the
`chinese_clip.`-prefixed keys, *.meta.bin {epoch, global_step, optimizer}, train_config.json, label_mapping.json, vocab.txt.
"""
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
        self.use_graph = bool(kwargs.get("use_cuda_graph", True))
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
        opt = meta.get("optimizer", {})
        P = self.engine.params
        if "exp_avg" in opt:
            P.exp_avg.copy_(opt["exp_avg"].to(P.exp_avg.device)); P.exp_avg_sq.copy_(opt["exp_avg_sq"].to(P.exp_avg.device))
            self.engine.set_step(int(opt.get("step", 0)))      # host AND device step counters (Adam bias correction, schedule, dropout stream)
        self._start_epoch = int(meta.get("epoch", 0))
        self._global_step = int(meta.get("global_step", 0)) + 1
        self._sched_step = self._global_step // max(1, args.gradient_accumulation_steps)   # unlike the reference (quirk A.4-6) the schedule survives a resume

    # ------------------------------------------------------------------ loop
    def train(self):
        args = self.args
        t_start = time.time()
        self.engine.zero_grad()
        for epoch in range(self._start_epoch, int(math.ceil(args.epoch_num))):
            self._current_epoch = epoch
            if D.world_size() > 1:
                D.barrier()
                if hasattr(self._train_loader.sampler, "set_epoch"):
                    self._train_loader.sampler.set_epoch(epoch)
            self._model.train()
            skip = self._global_step - epoch * len(self._train_loader) if epoch == self._start_epoch else 0
            for _step, batch in enumerate(self._train_loader):
                if _step < skip:
                    continue
                loss_val = self.train_step(batch)
                self.after_iter(_step, epoch, loss_val)
        self.after_train()
        print("Training Time: {}".format(time.time() - t_start))

    def train_step(self, batch):
        """One micro-step on a collated batch (host or device tensors): forward, loss, backward and -- every
        gradient_accumulation_steps -- clip + AdamW + schedule + zero_grad; returns the loss as a Python float (the reference
        reads loss.item() every micro-step too, trainer.py:342).  Without gradient accumulation the whole step runs as one
        captured CUDA graph (ClipEngine.train_step); otherwise through model(batch) / compute_loss / loss.backward()."""
        args = self.args
        label_ids = batch.pop("label_ids", None)
        if args.gradient_accumulation_steps == 1:
            # one fused step: a replayed CUDA graph on one GPU; eager launches with the gradient all-reduce overlapped with the
            # backward pass when the step contains collectives (use_cuda_graph=False)
            out = self.engine.train_step(batch["pixel_values"], batch["input_ids"], lr=args.learning_rate, weight_decay=args.weight_decay,
                                         max_grad_norm=args.max_grad_norm, warmup_steps=self._warmup_steps, t_total=self._t_total,
                                         distributed=getattr(self.model_module, "distributed_loss", False), use_graph=self.use_graph)
            self._sched_step += 1
            return out["loss"].item()
        forward_outputs = self._model(batch)
        loss_dict = self.model_module.compute_loss(forward_outputs, label_ids)
        _loss = loss_dict["loss"]
        if args.gradient_accumulation_steps > 1:
            _loss = _loss / args.gradient_accumulation_steps
        _loss.backward()
        loss_val = loss_dict["loss"].item()
        if (self._global_step + 1) % args.gradient_accumulation_steps == 0:
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
