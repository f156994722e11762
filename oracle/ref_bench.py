"""Timing legs that run the UNMODIFIED reference (through oracle/ref_loader.py): bench.py's `--impl reference` arm (CPU, all host
threads it can use) and the `gpu_baseline` leg (the north-star bar: the reference's own PyTorch path on the same B200 under
torch.autocast(bfloat16)).  TEST / BENCH INFRASTRUCTURE ONLY -- never imported by easynlp_b200.

One step = what easynlp/core/trainer.py:617-677 + :306-337 do per batch: model(batch) -> compute_loss -> backward ->
clip_grad_norm_(max_grad_norm) -> AdamW.step() -> scheduler.step() -> zero_grad(), with the optimizer / schedule built by the
reference's own get_optimizer (core/optimizers.py:472-539)."""
import json
import os
import tempfile
import time

import torch


def make_reference_app(cfg: dict, sd: dict, device="cpu"):
    from .ref_loader import import_reference
    R = import_reference()
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(cfg, f)
        torch.save({"chinese_clip." + k: v for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
        with open(os.path.join(d, "vocab.txt"), "w") as f:
            f.write("[PAD]\n[UNK]\n[CLS]\n[SEP]\n")
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints its config from __init__
            app = R["CLIPApp"](d)
    app.to(device)
    app.train()
    return app, R


def make_reference_optimizer(app, lr, weight_decay=1e-4, steps_per_epoch=1000):
    import contextlib
    import io
    from easynlp.core.optimizers import get_optimizer
    with contextlib.redirect_stdout(io.StringIO()):
        opt, sched = get_optimizer(optimizer_type="AdamW", learning_rate=lr, warmup_proportion=0.0, named_parameters=list(app.named_parameters()),
                                   gradient_accumulation_steps=1, num_steps_per_epoch=steps_per_epoch, epoch_num=1, weight_decay=weight_decay)
    return opt, sched


def reference_step(app, opt, sched, pixels, ids, max_grad_norm=1.0, autocast_dtype=None):
    """one Trainer iteration of the reference (trainer.py:617-677, 306-337); returns the loss tensor"""
    batch = {"pixel_values": pixels, "input_ids": ids, "label_ids": []}
    label_ids = batch.pop("label_ids")
    if autocast_dtype is not None:
        with torch.autocast(pixels.device.type if pixels.is_cuda else "cpu", dtype=autocast_dtype):
            out = app(batch)
            loss = app.compute_loss(out, label_ids)["loss"]
    else:
        out = app(batch)
        loss = app.compute_loss(out, label_ids)["loss"]
    loss.backward()
    torch.nn.utils.clip_grad_norm_(app.parameters(), max_grad_norm)
    opt.step()
    if sched is not None:
        sched.step()
    opt.zero_grad()
    return loss


def time_reference_cpu(cfg, sd, pixels, ids, steps, warmup, threads, budget_s=60.0, lr=1e-5):
    torch.set_num_threads(threads)
    app, _ = make_reference_app(cfg, sd, "cpu")
    opt, sched = make_reference_optimizer(app, lr)
    for _ in range(warmup):
        reference_step(app, opt, sched, pixels, ids)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        loss = reference_step(app, opt, sched, pixels, ids).item()      # Trainer reads loss.item() every step (trainer.py:342)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / max(1, done)
    return {"pairs_per_s": pixels.shape[0] / dt, "s_per_step": dt, "steps": done, "loss": loss}


def time_reference_gpu(cfg, sd, pixels_host, ids_host, steps, warmup, device, lr=1e-5, autocast_dtype=torch.bfloat16):
    """The reference modules .cuda() under torch.autocast(bfloat16) (BASELINE's dtype; the reference's own --use_amp is fp16 autocast +
    GradScaler, trainer.py:57-62,296-304), reference AdamW, host batch copied every step and loss.item() read back every step -- the
    same e2e protocol as this repo's `e2e` leg.  Returns pairs/s or raises (e.g. out of memory at this batch)."""
    app, _ = make_reference_app(cfg, sd, device)
    opt, sched = make_reference_optimizer(app, lr)
    for _ in range(warmup):
        reference_step(app, opt, sched, pixels_host.to(device, non_blocking=True), ids_host.to(device, non_blocking=True), autocast_dtype=autocast_dtype).item()
    torch.cuda.synchronize(device)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = reference_step(app, opt, sched, pixels_host.to(device, non_blocking=True), ids_host.to(device, non_blocking=True),
                              autocast_dtype=autocast_dtype).item()
    e1.record(); torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    mem = torch.cuda.max_memory_allocated(device)
    del app, opt, sched
    torch.cuda.empty_cache()
    return {"pairs_per_s": pixels_host.shape[0] / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "loss": loss, "max_mem_bytes": mem}
