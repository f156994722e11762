"""Data-parallel wiring of the contrastive step (one process per GPU, torch.distributed for the plumbing).

The reference never gathers embeddings: under DDP each rank takes InfoNCE over its LOCAL [b, b] logits
(appzoo/clip/model.py:148-164) and DDP averages gradients (core/trainer.py:103-108).  The north star asks for the global
batch instead: every rank all-gathers the two [local_B, E] embedding shards, evaluates its two CE strips
[local_B, global_B] against the gathered galleries, reduce-scatters the gallery gradients back to their owners and
all-reduces (SUM: the loss is already divided by the global batch) the flat parameter gradient.  At world size 1 this
is exactly the reference loss.

The collective helpers are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group when world > 1."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def gather_rows(local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """all-gather [b, E] shards into [world*b, E] (rank-major: rank r owns rows r*b .. (r+1)*b-1)."""
    w = world_size()
    if w == 1:
        return local
    if out is None:
        out = torch.empty((w * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def reduce_scatter_rows(full: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum [world*b, E] over ranks and return this rank's [b, E] slice."""
    w = world_size()
    if w == 1:
        return full
    b = full.shape[0] // w
    if out is None:
        out = torch.empty((b,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    if dist.get_backend() == "gloo":      # gloo has no reduce_scatter: all-reduce + slice (CPU tests only)
        tmp = full.clone()
        dist.all_reduce(tmp)
        out.copy_(tmp[get_rank() * b:(get_rank() + 1) * b])
    else:
        dist.reduce_scatter_tensor(out, full.contiguous())
    return out


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(flat)
    return flat


def barrier():
    if world_size() > 1:
        dist.barrier()


class OverlappedGradReducer:
    """SUM all-reduce of a flat gradient buffer issued in pieces WHILE the backward pass is still running (what DDP's bucketing does
    for the reference, core/trainer.py:103-108).  `ready(ranges)` launches an asynchronous all-reduce of slices whose gradients are
    final (the collective runs on the backend's own stream, ordered after the kernels launched so far); `finish()` sends whatever
    was never announced and makes the current stream wait for everything.  Each element is reduced exactly once."""

    def __init__(self, flat: torch.Tensor, n: int):
        self.flat = flat
        self.n = int(n)
        self.done = []      # (start, end) already sent
        self.works = []

    def _launch(self, a: int, b: int):
        if b > a:
            self.works.append(dist.all_reduce(self.flat[a:b], async_op=True))
            self.done.append((a, b))

    def ready(self, ranges):
        if world_size() == 1:
            return
        for a, b in ranges:
            a, b = max(0, int(a)), min(self.n, int(b))
            for c, d in self.done:                     # never send an element twice
                if a < d and c < b:
                    raise RuntimeError(f"gradient range [{a},{b}) overlaps an already reduced range [{c},{d})")
            self._launch(a, b)

    def finish(self):
        if world_size() > 1:
            pos = 0
            for a, b in sorted(self.done):
                self._launch(pos, a)
                pos = max(pos, b)
            self._launch(pos, self.n)
            for w in self.works:
                w.wait()
        self.done, self.works = [], []
