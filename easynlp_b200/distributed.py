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


class PeerGroup:
    """CUDA-IPC mapped peer buffers of the contrastive head (csrc/peer.cu): every rank owns one buffer
        [ gallery_image G x E | gallery_text G x E | dgallery_image G x E | dgallery_text G x E | flags 4 x world ]      (fp32 / uint32)
    and maps every other rank's buffer.  The embedding all-gather then is the l2-normalise kernel storing straight into all peers'
    galleries, the gradient reduce-scatter a kernel that pulls this rank's rows from all peers.  Construction is collective (handles are
    exchanged with all_gather_object); `PeerGroup.create` returns None where peer mapping is impossible (the caller keeps the
    torch.distributed collectives)."""

    CH_GALLERY, CH_GRADS = 0, 1

    def __init__(self, local_rows: int, E: int, device):
        import ctypes as C
        from . import _lib as L
        self.L = L; self.C = C
        self.world, self.rank = world_size(), get_rank()
        self.rows, self.E = int(local_rows), int(E)
        self.G = self.world * self.rows
        self.dev = torch.device(device)
        self.region = self.G * self.E * 4                      # bytes of one [G, E] fp32 region
        self.flag_off = 4 * self.region
        nbytes = self.flag_off + 4 * self.world * 4 + 256
        base = C.c_void_p()
        L.check(L.lib().clipk_peer_alloc(C.byref(base), nbytes), "peer_alloc")
        self.base = base.value
        handle = C.create_string_buffer(64)
        L.check(L.lib().clipk_peer_export(C.c_void_p(self.base), handle), "peer_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw))
        self.bases = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.bases.append(self.base)
            else:
                ptr = C.c_void_p()
                L.check(L.lib().clipk_peer_open(C.c_char_p(h), C.byref(ptr)), "peer_open")
                self.bases.append(ptr.value)
        tab = lambda off: torch.tensor([b + off for b in self.bases], dtype=torch.int64, device=self.dev)
        self.gi_ptrs, self.gt_ptrs = tab(0), tab(self.region)
        self.dgi_ptrs, self.dgt_ptrs = tab(2 * self.region), tab(3 * self.region)
        self.flag_ptrs = tab(self.flag_off)
        self.epoch = 0
        dist.barrier()

    @classmethod
    def create(cls, local_rows, E, device):
        if world_size() == 1 or not torch.cuda.is_available() or os.environ.get("CLIPK_PEER", "1") == "0":
            return None
        try:
            ok = torch.ones(1, device=device)
            try:
                pg = cls(local_rows, E, device)
            except Exception as ex:      # every rank must learn that SOME rank failed, or the others would wait forever later
                pg = None; ok.zero_()
                print(f"easynlp_b200: peer-memory collectives unavailable on rank {get_rank()}: {ex!r}", flush=True)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            return pg if ok.item() > 0 else None
        except Exception:
            return None

    def _view(self, off_regions):
        """torch view [G, E] fp32 of a region of the LOCAL buffer (device memory owned by the library, not by torch's allocator)"""
        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (self.G, self.E), "typestr": "<f4", "data": (self.base + off_regions * self.region, False), "version": 2}
        return torch.as_tensor(m, device=self.dev)

    def gallery_views(self):
        return self._view(0), self._view(1)

    def grad_views(self):
        return self._view(2), self._view(3)

    # ---- collectives (all asynchronous on the current stream)
    def l2norm_allgather(self, x, y_local, norm, which):
        from .ops import _stream
        ptrs = self.gi_ptrs if which == "image" else self.gt_ptrs
        self.L.check(self.L.lib().clipk_l2norm_allgather(self.C.c_void_p(x.data_ptr()), self.C.c_void_p(y_local.data_ptr()), self.C.c_void_p(norm.data_ptr()),
                                                         self.C.c_void_p(ptrs.data_ptr()), self.world, self.rank, self.rows, self.E, _stream()), "l2norm_allgather")

    def sync(self, channel):
        """every rank's writes issued so far are visible to every other rank once this returns (on the stream)"""
        from .ops import _stream
        if channel == self.CH_GALLERY:
            self.epoch += 1
        self.L.check(self.L.lib().clipk_peer_signal(self.C.c_void_p(self.flag_ptrs.data_ptr()), self.world, self.rank, channel, self.epoch, _stream()), "peer_signal")
        self.L.check(self.L.lib().clipk_peer_wait(self.C.c_void_p(self.base + self.flag_off), self.world, channel, self.epoch, _stream()), "peer_wait")

    def reduce_rows(self, which, out, accumulate=True):
        from .ops import _stream
        ptrs = self.dgi_ptrs if which == "image" else self.dgt_ptrs
        self.L.check(self.L.lib().clipk_peer_reduce_rows(self.C.c_void_p(ptrs.data_ptr()), self.world, self.rank, self.C.c_void_p(out.data_ptr()), self.rows, self.E,
                                                         int(accumulate), _stream()), "peer_reduce_rows")
