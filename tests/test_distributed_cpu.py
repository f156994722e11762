"""CPU, world_size 2 (gloo): the collective wiring of the data-parallel contrastive step.

The strip math itself lives in CUDA kernels; here the SAME decomposition is evaluated with a torch restatement of one CE strip so
that the algebra the GPU path relies on is pinned:  sum over ranks of the strip losses == the oracle's global-batch loss, and
local strip gradients + reduce-scattered gallery gradients == the oracle's gradient w.r.t. each rank's embeddings."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _strip(Q, K, scale, off):
    """one CE strip: sum_i (lse_i - s_{i,off+i})"""
    S = (Q @ K.t()) * scale
    lab = off + torch.arange(Q.shape[0])
    return torch.nn.functional.cross_entropy(S, lab, reduction="sum")


def _worker(rank, world, port, b, E, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from easynlp_b200 import distributed as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world) and D.world_size() == world and D.get_rank() == rank
    g = torch.Generator().manual_seed(0)
    T_all = torch.nn.functional.normalize(torch.randn(world * b, E, generator=g), dim=-1)
    I_all = torch.nn.functional.normalize(torch.randn(world * b, E, generator=g), dim=-1)
    T = T_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    I = I_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    scale = torch.tensor(1 / 0.07)
    GT = D.gather_rows(T.detach()); GI = D.gather_rows(I.detach())
    assert torch.equal(GT, T_all) and torch.equal(GI, I_all)            # rank-major gather order
    GTg = GT.clone().requires_grad_(True); GIg = GI.clone().requires_grad_(True)
    G = world * b
    local = (_strip(T, GIg, scale, rank * b) + _strip(I, GTg, scale, rank * b)) / (2 * G)
    local.backward()
    dT = T.grad + D.reduce_scatter_rows(GTg.grad)
    dI = I.grad + D.reduce_scatter_rows(GIg.grad)
    total = local.detach().clone()
    D.allreduce_sum_(total)
    # oracle: single process on the concatenated global batch
    Tr = T_all.clone().requires_grad_(True); Ir = I_all.clone().requires_grad_(True)
    ref = O.clip_loss((Tr @ Ir.t()) * scale)
    ref.backward()
    ok = (abs(total.item() - ref.item()) < 1e-5
          and torch.allclose(dT, Tr.grad[rank * b:(rank + 1) * b], atol=1e-6)
          and torch.allclose(dI, Ir.grad[rank * b:(rank + 1) * b], atol=1e-6))
    flat = torch.full((8,), float(rank + 1))
    D.allreduce_sum_(flat)
    ok = ok and bool((flat == sum(range(1, world + 1))).all())
    # overlapped gradient all-reduce: slices announced out of order + the never-announced remainder, each element exactly once
    n = 1000
    flat = torch.arange(n + 24, dtype=torch.float32) * (rank + 1)       # [n, n+24) lies outside the trainable range: untouched
    red = D.OverlappedGradReducer(flat, n)
    red.ready([(640, 768), (128, 256)])
    red.ready([(0, 64)])
    try:
        red.ready([(700, 800)])
        ok = False                                                      # overlapping an already sent range must raise
    except RuntimeError:
        pass
    red.finish()
    want = torch.arange(n + 24, dtype=torch.float32) * sum(range(1, world + 1))
    want[n:] = torch.arange(n, n + 24, dtype=torch.float32) * (rank + 1)
    ok = ok and torch.equal(flat, want) and red.done == [] and red.works == []
    D.barrier()
    q.put((rank, ok, total.item(), ref.item()))
    dist.destroy_process_group()


def test_global_contrastive_decomposition_world2():
    world, b, E = 2, 6, 32
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, b, E, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, tot, ref in res:
        assert ok, (rank, tot, ref)


def test_single_process_helpers_are_identity():
    from easynlp_b200 import distributed as D
    x = torch.randn(4, 8)
    assert D.world_size() == 1 and D.get_rank() == 0
    assert D.gather_rows(x) is x and D.reduce_scatter_rows(x) is x and D.allreduce_sum_(x) is x
    flat = torch.ones(16)
    red = D.OverlappedGradReducer(flat, 16)
    red.ready([(0, 8)]); red.finish()
    assert torch.equal(flat, torch.ones(16))


def test_layer_gradient_ranges_cover_each_parameter_once():
    """the overlapped all-reduce sends, per finished layer, the contiguous slices ParamStore.ranges_for names: per-layer prefixes +
    the remainder must tile the trainable range without overlap"""
    from easynlp_b200.params import ParamStore
    cfg = O.tiny_config()
    st = ParamStore(cfg, device="cpu", with_optimizer_state=False)
    seen = torch.zeros(st.n_trainable, dtype=torch.int32)
    prefixes = [f"visual.transformer.resblocks.{i}." for i in range(cfg["vision_layers"])] + \
               [f"bert.encoder.layer.{i}." for i in range(cfg["text_num_hidden_layers"])] + ["bert.embeddings."]
    for pre in prefixes:
        rs = st.ranges_for(pre)
        assert 1 <= len(rs) <= 2, (pre, rs)                             # decay group + no-decay group
        for a, b in rs:
            seen[a:b] += 1
        names = [n for n in st.trainable_names() if n.startswith(pre)]
        assert sum(b - a for a, b in rs) >= sum(st.p(n).numel() for n in names)
        for n in names:
            o = st.offsets[n]
            assert any(a <= o and o + st.p(n).numel() <= b for a, b in rs)
    assert int(seen.max()) == 1                                         # no element announced twice
    assert st.ranges_for("bert.pooler.") == []                           # never receives a gradient: outside the reduced range
