"""world_size-2 gloo tests (CPU) of the multi-GPU sampling plumbing (conditional_score_diffusion_b200.distributed):
weight broadcast, batch sharding, per-rank seeds, optional gather. No CUDA kernel runs here; the sampler is a
stand-in that follows the (model, y) -> (samples, info) calling convention of sampling/conditional.py:180-226."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conditional_score_diffusion_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        torch.manual_seed(100 + rank)                       # ranks start with DIFFERENT weights
        model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.GroupNorm(4, 8), torch.nn.Conv2d(8, 3, 3))
        before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
        nbytes = D.broadcast_parameters(model, src=0)
        after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.empty_like(after) for _ in range(world_size)]
        dist.all_gather(gathered, after)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        changed = not torch.equal(before, after)

        y = torch.arange(10, dtype=torch.float32).view(10, 1, 1, 1).expand(10, 1, 2, 2).contiguous()

        def fake_sampler(model, y_shard):
            # marks every sample with the rank-local noise so the per-rank seed is observable
            return y_shard * 10 + torch.rand(1).item(), {"n": y_shard.shape[0]}

        out, info = D.sample_sharded(fake_sampler, model, y, gather=True, seed=7)
        lo, hi = D.shard_range(10, rank, world_size)
        # training: gradient all-reduce, (a) gradients as views of one flat buffer (the engine's layout) reduced in
        # place with one collective, (b) ordinary per-parameter gradients through the bucketed path
        params = [p for p in model.parameters()]
        flat = torch.full((sum(p.numel() for p in params),), float(rank + 1))
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        n_flat = D.allreduce_gradients(model, average=True)
        flat_ok = bool(torch.all(flat == 1.5)) and n_flat == flat.numel() * 4
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
        n_b = D.allreduce_gradients(params, average=False, bucket_bytes=256)
        bucket_ok = all(bool(torch.all(p.grad == 3.0 * (i + 1))) for i, p in enumerate(params)) and n_b == n_flat
        results[rank] = {"same": same, "changed": changed, "nbytes": nbytes, "out": out.clone(), "n": info["n"],
                         "range": (lo, hi), "max": D.max_over_ranks(float(rank + 1), torch.device("cpu")),
                         "flat_ok": flat_ok, "bucket_ok": bucket_ok}
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_a_no_op():
    m = torch.nn.Linear(2, 2)
    assert D.world() == (0, 1)
    assert D.broadcast_parameters(m) == 0
    out, _ = D.sample_sharded(lambda model, y: (y + 1, {}), m, torch.zeros(4, 1))
    assert out.shape == (4, 1)


def test_broadcast_shard_gather_world2():
    world_size = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world_size, port, results), nprocs=world_size, join=True)
        r = dict(results)
    assert set(r) == {0, 1}
    assert r[0]["same"] and r[1]["same"], "parameters differ across ranks after the broadcast"
    assert not r[0]["changed"] and r[1]["changed"], "rank 0 is the source; rank 1 must have been overwritten"
    assert r[0]["nbytes"] == r[1]["nbytes"] > 0
    assert r[0]["range"] == (0, 5) and r[1]["range"] == (5, 10)
    assert r[0]["n"] == 5 and r[1]["n"] == 5
    assert torch.equal(r[0]["out"], r[1]["out"]) and r[0]["out"].shape[0] == 10
    # sample i came from rank i // 5, and the two ranks used different seeds (7 and 8)
    base = torch.arange(10, dtype=torch.float32) * 10
    frac = r[0]["out"][:, 0, 0, 0] - base
    assert torch.allclose(frac[:5], frac[:1].expand(5)) and torch.allclose(frac[5:], frac[5:6].expand(5))
    assert abs(frac[0].item() - frac[5].item()) > 1e-6
    assert r[0]["max"] == 2.0 and r[1]["max"] == 2.0
    assert r[0]["flat_ok"] and r[1]["flat_ok"], "in-place all-reduce of the flat gradient buffer"
    assert r[0]["bucket_ok"] and r[1]["bucket_ok"], "bucketed gradient all-reduce"


# ---- overlapped gradient all-reduce (training, DDP) ---------------------------------------------------------------
def _overlap_worker(rank, world_size, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        # the hook is fed what engine_train.TrainPlan.grad_segments produces: disjoint spans of one flat buffer, segment
        # by segment; after the last segment the buffer must equal the plain averaged all-reduce
        g = torch.Generator().manual_seed(rank)
        flat = torch.randn(1000, generator=g)
        ref = flat.clone()
        dist.all_reduce(ref)
        ref /= world_size
        sync = D.OverlappedGradSync(segments=3, average=True)
        segs = [[(600, 400)], [(200, 100), (350, 250)], [(0, 200), (300, 50)]]       # tail first, holes filled last
        for i, ranges in enumerate(segs):
            sync(flat, ranges, last=(i == len(segs) - 1))
        covered = sorted(r for s in segs for r in s)
        results[rank] = {"ok": bool(torch.allclose(flat, ref)), "bytes": sync.bytes_last, "calls": sync.calls_last,
                         "covered": covered}

        class Net(torch.nn.Module):      # engine-backed networks carry `_engine`; allreduce_gradients defers to the hook
            pass
        net = Net()
        net._engine = type("E", (), {"grad_sync": None})()
        hook = D.enable_gradient_overlap(net, segments=2)
        results[rank]["installed"] = hook is not None and net._engine.grad_sync is hook
        results[rank]["noop"] = D.allreduce_gradients(net) == 0
        D.disable_gradient_overlap(net)
        results[rank]["removed"] = net._engine.grad_sync is None
    finally:
        dist.destroy_process_group()


def _run_overlap(rank, world_size, port, q):
    res = {}
    _overlap_worker(rank, world_size, port, res)
    q.put((rank, res[rank]))


def test_overlapped_gradient_sync_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert out[r]["ok"], "segment-wise all-reduce differs from the plain averaged all-reduce"
        assert out[r]["bytes"] == 1000 * 4 and out[r]["calls"] == 5
        assert out[r]["installed"] and out[r]["noop"] and out[r]["removed"]


def test_grad_segments_cover_the_flat_buffer_once():
    """TrainPlan.grad_segments on a stand-in plan: spans are disjoint, cover every parameter exactly once, and a parameter
    lands in the segment of the LAST tape entry that touched it (late writers such as Dense_0 stay out of early ones)."""
    from conditional_score_diffusion_b200 import kernels as K
    from conditional_score_diffusion_b200.engine_train import TrainPlan

    class P:      # parameter stand-in
        def __init__(self, n, ptr):
            self._n, self._ptr = n, ptr

        def numel(self):
            return self._n

        def data_ptr(self):
            return self._ptr

    plan = TrainPlan.__new__(TrainPlan)
    params = [P(10, 1), P(7, 2), P(64, 3), P(5, 4), P(32, 5), P(9, 6)]
    plan.param_list = params
    plan.param_offsets, off = [], 0
    for p in params:
        plan.param_offsets.append(off)
        off += K.ceil_to(p.numel(), 4)
    plan.scratch = {}
    plan.bwd = type("R", (), {"ops": [None] * 40})()
    plan._entry_end = [4, 9, 15, 22, 30, 40]                   # op count after each tape entry (backward order)
    # backward order: last layers first. param 6 (tail) touched by entry 0, param 5 by entry 1, param 4 by entries 1 AND 5
    # (a late writer), param 3 by entry 2, param 2 by entry 4, param 1 by entry 5
    plan._grad_touch = {6: 0, 5: 1, 4: 5, 3: 2, 2: 4, 1: 5}
    segs = plan.grad_segments(3)
    assert [s[:2] for s in segs][0][0] == 0 and segs[-1][1] == 40
    assert all(a[1] == b[0] for a, b in zip(segs[:-1], segs[1:]))
    spans = sorted(sp for _, _, r in segs for sp in r)
    assert spans[0][0] == 0 and all(a[0] + a[1] <= b[0] for a, b in zip(spans[:-1], spans[1:]))
    assert sum(n for _, n in spans) == off
    seg_of = {}
    for i, (_, _, r) in enumerate(segs):
        for o, n in r:
            for p, po in zip(params, plan.param_offsets):
                if o <= po < o + n:
                    seg_of[p.data_ptr()] = i
    assert seg_of[6] == 0 and seg_of[4] == len(segs) - 1 and seg_of[1] == len(segs) - 1
    assert seg_of[5] <= seg_of[3] <= seg_of[2]
