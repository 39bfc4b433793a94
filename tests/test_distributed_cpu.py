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
