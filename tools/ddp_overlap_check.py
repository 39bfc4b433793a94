"""N-GPU check (torchrun) of the overlapped gradient all-reduce: same gradients as the post-backward all-reduce, and the
exposed communication time of both against a run without any collective (config 4 as shipped, batch 50 per rank).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_overlap_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import distributed as D, losses, sde_lib, workloads
from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
cfg = workloads.config4_ddpm_sr3_64()
cfg.model.dropout = 0.0
torch.manual_seed(0)
model = utils.create_model(cfg).to(dev).train()
D.broadcast_parameters(model, src=0)
sde = sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000)
loss_fn = losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                         likelihood_weighting=True)
B = 50
g = torch.Generator().manual_seed(100 + rank)
x, y = torch.rand(B, 3, 64, 64, generator=g).to(dev), torch.rand(B, 3, 64, 64, generator=g).to(dev)
noise = {"t": (torch.rand(B, generator=g) * 0.9 + 0.05).to(dev), "z": torch.randn(B, 3, 64, 64, generator=g).to(dev)}


def grads(mode):
    for p in model.parameters():
        p.grad = None
    if mode == "overlap":
        D.enable_gradient_overlap(model, segments=4)
    else:
        D.disable_gradient_overlap(model)
    loss_fn(model, (y, x), noise=noise).backward()
    if mode == "post":
        D.allreduce_gradients(model, average=True)
    return torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).clone()


def timed(mode, steps=20):
    for _ in range(3):
        grads(mode)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        grads(mode)
    e1.record()
    torch.cuda.synchronize()
    return D.max_over_ranks(e0.elapsed_time(e1) / steps, dev)


ref = grads("post")
got = grads("overlap")
got2 = grads("overlap")
err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
rep = torch.equal(got, got2)
t_none, t_post, t_ovl = timed("none"), timed("post"), timed("overlap")
hook = model._engine.grad_sync
if rank == 0:
    plan = next(iter(model._engine.train_plans.values()))
    segs = plan.grad_segments(4)
    print(json.dumps({"world": world, "max_rel_diff_overlap_vs_post": err, "overlap_reproducible": rep,
                      "ms_fwd_bwd_no_collective": t_none, "ms_post_backward_allreduce": t_post, "ms_overlapped": t_ovl,
                      "exposed_allreduce_ms_post": t_post - t_none, "exposed_allreduce_ms_overlapped": t_ovl - t_none,
                      "gradient_bytes": int(plan.gflat.numel() * 4),
                      "segments": [{"launches": hi - lo, "spans": len(r), "mbytes": sum(n for _, n in r) * 4 / 1e6}
                                   for lo, hi, r in segs]}))
dist.destroy_process_group()
