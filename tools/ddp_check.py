"""2-GPU check (torchrun): torch's DistributedDataParallel around the engine-backed network (what Lightning's
accelerator='ddp' does, run_lib.py:55-57) against distributed.allreduce_gradients on the same shards.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from golden_utils import golden, to_namespace
from conditional_score_diffusion_b200 import distributed as D, losses, sde_lib
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
f = golden()["ncsnpp_cifar"]
cfg = to_namespace(f["config"])
cfg.model.dropout = 0.0


def make():
    m = utils.create_model(cfg)
    m.load_state_dict(f["state_dict"], strict=True)
    return m.cuda()


sde = sde_lib.VESDE(0.01, 50, 1000)
fn = losses.get_sde_loss_fn(sde, train=True, reduce_mean=True, continuous=True, likelihood_weighting=False)
g = torch.Generator().manual_seed(100 + rank)          # every rank its own shard
x = torch.rand(2, 3, 16, 16, generator=g).cuda()
noise = {"t": (torch.rand(2, generator=g) * 0.9 + 0.05).cuda(), "z": torch.randn(2, 3, 16, 16, generator=g).cuda()}

m1 = make()
fn(m1, x, noise=noise).backward()
nbytes = D.allreduce_gradients(m1, average=True)
ref = torch.cat([p.grad.flatten() for p in m1.parameters() if p.grad is not None])

m2 = make()
ddp = torch.nn.parallel.DistributedDataParallel(m2, device_ids=[local])


class Wrapper(torch.nn.Module):      # loss_fn calls model(x, labels) and reads model.embedding_type / .train()
    pass


ddp.embedding_type = m2.embedding_type
ddp.forward_scaled = None
del ddp.forward_scaled
loss = fn(ddp, x, noise=noise)
loss.backward()
got = torch.cat([p.grad.flatten() for p in m2.parameters() if p.grad is not None])
err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
gathered = [torch.empty_like(got) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, got)
same = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    print(f"DDP_CHECK flat all-reduce bytes {nbytes}, DDP vs allreduce_gradients max rel err {err:.3e}, identical across ranks {same}")
    assert err < 1e-5 and same
dist.destroy_process_group()
