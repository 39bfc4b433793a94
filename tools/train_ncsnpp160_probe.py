"""GPU probe: one CMDE training step of the bench network (ncsnpp_paired nf96, 160x160) - memory and step time.
usage: python tools/train_ncsnpp160_probe.py [batch]"""
import math
import sys

import torch

sys.path.insert(0, ".")
import bench
from conditional_score_diffusion_b200 import losses, optim, sde_lib
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = bench.workload_config()
cfg.model.dropout = 0.1
torch.manual_seed(0)
model = utils.create_model(cfg).cuda().train()
sde = {"x": sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000),
       "y": sde_lib.VESDE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)}
loss_fn = losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                         likelihood_weighting=True)
opt = optim.FusedAdamEMA(model.parameters(), lr=2e-4, grad_clip=1.0, ema_decay=0.999, model=model)
x = torch.rand(B, 3, 160, 160, device="cuda")
y = torch.rand(B, 3, 160, 160, device="cuda")
vals = []
for i in range(4):
    opt.zero_grad()
    loss = loss_fn(model, (y, x))
    loss.backward()
    opt.step()
    vals.append(loss.item())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for i in range(n):
    opt.zero_grad()
    loss = loss_fn(model, (y, x))
    loss.backward()
    opt.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
plan = next(iter(model._engine.train_plans.values()))
print(f"ncsnpp_paired nf96 160px CMDE training step: batch {B}, {ms:.1f} ms, {B / ms * 1e3:.1f} images/s, losses {['%.3f' % v for v in vals]}, "
      f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB, activations {plan.pool.nbytes() / 1e9:.1f} GB, "
      f"launches {len(plan.rec.ops)} + {len(plan.bwd.ops)}")
assert all(math.isfinite(v) for v in vals)
