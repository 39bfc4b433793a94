"""Diagnostic: are the engine's parameter gradients consumed in place by FusedAdamEMA (views of one buffer)?
Run plain and under compute-sanitizer; prints which condition of optim.FusedAdamEMA._flat_grads fails, if any."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_utils import golden, to_namespace  # noqa: E402
from conditional_score_diffusion_b200 import losses, optim, sde_lib  # noqa: E402
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: E402,F401

f = golden()["ncsnpp_cifar"]
m = utils.create_model(to_namespace(f["config"]))
m.load_state_dict(f["state_dict"], strict=True)
m = m.cuda().train()
sde = sde_lib.VESDE(0.01, 50, 1000)
fn = losses.get_sde_loss_fn(sde, train=True, reduce_mean=True, continuous=True, likelihood_weighting=False, eps=1e-5)
opt = optim.FusedAdamEMA(m.parameters(), lr=1e-3, grad_clip=1.0, ema_decay=0.999, model=m)
x = torch.rand(2, 3, 16, 16).cuda()
for it in range(3):
    opt.zero_grad()
    loss = fn(m, x)
    loss.backward()
    grads = [p.grad for p in opt.params]
    bases = {g.untyped_storage().data_ptr() for g in grads if g is not None}
    none = sum(g is None for g in grads)
    first = next(i for i, g in enumerate(grads) if g is not None)
    start = grads[first].storage_offset() - opt.offsets[first]
    bad_off = sum(1 for g, o in zip(grads, opt.offsets) if g is not None and g.storage_offset() - o != start)
    flat_g = opt._flat_grads()
    print(f"iter {it}: storages={len(bases)} none={none} start={start} bad_offsets={bad_off} "
          f"contig={all(g.is_contiguous() for g in grads if g is not None)} "
          f"zero_copy={flat_g.data_ptr() != opt.gflat.data_ptr()} loss={loss.item():.4f}")
    opt.step()
