"""GPU probe: one training step of BASELINE config 4 (edges2shoes_SR3.py: ddpm_paired_SR3, nf 128, ch_mult (1,1,2,2),
attention at 16/8, 64x64, dropout 0.1, batch 50 per GPU, SR3 loss with likelihood weighting, Adam 2e-4, clip 1.0):
step time with CUDA events and a per-kernel breakdown of the forward and the backward launch lists.
usage: python tools/train_step_profile.py [batch] [image_size] [fused]   (fused = optim.FusedAdamEMA instead of torch Adam)"""
import collections
import math
import sys
import time
from types import SimpleNamespace as NS

import torch

sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as K, losses, sde_lib
from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 64


def sr3_config(hw):
    c = NS()
    c.training = NS(continuous=True)
    c.data = NS(image_size=hw, effective_image_size=hw, num_channels=6, centered=False)
    c.model = NS(name="ddpm_paired_SR3", nf=128, ch_mult=(1, 1, 2, 2), num_res_blocks=2, attn_resolutions=(16, 8),
                 dropout=0.1, resamp_with_conv=True, conditional=True, nonlinearity="swish", input_channels=6,
                 output_channels=3, num_scales=1000)
    c.optim = NS(weight_decay=0, optimizer="Adam", lr=2e-4, beta1=0.9, eps=1e-8, warmup=2500, grad_clip=1.0)
    return c


def main():
    cfg = sr3_config(HW)
    torch.manual_seed(0)
    model = utils.create_model(cfg).cuda().train()
    nparams = sum(p.numel() for p in model.parameters())
    sde = sde_lib.cVESDE(5e-3, math.sqrt(3 * HW * HW), 1000)
    loss_fn = losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                             likelihood_weighting=True)
    fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
    if fused:
        from conditional_score_diffusion_b200 import optim
        opt = optim.FusedAdamEMA(model.parameters(), lr=2e-4, grad_clip=1.0, ema_decay=0.999, warmup=2500, model=model)
        optimize_fn = lambda o, p, step: o.step()
    else:
        opt = losses.get_optimizer(cfg, model.parameters())
        optimize_fn = losses.optimization_manager(cfg)
    x = torch.rand(B, 3, HW, HW, device="cuda")
    y = torch.rand(B, 3, HW, HW, device="cuda")

    def step(i):
        opt.zero_grad()
        loss = loss_fn(model, (y, x))
        loss.backward()
        optimize_fn(opt, model.parameters(), step=i)
        return loss

    for i in range(3):
        l = step(i + 1)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(n):
        l = step(i + 4)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.time() - t0) / n * 1e3
    ms = e0.elapsed_time(e1) / n
    print(f"train step: {ms:.2f} ms device, {wall:.2f} ms wall, batch {B}, {HW}px, {nparams / 1e6:.1f} M params, "
          f"{B / ms * 1e3:.1f} images/s, loss {l.item():.4f}")
    # phases
    for name, fn in (("zero_grad", lambda: opt.zero_grad()),):
        pass
    plan = next(iter(model._engine.train_plans.values()))

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    print(f"  forward list   {timed(plan.rec.run):8.3f} ms ({len(plan.rec.ops)} ops)")
    print(f"  backward list  {timed(plan.bwd.run):8.3f} ms ({len(plan.bwd.ops)} ops)")
    print(f"  weight refresh {timed(model._engine._refresh):8.3f} ms")
    print(f"  param_grads    {timed(plan.param_grads):8.3f} ms")
    print(f"  clip + adam    {timed(lambda: optimize_fn(opt, model.parameters(), step=100)):8.3f} ms")
    for label, rec in (("forward", plan.rec), ("backward", plan.bwd)):
        evs = []
        for fn, a, kw in rec.ops:
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); fn(*a, **kw); a1.record()
            evs.append((fn, a, kw, a0, a1))
        torch.cuda.synchronize()
        by_kind = collections.defaultdict(lambda: [0, 0.0])
        for fn, a, kw, a0, a1 in evs:
            name = getattr(fn, "__name__", str(fn))
            by_kind[name][0] += 1; by_kind[name][1] += a0.elapsed_time(a1)
        total = sum(v[1] for v in by_kind.values())
        print(f"{label}: {total:.3f} ms (event-timed per op, includes launch gaps)")
        for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f"    {k:24s} n={v[0]:4d} {v[1]:8.3f} ms {100 * v[1] / total:5.1f}%")
    print(f"  activation pool {plan.pool.nbytes() / 1e9:.2f} GB, pixmajor buffers "
          f"{sum(t.numel() * 2 for t in plan.pm_bufs.values()) / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
