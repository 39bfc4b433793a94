"""Score-matching losses (reference: losses.py:26-35, 55-97, 99-234, 236-265, 320-340, 345-407).

Same factory signatures and loss_fn(model, batch) contract as the reference, for evaluation (train=False, what
`validation_step` calls, lightning_modules/BaseSdeGenerativeModel.py:62-65) and for training (train=True:
`loss.backward()` fills the parameters' `.grad` like the reference's autograd does).

The tensor-sized arithmetic runs in libcsd_b200 kernels: the forward perturbation `mean + std * z` (csd_sde_perturb_f32),
the score network (one autograd node whose backward is the engine's planned reverse launch list), and the per-sample
weighted residual reduction (csd_dsm_loss_f32, adjoint csd_dsm_loss_bwd_f32). Per-sample scalars (t, std, g^2: `batch`
floats) are computed with the SDE objects as in the reference. There is no PyTorch fallback for the network.

loss_fn accepts an optional `noise` dict ({'t', 'z'} or {'t', 'z_x', 'z_y'}) that replaces the random draws
(parity tests); without it the draws follow the reference's order: t = torch.rand(B) on the host, then randn_like
on the data's device (losses.py:124-131,188-190,216-218).
"""
import torch
import torch.optim as optim

from . import kernels as K
from .models import utils as mutils


def get_optimizer(config, params):
    """losses.py:26-35."""
    if config.optim.optimizer == "Adam":
        return optim.Adam(params, lr=config.optim.lr, betas=(config.optim.beta1, 0.999), eps=config.optim.eps,
                          weight_decay=config.optim.weight_decay)
    raise NotImplementedError(f"Optimizer {config.optim.optimizer} not supported yet!")


def _scalars(sde, t):
    """Per-sample (mean coefficient, std, g^2) of an SDE at times t [B]."""
    mean_coef, std = sde.marginal_prob(torch.ones_like(t), t)
    g2 = sde.sde(torch.zeros_like(t), t)[1] ** 2
    return mean_coef.float().contiguous(), std.float().contiguous(), g2.float().contiguous()


class _DsmLoss(torch.autograd.Function):
    """losses[b] = w[b] * sum_i (a[b] * score[b,i] + c[b] * z[b,i])^2 with the analytic adjoint w.r.t. score."""

    @staticmethod
    def forward(ctx, score, z, a, c, w):
        score = score.contiguous()
        losses = torch.zeros(score.shape[0], device=score.device, dtype=torch.float32)
        K.dsm_loss(score, z, a, c, w, losses)
        ctx.save_for_backward(score, z, a, c, w)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        score, z, a, c, w = ctx.saved_tensors
        dscore = torch.empty_like(score)
        K.dsm_loss_bwd(score, z, a, c, w, grad_losses.contiguous().float(), dscore)
        return dscore, None, None, None, None


def _draw_t(sde_T, eps, like, noise):
    if noise is not None and "t" in noise:
        return noise["t"].to(like.device, torch.float32)
    return (torch.rand(like.shape[0]).type_as(like) * (sde_T - eps) + eps).float()


def _draw_z(like, noise, key):
    if noise is not None and key in noise:
        return noise[key].to(like.device, torch.float32).contiguous()
    return torch.randn_like(like)


def _residual(score, z, std, g2, n_total, reduce_mean, likelihood_weighting):
    """One tensor's per-sample loss term [B] (differentiable w.r.t. score)."""
    ones = torch.ones_like(std)
    red = (1.0 / n_total) if reduce_mean else 0.5
    if likelihood_weighting:      # square(score + z / std) * g^2
        return _DsmLoss.apply(score, z, ones, (1.0 / std).contiguous(), (g2 * red).contiguous())
    return _DsmLoss.apply(score, z, std, ones, torch.full_like(std, red))   # square(score * std + z)


def get_general_sde_loss_fn(sde, train, conditional=False, reduce_mean=True, continuous=True,
                            likelihood_weighting=True, eps=1e-5):
    """losses.py:99-234: unconditional, SR3 (one conditional SDE) and CMDE (x and y SDEs) estimators."""
    if conditional and isinstance(sde, dict):
        if len(sde.keys()) != 2:
            raise NotImplementedError("multi-speed losses with >= 3 SDEs (losses.py:148-183) are not on the B200 path")
        assert likelihood_weighting, ("For the variance reduction technique in inverse problems, we only support "
                                      "likelihood weighting for the time being.")

        def loss_fn(model, batch, noise=None):
            y, x = (b.float().contiguous() for b in batch)
            score_fn = mutils.get_score_fn(sde, model, conditional=conditional, train=train, continuous=continuous)
            t = _draw_t(sde["x"].T, eps, x, noise)
            z_y = _draw_z(y, noise, "z_y")
            my, std_y, g2_y = _scalars(sde["y"], t)
            z_x = _draw_z(x, noise, "z_x")
            mx, std_x, g2_x = _scalars(sde["x"], t)
            pert = {"x": K.sde_perturb(x, z_x, torch.empty_like(x), mx, std_x),
                    "y": K.sde_perturb(y, z_y, torch.empty_like(y), my, std_y)}
            score = score_fn(pert, t)
            n_total = x[0].numel() + y[0].numel()       # losses are concatenated before the reduction (:143-145)
            losses = (_residual(score["x"], z_x, std_x, g2_x, n_total, reduce_mean, True)
                      + _residual(score["y"], z_y, std_y, g2_y, n_total, reduce_mean, True))
            return losses.mean()

        return loss_fn if train else torch.no_grad()(loss_fn)

    def loss_fn(model, batch, noise=None):
        if conditional:               # SR3 estimator (losses.py:185-206): batch = (y, x), only x is perturbed
            y, x = (b.float().contiguous() for b in batch)
        else:
            y, x = None, batch.float().contiguous()
        score_fn = mutils.get_score_fn(sde, model, conditional=conditional, train=train, continuous=continuous)
        t = _draw_t(sde.T, eps, x, noise)
        z = _draw_z(x, noise, "z")
        mean_coef, std, g2 = _scalars(sde, t)
        perturbed = K.sde_perturb(x, z, torch.empty_like(x), mean_coef, std)
        score = score_fn({"x": perturbed, "y": y} if conditional else perturbed, t)
        losses = _residual(score, z, std, g2, x[0].numel(), reduce_mean, likelihood_weighting)
        return losses.mean()

    return loss_fn if train else torch.no_grad()(loss_fn)


def get_sde_loss_fn(sde, train, reduce_mean=True, continuous=True, likelihood_weighting=True, eps=1e-5):
    """losses.py:55-97: the unconditional estimator of get_general_sde_loss_fn (same arithmetic)."""
    return get_general_sde_loss_fn(sde, train, conditional=False, reduce_mean=reduce_mean, continuous=continuous,
                                   likelihood_weighting=likelihood_weighting, eps=eps)


def optimization_manager(config):
    """losses.py:38-52: warm-up and gradient clipping around optimizer.step()."""
    import numpy as np

    def optimize_fn(optimizer, params, step, lr=config.optim.lr, warmup=config.optim.warmup,
                    grad_clip=config.optim.grad_clip):
        if warmup > 0:
            for g in optimizer.param_groups:
                g["lr"] = lr * np.minimum(step / warmup, 1.0)
        if grad_clip >= 0:
            torch.nn.utils.clip_grad_norm_(params, max_norm=grad_clip)
        optimizer.step()

    return optimize_fn


def _wrap(loss_fn, train):
    return loss_fn if train else torch.no_grad()(loss_fn)


def get_smld_loss_fn(vesde, train, reduce_mean=False, likelihood_weighting=False):
    """losses.py:55-85 (legacy SMLD): square(score + noise / sigma^2) * sigma^2 with discrete sigmas. Both weighting
    branches of the reference reduce to the same per-sample value."""
    from . import sde_lib
    assert isinstance(vesde, sde_lib.VESDE), "SMLD training only works for VESDEs."

    def loss_fn(model, batch, noise=None):
        x = batch.float().contiguous()
        score_fn = mutils.get_score_fn(vesde, model, train=train)
        if noise is not None and "labels" in noise:
            labels = noise["labels"].to(x.device)
        else:
            labels = torch.randint(0, vesde.N, (x.shape[0],), device=x.device)
        sigmas = vesde.discrete_sigmas.to(x.device)[labels].float().contiguous()
        z = _draw_z(x, noise, "z")
        perturbed = K.sde_perturb(x, z, torch.empty_like(x), None, sigmas)
        score = score_fn(perturbed, labels / (vesde.N - 1))
        red = (1.0 / x[0].numel()) if reduce_mean else 0.5
        ones = torch.ones_like(sigmas)
        return _DsmLoss.apply(score, z, ones, (1.0 / sigmas).contiguous(), (sigmas ** 2 * red).contiguous()).mean()

    return _wrap(loss_fn, train)


def get_inverse_problem_smld_loss_fn(sde, train, reduce_mean=False, likelihood_weighting=True):
    """losses.py:87-139: the discrete two-SDE loss (x and y perturbed with their own sigma tables)."""

    def loss_fn(model, batch, noise=None):
        y, x = (b.float().contiguous() for b in batch)
        score_fn = mutils.get_score_fn(sde, model, train=train)
        if noise is not None and "labels" in noise:
            labels = noise["labels"].to(x.device)
        else:
            labels = torch.randint(0, sde["x"].N, (x.shape[0],), device=x.device)
        sig_y = sde["y"].discrete_sigmas.to(y.device)[labels].float().contiguous()
        sig_x = sde["x"].discrete_sigmas.to(x.device)[labels].float().contiguous()
        z_y = _draw_z(y, noise, "z_y")
        z_x = _draw_z(x, noise, "z_x")
        pert = {"x": K.sde_perturb(x, z_x, torch.empty_like(x), None, sig_x),
                "y": K.sde_perturb(y, z_y, torch.empty_like(y), None, sig_y)}
        score = score_fn(pert, labels / (sde["x"].N - 1))
        n_total = x[0].numel() + y[0].numel()
        red = (1.0 / n_total) if reduce_mean else 0.5
        ones = torch.ones_like(sig_x)
        if likelihood_weighting:       # square(score - target) * sigma^2, concatenated, then reduced
            wx, wy = sig_x ** 2 * red, sig_y ** 2 * red
        else:                          # reduce(square(score - target)) * sigma_x^2 sigma_y^2 / (sigma_x^2 + sigma_y^2)
            smld = (sig_x ** 2 * sig_y ** 2) / (sig_x ** 2 + sig_y ** 2)
            wx = wy = smld * red
        losses = (_DsmLoss.apply(score["x"], z_x, ones, (1.0 / sig_x).contiguous(), wx.contiguous())
                  + _DsmLoss.apply(score["y"], z_y, ones, (1.0 / sig_y).contiguous(), wy.contiguous()))
        return losses.mean()

    return _wrap(loss_fn, train)


def get_ddpm_loss_fn(vpsde, train, reduce_mean=True):
    """losses.py:320-340 (legacy DDPM): square(model(x_t, labels) - noise)."""
    from . import sde_lib
    assert isinstance(vpsde, sde_lib.VPSDE), "DDPM training only works for VPSDEs."

    def loss_fn(model, batch, noise=None):
        x = batch.float().contiguous()
        model_fn = mutils.get_model_fn(model, train=train)
        if noise is not None and "labels" in noise:
            labels = noise["labels"].to(x.device)
        else:
            labels = torch.randint(0, vpsde.N, (x.shape[0],), device=x.device)
        a = vpsde.sqrt_alphas_cumprod.to(x.device)[labels].float().contiguous()
        s = vpsde.sqrt_1m_alphas_cumprod.to(x.device)[labels].float().contiguous()
        z = _draw_z(x, noise, "z")
        perturbed = K.sde_perturb(x, z, torch.empty_like(x), a, s)
        out = model_fn(perturbed, labels)
        red = (1.0 / x[0].numel()) if reduce_mean else 0.5
        ones = torch.ones_like(a)
        return _DsmLoss.apply(out, z, ones, -ones, torch.full_like(a, red)).mean()

    return _wrap(loss_fn, train)


def get_step_fn(sde, train, optimize_fn=None, reduce_mean=False, continuous=True, likelihood_weighting=False):
    """losses.py:345-407: one training / evaluation step on a {'model', 'optimizer', 'ema', 'step'} state."""
    from . import sde_lib
    if continuous:
        loss_fn = get_sde_loss_fn(sde, train, reduce_mean=reduce_mean, continuous=True,
                                  likelihood_weighting=likelihood_weighting)
    elif isinstance(sde, dict):
        loss_fn = get_inverse_problem_smld_loss_fn(sde, train, reduce_mean=reduce_mean,
                                                   likelihood_weighting=likelihood_weighting)
    elif isinstance(sde, sde_lib.VESDE):
        loss_fn = get_smld_loss_fn(sde, train, reduce_mean=reduce_mean)
    elif isinstance(sde, sde_lib.VPSDE):
        loss_fn = get_ddpm_loss_fn(sde, train, reduce_mean=reduce_mean)
    else:
        raise ValueError(f"Discrete training for {sde.__class__.__name__} is not recommended.")

    def step_fn(state, batch):
        model = state["model"]
        if train:
            optimizer = state["optimizer"]
            optimizer.zero_grad()
            loss = loss_fn(model, batch)
            loss.backward()
            optimize_fn(optimizer, model.parameters(), step=state["step"])
            state["step"] += 1
            state["ema"].update(model.parameters())
        else:
            with torch.no_grad():
                ema = state["ema"]
                ema.store(model.parameters())
                ema.copy_to(model.parameters())
                loss = loss_fn(model, batch)
                ema.restore(model.parameters())
        return loss

    return step_fn
