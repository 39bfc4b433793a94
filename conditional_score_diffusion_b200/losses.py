"""Score-matching losses (reference: losses.py:26-35, 55-97, 99-234, 236-265, 320-340).

Same factory signatures and loss_fn(model, batch) contract as the reference. This round the losses are
EVALUATION ONLY (what `validation_step` calls, lightning_modules/BaseSdeGenerativeModel.py:62-65): the score
network has no backward on the B200 engine yet, so a loss function built with train=True raises
NotImplementedError when called - it does not fall back to PyTorch autograd.

The tensor-sized arithmetic runs in libcsd_b200 kernels: the forward perturbation `mean + std * z` (csd_sde_perturb_f32),
the score network, and the per-sample weighted residual reduction (csd_dsm_loss_f32). Per-sample scalars
(t, std, g^2: `batch` floats) are computed with the SDE objects as in the reference.

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


def _check_eval(train):
    if train:
        raise NotImplementedError(
            "training losses need the score network's backward pass, which the B200 engine does not implement yet "
            "(SURVEY.md §8 a16); build the loss with train=False for evaluation. No PyTorch fallback on purpose.")


def _draw_t(sde_T, eps, like, noise):
    if noise is not None and "t" in noise:
        return noise["t"].to(like.device, torch.float32)
    return (torch.rand(like.shape[0]).type_as(like) * (sde_T - eps) + eps).float()


def _draw_z(like, noise, key):
    if noise is not None and key in noise:
        return noise[key].to(like.device, torch.float32).contiguous()
    return torch.randn_like(like)


def _residual(losses, score, z, std, g2, n_total, reduce_mean, likelihood_weighting):
    """Accumulate one tensor's per-sample loss term into `losses` [B]."""
    ones = torch.ones_like(std)
    red = (1.0 / n_total) if reduce_mean else 0.5
    if likelihood_weighting:      # square(score + z / std) * g^2
        K.dsm_loss(score, z, ones, (1.0 / std).contiguous(), (g2 * red).contiguous(), losses)
    else:                         # square(score * std + z)
        K.dsm_loss(score, z, std, ones, torch.full_like(std, red), losses)


def get_general_sde_loss_fn(sde, train, conditional=False, reduce_mean=True, continuous=True,
                            likelihood_weighting=True, eps=1e-5):
    """losses.py:99-234: unconditional, SR3 (one conditional SDE) and CMDE (x and y SDEs) estimators."""
    if conditional and isinstance(sde, dict):
        if len(sde.keys()) != 2:
            raise NotImplementedError("multi-speed losses with >= 3 SDEs (losses.py:148-183) are not on the B200 path")
        assert likelihood_weighting, ("For the variance reduction technique in inverse problems, we only support "
                                      "likelihood weighting for the time being.")

        @torch.no_grad()
        def loss_fn(model, batch, noise=None):
            _check_eval(train)
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
            losses = torch.zeros(x.shape[0], device=x.device, dtype=torch.float32)
            _residual(losses, score["x"].contiguous(), z_x, std_x, g2_x, n_total, reduce_mean, True)
            _residual(losses, score["y"].contiguous(), z_y, std_y, g2_y, n_total, reduce_mean, True)
            return losses.mean()

        return loss_fn

    @torch.no_grad()
    def loss_fn(model, batch, noise=None):
        _check_eval(train)
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
        losses = torch.zeros(x.shape[0], device=x.device, dtype=torch.float32)
        _residual(losses, score.contiguous(), z, std, g2, x[0].numel(), reduce_mean, likelihood_weighting)
        return losses.mean()

    return loss_fn


def get_sde_loss_fn(sde, train, reduce_mean=True, continuous=True, likelihood_weighting=True, eps=1e-5):
    """losses.py:55-97: the unconditional estimator of get_general_sde_loss_fn (same arithmetic)."""
    return get_general_sde_loss_fn(sde, train, conditional=False, reduce_mean=reduce_mean, continuous=continuous,
                                   likelihood_weighting=likelihood_weighting, eps=eps)
