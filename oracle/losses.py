"""Oracle restatement of the denoising score-matching losses (CPU, fp32). TEST INFRASTRUCTURE ONLY.

Follows losses.py:99-234 with the random draws (t, z) passed in, so the RNG is factored out of the comparison."""
import torch


def _bc(v, like):
    return v[(...,) + (None,) * (like.ndim - 1)]


def _reduce(losses, reduce_mean):
    flat = losses.reshape(losses.shape[0], -1)
    return flat.mean(dim=-1) if reduce_mean else 0.5 * flat.sum(dim=-1)


def sr3_loss(score_fn, sde, y, x, t, z, reduce_mean=True, likelihood_weighting=True):
    """SR3 branch (losses.py:185-206): score_fn({'x': perturbed x, 'y': y}, t) -> score of x. `sde` is an oracle.sde
    object (sigma(t) / diffusion(t))."""
    std = sde.sigma(t)
    perturbed = x + _bc(std, x) * z
    score = score_fn({"x": perturbed, "y": y}, t)
    if likelihood_weighting:
        g2 = sde.diffusion(t) ** 2
        losses = _reduce(torch.square(score + z / _bc(std, x)), reduce_mean) * g2
    else:
        losses = _reduce(torch.square(score * _bc(std, x) + z), reduce_mean)
    return losses.mean()


def cmde_loss(score_fn, sde_x, sde_y, y, x, t, z_x, z_y, reduce_mean=True):
    """Two-SDE branch (losses.py:119-146): both x and y are perturbed, the two residuals are concatenated."""
    std_y, std_x = sde_y.sigma(t), sde_x.sigma(t)
    pert = {"x": x + _bc(std_x, x) * z_x, "y": y + _bc(std_y, y) * z_y}
    score = score_fn(pert, t)
    g2_y, g2_x = sde_y.diffusion(t) ** 2, sde_x.diffusion(t) ** 2
    ly = (torch.square(score["y"] + z_y / _bc(std_y, y)) * _bc(g2_y, y)).reshape(y.shape[0], -1)
    lx = (torch.square(score["x"] + z_x / _bc(std_x, x)) * _bc(g2_x, x)).reshape(x.shape[0], -1)
    cat = torch.cat((lx, ly), dim=-1)
    per = cat.mean(dim=-1) if reduce_mean else 0.5 * cat.sum(dim=-1)
    return per.mean()


def uncond_loss(score_fn, sde, x, t, z, reduce_mean=True, likelihood_weighting=False):
    """Unconditional branch (losses.py:207-232)."""
    std = sde.sigma(t)
    score = score_fn(x + _bc(std, x) * z, t)
    if likelihood_weighting:
        g2 = sde.diffusion(t) ** 2
        losses = _reduce(torch.square(score + z / _bc(std, x)), reduce_mean) * g2
    else:
        losses = _reduce(torch.square(score * _bc(std, x) + z), reduce_mean)
    return losses.mean()
