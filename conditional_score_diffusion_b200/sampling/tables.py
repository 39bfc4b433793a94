"""Per-step scalar tables for the fused predictor-corrector loop.

Every scalar the reference recomputes on the host at each step (time labels, sigma(t), G(t), alpha,
1/std; sampling/conditional.py:104-110, sde_lib.py:186-194,310-321,349-360, models/utils.py:172-262)
is evaluated ONCE here, with the same torch fp32 formulas on the same `torch.linspace` time grid,
and uploaded as small device tables that the CUDA update kernels index with a device-resident step
counter.
"""
import torch

from .. import sde_lib


def _is_ve(sde):
    return isinstance(sde, (sde_lib.VESDE, sde_lib.cVESDE))


def _is_vp(sde):
    return isinstance(sde, (sde_lib.VPSDE, sde_lib.cVPSDE))


def time_grid(sde, eps, p_steps):
    return torch.linspace(sde.T, eps, p_steps)


def model_time_tables(sde, timesteps, continuous, conditional, embedding_type="positional"):
    """(labels, inv_std) per step: what the score wrappers of models/utils.py:156-267 feed the network
    and divide its output by. sde may be a {'x','y'} dict (conditional pair) -> inv_std is a dict."""
    t = timesteps
    if isinstance(sde, dict):
        sx = sde["x"]
        if continuous:
            labels = t * (sx.N - 1)
            inv = {k: 1.0 / sde[k].marginal_prob(t, t)[1] for k in ("x", "y")}
        else:
            lab = torch.round((t * (sx.N - 1)).float()).long()
            labels = lab.float()
            inv = {k: 1.0 / sde[k].discrete_sigmas[lab] for k in ("x", "y")}
        return labels, inv
    if _is_ve(sde):
        if conditional:
            if continuous:
                return t * (sde.N - 1), 1.0 / sde.marginal_prob(t, t)[1]
            lab = torch.round((t * (sde.N - 1)).float()).long()
            return lab.float(), 1.0 / sde.discrete_sigmas[lab]
        if continuous:
            std = sde.marginal_prob(t, t)[1]
            return (torch.log(std) if embedding_type == "fourier" else std), 1.0 / std
        lab = torch.round(t * (sde.N - 1)).long()
        std = sde.discrete_sigmas[lab]
        return std, 1.0 / std
    if _is_vp(sde) or isinstance(sde, sde_lib.subVPSDE):
        labels = t * (sde.N - 1)
        if continuous or isinstance(sde, sde_lib.subVPSDE):
            std = sde.marginal_prob(t, t)[1]
        else:
            std = sde.sqrt_1m_alphas_cumprod[labels.long()]
        return labels, 1.0 / std
    raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")


def predictor_tables(sde, timesteps, kind):
    """Coefficients of one predictor step. kind: 'reverse_diffusion' | 'euler_maruyama'.
    Returns (lin_coef or None, g): rev_f = lin*x - g^2*score (RD) / drift = lin*x - g^2*score (EM)."""
    t = timesteps
    dummy = torch.zeros_like(t)
    if kind == "reverse_diffusion":
        f, G = sde.discretize(torch.ones_like(t), t)   # x = 1 -> f is the linear coefficient
        return (None if _is_ve(sde) else f), G
    if kind == "euler_maruyama":
        drift, g = sde.sde(torch.ones_like(t), t)
        return (None if _is_ve(sde) else drift), g
    raise NotImplementedError(kind)


def langevin_alpha_table(sde, timesteps):
    """alpha of LangevinCorrector.update_fn (sampling/correctors.py:64-68): 1 for VE."""
    if _is_vp(sde) or isinstance(sde, sde_lib.subVPSDE):
        step = (timesteps * (sde.N - 1) / sde.T).long()
        return sde.alphas[step]
    return None
