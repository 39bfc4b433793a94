"""Oracle restatement of the predictor / corrector updates and the PC sampling loops (CPU fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Noise is drawn through `randn_like`, by default torch's global generator in exactly the order the
reference draws it, so that a seeded oracle run is comparable with a seeded reference run; tests
that compare against the CUDA path inject recorded noise instead.
"""
import torch

from .sde import _bc


def reverse_diffusion_update(sde, score, x, t, z, probability_flow=False):
    """ReverseDiffusionPredictor.update_fn (sampling/predictors.py:79-92) with RSDE.discretize
    (sde_lib.py:87-92). `score` is score_fn(x, t) already evaluated. Returns (x, x_mean)."""
    if sde.kind == "ve":
        f = torch.zeros_like(x)
        g = sde.discretize_g(t)
    else:
        fc, g = sde.discretize_fg(t)
        f = _bc(fc, x) * x
    rev_f = f - _bc(g, x) ** 2 * score * (0.5 if probability_flow else 1.0)
    rev_g = torch.zeros_like(g) if probability_flow else g
    x_mean = x - rev_f
    return x_mean + _bc(rev_g, x) * z, x_mean


def euler_maruyama_update(sde, score, x, t, z, probability_flow=False):
    """EulerMaruyamaPredictor.update_fn (sampling/predictors.py:52-63) with RSDE.sde (sde_lib.py:78-85)."""
    dt = -1.0 / sde.N
    g = sde.diffusion(t)
    drift = torch.zeros_like(x) if sde.kind == "ve" else -0.5 * _bc(sde.beta(t), x) * x
    drift = drift - _bc(g, x) ** 2 * score * (0.5 if probability_flow else 1.0)
    g = torch.zeros_like(g) if probability_flow else g
    x_mean = x + drift * dt
    return x_mean + _bc(g, x) * (-dt) ** 0.5 * z, x_mean


def langevin_update(sde, grad, x, t, z, snr):
    """One inner iteration of LangevinCorrector.update_fn (sampling/correctors.py:58-78):
    batch-mean norms, step = 2 alpha (snr |z| / |g|)^2."""
    if sde.kind == "vp":
        alpha = sde.alphas[(t * (sde.N - 1) / sde.T).long()]
    else:
        alpha = torch.ones_like(t)
    grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(z.reshape(z.shape[0], -1), dim=-1).mean()
    step = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
    x_mean = x + _bc(step, x) * grad
    return x_mean + _bc(torch.sqrt(step * 2), x) * z, x_mean


def pc_sampler(score_fn, sde, shape, snr, p_steps, c_steps=1, eps=1e-3, denoise=True,
               predictor="reverse_diffusion", corrector="langevin", probability_flow=False,
               randn_like=torch.randn_like, x_init=None, record=None):
    """get_pc_sampler(...)(model) (sampling/unconditional.py:161-228): corrector then predictor."""
    x = sde.prior_sampling(shape).float() if x_init is None else x_init.clone()
    timesteps = torch.linspace(sde.T, eps, p_steps)
    x_mean = x
    for i in range(p_steps):
        vec_t = torch.ones(shape[0]) * timesteps[i]
        if corrector == "langevin":
            for _ in range(c_steps):
                grad = score_fn(x, vec_t)
                z = randn_like(x)
                x, x_mean = langevin_update(sde, grad, x, vec_t, z, snr)
        if predictor == "reverse_diffusion":
            score = score_fn(x, vec_t)     # rsde.discretize evaluates the score first (predictors.py:86)
            z = randn_like(x)
            x, x_mean = reverse_diffusion_update(sde, score, x, vec_t, z, probability_flow)
        elif predictor == "euler_maruyama":
            z = randn_like(x)              # predictors.py:59 draws z before the score
            score = score_fn(x, vec_t)
            x, x_mean = euler_maruyama_update(sde, score, x, vec_t, z, probability_flow)
        if record is not None:
            record.append(x.clone())
    return (x_mean if denoise else x), {"times": timesteps, "steps": p_steps * (c_steps + 1)}


def pc_conditional_sampler(score_fn, sde_x, sde_y, y, shape, snr, p_steps, c_steps=1, eps=1e-5,
                           denoise=True, randn_like=torch.randn_like, x_init=None, record=None):
    """get_pc_conditional_sampler(..., use_path=False)(model, y) for sde = {'x','y'}
    (sampling/conditional.py:104-110,180-226): before every corrector and predictor call the
    condition is re-noised, y_t = y + sigma_y(t) z; conditional Langevin then conditional
    reverse diffusion. score_fn(x, y_t, t) -> x-score."""
    x = sde_x.prior_sampling(shape) if x_init is None else x_init.clone()
    timesteps = torch.linspace(sde_x.T, eps, p_steps)
    x_mean = x
    for i in range(p_steps):
        vec_t = torch.ones(x.shape[0]) * timesteps[i]
        # corrector (conditional.py:104-110 then correctors.py:88-108)
        y_t = y + randn_like(y) * _bc(sde_y.sigma(vec_t), y)
        for _ in range(c_steps):
            grad = score_fn(x, y_t, vec_t)
            z = randn_like(x)
            x, x_mean = langevin_update(sde_x, grad, x, vec_t, z, snr)
        # predictor (fresh y_t)
        y_t = y + randn_like(y) * _bc(sde_y.sigma(vec_t), y)
        score = score_fn(x, y_t, vec_t)
        z = randn_like(x)
        x, x_mean = reverse_diffusion_update(sde_x, score, x, vec_t, z)
        if record is not None:
            record.append(x.clone())
    return (x_mean if denoise else x), {}


def pc_conditional_sampler_path(score_fn, sde_x, sde_y, y, shape, snr, p_steps, eps=1e-5, denoise=True,
                                randn_like=torch.randn_like, x_init=None, record=None):
    """get_pc_conditional_sampler(..., use_path=True)(model, y) (sampling/conditional.py:87-94,124-176): y follows one
    path through p(y_t | y_0, y_{t+tau}) (VESDE.compute_backward_kernel, sde_lib.py:323-339); predictor, then corrector."""
    x = sde_x.prior_sampling(shape) if x_init is None else x_init.clone()
    timesteps = torch.linspace(sde_x.T, eps, p_steps)
    tau = timesteps[0] - timesteps[1]
    ones = torch.ones(x.shape[0])
    y_tpt = y + randn_like(y) * _bc(sde_y.sigma(ones * (timesteps[0] + tau)), y)
    x_mean = x
    for i in range(p_steps):
        vec_t = ones * timesteps[i]
        s_t, s_tau = sde_y.sigma(vec_t) ** 2, sde_y.sigma(vec_t + tau) ** 2
        mean = y * _bc((s_tau - s_t) / s_tau, y) + y_tpt * _bc(s_t / s_tau, y)
        y_tpt = mean + randn_like(y) * _bc(torch.sqrt(s_t * (s_tau - s_t) / s_tau), y)
        score = score_fn(x, y_tpt, vec_t)
        x, x_mean = reverse_diffusion_update(sde_x, score, x, vec_t, randn_like(x))
        grad = score_fn(x, y_tpt, vec_t)
        x, x_mean = langevin_update(sde_x, grad, x, vec_t, randn_like(x), snr)
        if record is not None:
            record.append(x.clone())
    return (x_mean if denoise else x), {}


def pc_inpainter(score_fn, sde, data, mask, snr, eps=1e-5, denoise=True, randn_like=torch.randn_like, x_init=None,
                 record=None):
    """get_pc_inpainter(...)(model, data, mask) (sampling/unconditional.py:230-345) for a VE SDE with the Langevin
    corrector and the reverse-diffusion predictor: after each update the known pixels are re-imposed at the current
    noise level; x_mean is rebuilt from the merged x (reference quirk, :275)."""
    prior = sde.prior_sampling(data.shape) if x_init is None else x_init.clone()
    x = data * mask + prior * (1.0 - mask)
    timesteps = torch.linspace(sde.T, eps, sde.N)
    x_mean = x

    def merge(x, vec_t):
        std = sde.sigma(vec_t)
        masked = data + randn_like(x) * _bc(std, x)
        x = x * (1.0 - mask) + masked * mask
        return x, x * (1.0 - mask) + data * mask

    for i in range(sde.N):
        vec_t = torch.ones(data.shape[0]) * timesteps[i]
        grad = score_fn(x, vec_t)
        x, _ = langevin_update(sde, grad, x, vec_t, randn_like(x), snr)
        x, x_mean = merge(x, vec_t)
        score = score_fn(x, vec_t)
        x, _ = reverse_diffusion_update(sde, score, x, vec_t, randn_like(x))
        x, x_mean = merge(x, vec_t)
        if record is not None:
            record.append(x.clone())
    return (x_mean if denoise else x), {}
