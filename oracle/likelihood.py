"""Oracle restatement of the probability-flow likelihood and ODE sampler (CPU, fp32). TEST INFRASTRUCTURE ONLY.

Follows likelihood.py:26-113 and sampling/unconditional.py:93-158 for a VE SDE: drift = -1/2 g(t)^2 score(x, t), the
Hutchinson term from autograd, scipy.integrate.solve_ivp as the integrator. The Hutchinson probe is passed in."""
import numpy as np
import torch
from scipy import integrate


def _drift(score_fn, sde, x, t):
    g = sde.diffusion(t)
    return -0.5 * (g ** 2)[:, None, None, None] * score_fn(x, t)


def likelihood(score_fn, sde, data, epsilon, inverse_scaler, rtol=1e-5, atol=1e-5, eps=1e-5, method="RK45"):
    shape = data.shape

    def ode_func(t, x):
        sample = torch.from_numpy(x[:-shape[0]].reshape(shape)).float()
        vec_t = torch.ones(shape[0]) * t
        with torch.no_grad():
            drift = _drift(score_fn, sde, sample, vec_t)
        with torch.enable_grad():
            xs = sample.clone().requires_grad_(True)
            grad = torch.autograd.grad(torch.sum(_drift(score_fn, sde, xs, vec_t) * epsilon), xs)[0]
        div = torch.sum(grad * epsilon, dim=(1, 2, 3))
        return np.concatenate([drift.numpy().reshape(-1), div.detach().numpy().reshape(-1)], axis=0)

    init = np.concatenate([data.numpy().reshape(-1), np.zeros((shape[0],))], axis=0)
    sol = integrate.solve_ivp(ode_func, (eps, 1.0), init, rtol=rtol, atol=atol, method=method)
    zp = sol.y[:, -1]
    z = torch.from_numpy(zp[:-shape[0]].reshape(shape)).float()
    delta_logp = torch.from_numpy(zp[-shape[0]:]).float()
    n = np.prod(shape[1:])
    prior_logp = -n / 2.0 * np.log(2 * np.pi * sde.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * sde.sigma_max ** 2)
    bpd = -(prior_logp + delta_logp) / np.log(2) / n + (7.0 - inverse_scaler(-1.0))
    return bpd, z, sol.nfev


def ode_sampler(score_fn, sde, z, rtol=1e-5, atol=1e-5, eps=1e-3, method="RK45"):
    shape = z.shape

    def ode_func(t, x):
        xs = torch.from_numpy(x.reshape(shape)).float()
        with torch.no_grad():
            return _drift(score_fn, sde, xs, torch.ones(shape[0]) * t).numpy().reshape(-1)

    sol = integrate.solve_ivp(ode_func, (1.0, eps), z.numpy().reshape(-1), rtol=rtol, atol=atol, method=method)
    return torch.tensor(sol.y[:, -1]).reshape(shape).float(), sol.nfev
