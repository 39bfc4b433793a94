"""Log-likelihood through the probability-flow ODE (reference: likelihood.py:26-113).

Same `get_div_fn` / `get_likelihood_fn` signatures and return values as the reference. The reference evaluates the
score network twice per ODE right-hand side (once for the drift, once more under autograd for the Hutchinson term,
likelihood.py:58-67,91-96); here one evaluation of the engine's autograd node gives the drift and, through the planned
backward pass w.r.t. the INPUT only (engine_train.TrainPlan(want_params=False, want_input=True)), the vector-Jacobian
product eps^T d(drift)/dx. The black-box integrator is scipy's, as in the reference (`method` is passed through); the
state crosses the host once per right-hand side like it does there.

`likelihood_fn(model, data, epsilon=None)`: `epsilon` optionally injects the Hutchinson probe (parity tests).
"""
import numpy as np
import torch
from scipy import integrate

from .models import utils as mutils


def get_div_fn(fn):
    """Hutchinson-Skilling divergence estimator of `fn` (likelihood.py:26-37)."""

    def div_fn(x, t, eps):
        with torch.enable_grad():
            x.requires_grad_(True)
            fn_eps = torch.sum(fn(x, t) * eps)
            grad_fn_eps = torch.autograd.grad(fn_eps, x)[0]
        x.requires_grad_(False)
        return torch.sum(grad_fn_eps * eps, dim=tuple(range(1, len(x.shape))))

    return div_fn


def get_likelihood_fn(sde, inverse_scaler, hutchinson_type="Rademacher", rtol=1e-5, atol=1e-5, method="RK45", eps=1e-5,
                      device_integrator=True):
    """likelihood.py:40-113. Returns likelihood_fn(model, data) -> (bpd [B], z, nfe). method='RK45' on CUDA data runs
    `ode.solve_rk45` (scipy's RK45 restated with the state resident in HBM) unless device_integrator=False."""

    def drift_fn(model, x, t):
        score_fn = mutils.get_score_fn(sde, model, train=False, continuous=True)
        rsde = sde.reverse(score_fn, probability_flow=True)
        return rsde.sde(x, t)[0]

    def drift_and_div(model, x, t, noise):
        """One network evaluation: drift and eps^T (d drift / dx) eps."""
        frozen = [p for p in model.parameters() if p.requires_grad]
        for p in frozen:                       # the divergence needs the input gradient only
            p.requires_grad_(False)
        try:
            with torch.enable_grad():
                x = x.detach().requires_grad_(True)
                drift = drift_fn(model, x, t)
                grad = torch.autograd.grad(torch.sum(drift * noise), x)[0]
        finally:
            for p in frozen:
                p.requires_grad_(True)
        return drift.detach(), torch.sum(grad * noise, dim=tuple(range(1, x.dim())))

    def likelihood_fn(model, data, epsilon=None):
        with torch.no_grad():
            shape = data.shape
            if epsilon is not None:
                noise = epsilon.to(data.device, torch.float32)
            elif hutchinson_type == "Gaussian":
                noise = torch.randn_like(data)
            elif hutchinson_type == "Rademacher":
                noise = torch.randint_like(data, low=0, high=2).float() * 2 - 1.0
            else:
                raise NotImplementedError(f"Hutchinson type {hutchinson_type} unknown.")

            def ode_func(t, x):
                sample = mutils.from_flattened_numpy(x[:-shape[0]], shape).to(data.device).type(torch.float32)
                vec_t = torch.ones(sample.shape[0], device=sample.device) * t
                drift, logp_grad = drift_and_div(model, sample, vec_t, noise)
                return np.concatenate([mutils.to_flattened_numpy(drift), mutils.to_flattened_numpy(logp_grad)], axis=0)

            if device_integrator and method == "RK45" and data.is_cuda:
                from . import ode
                n = data.numel()
                out = torch.empty(n + shape[0], device=data.device, dtype=torch.float32)

                def rhs(t, yflat):
                    vec_t = torch.ones(shape[0], device=data.device) * t
                    drift, logp_grad = drift_and_div(model, yflat[:n].view(shape), vec_t, noise)
                    out[:n].copy_(drift.reshape(-1))
                    out[n:].copy_(logp_grad)
                    return out

                y0 = torch.cat([data.reshape(-1).float(), torch.zeros(shape[0], device=data.device)])
                yT, nfe = ode.solve_rk45(rhs, float(eps), float(sde.T), y0, rtol, atol)
                z = yT[:n].view(shape).clone()
                delta_logp = yT[n:].clone()
            else:
                init = np.concatenate([mutils.to_flattened_numpy(data), np.zeros((shape[0],))], axis=0)
                solution = integrate.solve_ivp(ode_func, (eps, sde.T), init, rtol=rtol, atol=atol, method=method)
                nfe = solution.nfev
                zp = solution.y[:, -1]
                z = mutils.from_flattened_numpy(zp[:-shape[0]], shape).to(data.device).type(torch.float32)
                delta_logp = mutils.from_flattened_numpy(zp[-shape[0]:], (shape[0],)).to(data.device).type(torch.float32)
            prior_logp = sde.prior_logp(z)
            bpd = -(prior_logp + delta_logp) / np.log(2)
            bpd = bpd / np.prod(shape[1:])
            offset = 7.0 - inverse_scaler(-1.0)      # the reference's conversion to bits/dim (likelihood.py:108-110)
            return bpd + offset, z, nfe

    return likelihood_fn
