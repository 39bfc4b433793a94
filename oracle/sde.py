"""Oracle restatement of the SDE maths and the score wrappers (CPU fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math

import numpy as np
import torch


def _bc(v, x):
    return v[(...,) + (None,) * (x.ndim - 1)]


class VE:
    """VESDE / cVESDE (sde_lib.py:290-418): sigma(t) = smin (smax/smin)^t."""

    kind = "ve"

    def __init__(self, sigma_min=0.01, sigma_max=50.0, N=1000):
        self.sigma_min, self.sigma_max, self.N, self.T = sigma_min, sigma_max, N, 1
        # sde_lib.py:301 - note: exp(linspace(log smin, log smax, N)) in fp32
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    def sigma(self, t):
        """marginal_prob std (sde_lib.py:316-321)."""
        smin = torch.tensor(self.sigma_min).type_as(t)
        smax = torch.tensor(self.sigma_max).type_as(t)
        return smin * (smax / smin) ** t

    def diffusion(self, t):
        """g(t) of `sde` (sde_lib.py:310-314)."""
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min))).type_as(t))

    def discretize_g(self, t):
        """G of `discretize` (sde_lib.py:349-360); f = 0."""
        step = (t * (self.N - 1) / self.T).long()
        sigma = self.discrete_sigmas[step]
        adj = torch.where(step == 0, torch.zeros_like(t), self.discrete_sigmas[step - 1])
        return torch.sqrt(sigma ** 2 - adj ** 2)

    def prior_sampling(self, shape):
        return torch.randn(*shape) * self.sigma_max  # sde_lib.py:341-347 (no data mean)


class VP:
    """VPSDE / cVPSDE (sde_lib.py:144-249)."""

    kind = "vp"

    def __init__(self, beta_min=0.1, beta_max=20.0, N=1000):
        self.beta_0, self.beta_1, self.N, self.T = beta_min, beta_max, N, 1
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    def beta(self, t):
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def marginal(self, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(lmc), torch.sqrt(1.0 - torch.exp(2.0 * lmc))

    def sigma(self, t):
        return self.marginal(t)[1]

    def diffusion(self, t):
        return torch.sqrt(self.beta(t))

    def discretize_fg(self, t):
        """DDPM discretisation (sde_lib.py:186-194): f = (sqrt(alpha) - 1) x, G = sqrt(beta)."""
        step = (t * (self.N - 1) / self.T).long()
        beta = self.discrete_betas[step]
        alpha = self.alphas[step]
        return torch.sqrt(alpha) - 1.0, torch.sqrt(beta)

    def prior_sampling(self, shape):
        return torch.randn(*shape)


def score_fn_unconditional(model_fn, sde, continuous, embedding_type="positional"):
    """get_score_fn, unconditional branch (models/utils.py:227-262).

    model_fn(x, labels) -> network output.
    """
    def score(x, t):
        if sde.kind == "ve":
            if continuous:
                std = sde.sigma(t)
                emb = torch.log(std) if embedding_type == "fourier" else std
                return model_fn(x, emb) / _bc(std, x)
            labels = torch.round(t * (sde.N - 1)).long()
            std = sde.discrete_sigmas[labels]
            return model_fn(x, std) / _bc(std, x)
        labels = t * (sde.N - 1)
        out = model_fn(x, labels)
        std = sde.sigma(t) if continuous else sde.sqrt_1m_alphas_cumprod[labels.long()]
        return out / _bc(std, x)
    return score


def score_fn_conditional_pair(model_fn, sde_x, sde_y, continuous):
    """get_score_fn, conditional dict branch for {'x': cVESDE, 'y': VESDE}
    (models/utils.py:172-186) composed with get_conditional_score_fn(target 'x') (:270-278).

    model_fn({'x','y'}, labels) -> {'x','y'}; returns the x-score.
    """
    def score(x, y, t):
        if continuous:
            labels = t * (sde_x.N - 1)
            out = model_fn({"x": x, "y": y}, labels)
            return out["x"] / _bc(sde_x.sigma(t), x)
        labels = torch.round((t * (sde_x.N - 1)).float()).long()
        out = model_fn({"x": x, "y": y}, labels)
        return out["x"] / _bc(sde_x.discrete_sigmas[labels], x)
    return score
