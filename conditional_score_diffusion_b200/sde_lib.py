"""SDE definitions with the reference's class names and methods (reference: sde_lib.py:7-418).

Per-sample scalar maths only ([B]-shaped tensors); the heavy elementwise work of sampling is done by
the fused CUDA update kernels, which take these scalars as small device tables
(`sampling.tables`). Differences from the reference, none of which change results:
  * lookup tables (discrete_sigmas, alphas, ...) are moved to the target device once and cached
    instead of `.to(device)` on every call (sde_lib.py:188-189, 357-359);
  * the conditional variants (cSDE / cVESDE / cVPSDE) share their maths with the unconditional
    classes through mixins instead of duplicating it.
"""
import abc

import numpy as np
import torch


def _bc(v, x):
    """Broadcast a per-sample vector over the trailing dims of x."""
    return v[(...,) + (None,) * len(x.shape[1:])]


class SDE(abc.ABC):
    """SDE abstract class (sde_lib.py:7-101)."""

    def __init__(self, N):
        super().__init__()
        self.N = N
        self._dev_cache = {}

    def _on(self, name, device):
        """Cached device copy of a lookup table attribute."""
        key = (name, str(device))
        t = self._dev_cache.get(key)
        if t is None:
            t = getattr(self, name).to(device)
            self._dev_cache[key] = t
        return t

    @property
    @abc.abstractmethod
    def T(self):
        pass

    @abc.abstractmethod
    def sde(self, x, t):
        pass

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        pass

    @abc.abstractmethod
    def prior_sampling(self, shape):
        pass

    @abc.abstractmethod
    def prior_logp(self, z):
        pass

    def discretize(self, x, t):
        """Euler-Maruyama discretisation x_{i+1} = x_i + f_i + G_i z_i (sde_lib.py:49-63)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def _reverse_class(self, score_call, probability_flow):
        """Build the reverse-time SDE object (sde_lib.py:65-101 / 104-142)."""
        N, T = self.N, self.T
        fwd_sde, fwd_discretize = self.sde, self.discretize
        scale = 0.5 if probability_flow else 1.0

        class RSDE(self.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def sde(self, x, *cond_and_t):
                t = cond_and_t[-1]
                drift, diffusion = fwd_sde(x, t)
                score = score_call(x, *cond_and_t)
                drift = drift - _bc(diffusion, x) ** 2 * score * scale
                return drift, (0.0 if probability_flow else diffusion)

            def discretize(self, x, *cond_and_t):
                t = cond_and_t[-1]
                f, G = fwd_discretize(x, t)
                rev_f = f - _bc(G, x) ** 2 * score_call(x, *cond_and_t) * scale
                return rev_f, (torch.zeros_like(G) if probability_flow else G)

        return RSDE()

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE/ODE with score_fn(x, t) (sde_lib.py:65-101)."""
        return self._reverse_class(score_fn, probability_flow)


class cSDE(SDE):
    """Conditional setting: the reverse SDE takes (x, y, t) and score_fn(x, y, t) (sde_lib.py:104-142)."""

    def reverse(self, score_fn, probability_flow=False):
        return self._reverse_class(score_fn, probability_flow)


class _VPMath:
    def _vp_init(self, beta_min, beta_max, N):
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * _bc(beta_t, x) * x, torch.sqrt(beta_t)

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(_bc(log_mean_coeff, x)) * x, torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.0

    def discretize(self, x, t):
        """DDPM discretisation (sde_lib.py:186-194)."""
        timestep = (t * (self.N - 1) / self.T).long()
        beta = self._on("discrete_betas", x.device)[timestep]
        alpha = self._on("alphas", x.device)[timestep]
        return _bc(torch.sqrt(alpha), x) * x - x, torch.sqrt(beta)


class VPSDE(_VPMath, SDE):
    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        SDE.__init__(self, N)
        self._vp_init(beta_min, beta_max, N)


class cVPSDE(_VPMath, cSDE):
    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        cSDE.__init__(self, N)
        self._vp_init(beta_min, beta_max, N)


class subVPSDE(SDE):
    """sde_lib.py:251-287."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1.0 - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * _bc(beta_t, x) * x, torch.sqrt(beta_t * discount)

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return _bc(torch.exp(log_mean_coeff), x) * x, 1 - torch.exp(2.0 * log_mean_coeff)

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.0


class _VEMath:
    def _ve_init(self, sigma_min, sigma_max, N, data_mean):
        self.sigma_min, self.sigma_max, self.N = sigma_min, sigma_max, N
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))
        self.diffused_mean = data_mean

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        g = sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min))).type_as(t))
        return torch.zeros_like(x), g

    def marginal_prob(self, x, t):
        smin = torch.tensor(self.sigma_min).type_as(t)
        smax = torch.tensor(self.sigma_max).type_as(t)
        return x, smin * (smax / smin) ** t

    def compute_backward_kernel(self, x0, x_tplustau, t, tau):
        """Parameters of p(x(t) | x(0), x(t+tau)) (sde_lib.py:323-339)."""
        smin, smax = torch.tensor(self.sigma_min).type_as(t), torch.tensor(self.sigma_max).type_as(t)
        s_t = (smin * (smax / smin) ** t) ** 2
        s_tau = (smin * (smax / smin) ** (t + tau)) ** 2
        std_backward = torch.sqrt(s_t * (s_tau - s_t) / s_tau)
        mean = x0 * _bc((s_tau - s_t) / s_tau, x0) + x_tplustau * _bc(s_t / s_tau, x0)
        return mean, std_backward

    def prior_sampling(self, shape):
        z = torch.randn(*shape) * self.sigma_max
        if self.diffused_mean is not None:
            z = z + self.diffused_mean.unsqueeze(0).repeat(tuple([shape[0]] + [1] * (len(shape) - 1)))
        return z

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (
            2 * self.sigma_max ** 2)

    def discretize(self, x, t):
        """SMLD discretisation (sde_lib.py:349-360): f = 0, G = sqrt(sigma_i^2 - sigma_{i-1}^2)."""
        timestep = (t * (self.N - 1) / self.T).long()
        sig = self._on("discrete_sigmas", t.device)
        sigma = sig[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(t), sig[timestep - 1])
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)


class VESDE(_VEMath, SDE):
    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, data_mean=None):
        SDE.__init__(self, N)
        self._ve_init(sigma_min, sigma_max, N, data_mean)


class cVESDE(_VEMath, cSDE):
    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, data_mean=None):
        cSDE.__init__(self, N)
        self._ve_init(sigma_min, sigma_max, N, data_mean)
