"""Host logic of likelihood.get_likelihood_fn and sampling.get_ode_sampler on a problem with a closed-form answer.

Data ~ N(0, s^2 I) has perturbed score -x / (s^2 + sigma(t)^2); the probability-flow ODE of the VE SDE is then linear,
its divergence is exact under Rademacher probes, and the log-likelihood of a point is the Gaussian log-density with
variance s^2 + sigma(eps)^2. The stand-in model is plain torch (the engine-backed networks are CUDA-only and are
covered by tests/test_gpu_training.py); what is tested here is the integrator plumbing, the divergence estimator and
the bits/dim conversion (likelihood.py:26-113), and the ODE sampler's direction / denoising step."""
import math

import numpy as np
import torch

from conditional_score_diffusion_b200 import likelihood, sde_lib
from conditional_score_diffusion_b200.sampling import unconditional


class GaussianScore(torch.nn.Module):
    embedding_type = "positional"     # models.utils.get_score_fn then passes sigma(t) as the label

    def __init__(self, s):
        super().__init__()
        self.s = s
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    @property
    def device(self):
        return self.dummy.device

    def forward(self, x, labels):     # network output = score * sigma (the wrapper divides by sigma)
        std = labels.view(-1, 1, 1, 1)
        return -x * std / (self.s ** 2 + std ** 2)


def test_bits_per_dim_of_a_gaussian():
    s, shape = 1.5, (3, 2, 4, 4)
    sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=1000)
    torch.manual_seed(0)
    data = torch.randn(*shape) * s
    fn = likelihood.get_likelihood_fn(sde, lambda v: v, rtol=1e-6, atol=1e-6, eps=1e-5)
    bpd, z, nfe = fn(GaussianScore(s), data)
    n = np.prod(shape[1:])
    var0 = s ** 2 + (0.01 * (50 / 0.01) ** 1e-5) ** 2
    logp = -0.5 * n * math.log(2 * math.pi * var0) - data.pow(2).sum(dim=(1, 2, 3)) / (2 * var0)
    expect = -logp / math.log(2) / n + 8.0          # offset = 7 - inverse_scaler(-1) = 8 for the identity scaler
    print("bpd", bpd.tolist(), "expected", expect.tolist(), "nfe", nfe)
    assert torch.allclose(bpd, expect.float(), atol=5e-3)
    # the latent is the data scaled to the prior's standard deviation
    ratio = (z.std() / data.std()).item()
    assert abs(ratio - math.sqrt(s ** 2 + 50 ** 2) / math.sqrt(var0)) < 0.05 * ratio
    assert nfe > 10


def test_div_fn_matches_exact_divergence():
    fn = lambda x, t: 3.0 * x + x ** 2
    x = torch.randn(2, 1, 3, 3)
    eps = torch.randint(0, 2, x.shape).float() * 2 - 1
    div = likelihood.get_div_fn(fn)(x, torch.zeros(2), eps)
    assert torch.allclose(div, (3.0 + 2 * x).sum(dim=(1, 2, 3)), atol=1e-5)
    assert not x.requires_grad


def test_ode_sampler_maps_prior_to_data_scale():
    s, shape = 1.5, (4, 1, 8, 8)
    sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=1000)
    # denoise=False: the one-step denoiser is a libcsd_b200 kernel (CUDA only; tests/test_gpu_training.py runs it)
    sampler = unconditional.get_ode_sampler(sde, shape, denoise=False, rtol=1e-5, atol=1e-5, eps=1e-3)
    torch.manual_seed(1)
    z = torch.randn(*shape) * math.sqrt(s ** 2 + 50 ** 2)
    x, nfe = sampler(GaussianScore(s), z=z.clone())
    # linear flow: x(eps) = z * sqrt((s^2 + sigma(eps)^2) / (s^2 + sigma_max^2))
    expect = z * math.sqrt((s ** 2 + (0.01 * 5000 ** 1e-3) ** 2) / (s ** 2 + 50 ** 2))
    assert torch.allclose(x, expect, rtol=2e-3, atol=2e-3), (x - expect).abs().max()
    assert nfe > 10
