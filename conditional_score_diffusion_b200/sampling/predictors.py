"""Predictor classes and registry (reference: sampling/predictors.py:6-200).

Same registry keys and `update_fn` signatures. Each update evaluates the score function and then
applies ONE fused CUDA kernel (csd_reverse_diffusion_update_f32 / csd_euler_maruyama_update_f32)
instead of the reference's 5-8 elementwise ATen ops; per-sample coefficients (f, G, drift, g) are the
[B]-vector outputs of the SDE's own `discretize` / `sde` methods.
"""
import abc

import torch

from .. import kernels as K
from .. import sde_lib

_PREDICTORS = {}


def register_predictor(cls=None, *, name=None):
    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _PREDICTORS:
            raise ValueError(f"Already registered model with name: {local_name}")
        _PREDICTORS[local_name] = cls
        return cls

    return _register if cls is None else _register(cls)


def get_predictor(name):
    return _PREDICTORS[name]


def _is_ve(sde):
    return isinstance(sde, (sde_lib.VESDE, sde_lib.cVESDE))


def _prep(x):
    if x.device.type != "cuda":
        raise RuntimeError("predictor/corrector updates run on CUDA tensors only (libcsd_b200)")
    return x.contiguous().float()


class Predictor(abc.ABC):
    """The abstract class for a predictor algorithm (predictors.py:30-50)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        self.rsde = sde.reverse(score_fn, probability_flow)
        self.score_fn = score_fn
        self.probability_flow = probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t):
        pass


class _ReverseDiffusion:
    """x_mean = x - (f - G^2 score), x = x_mean + G z (predictors.py:79-102, sde_lib.py:87-92)."""

    def _update(self, x, t, *cond):
        x = _prep(x)
        score = self.score_fn(x, *cond, t)
        z = torch.randn_like(x)
        f, G = self.sde.discretize(torch.ones_like(t), t)  # per-sample linear coefficient and G
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        K.reverse_diffusion_update(x, score.contiguous(), z, x_out, x_mean, None if _is_ve(self.sde) else f.contiguous(),
                                   G.contiguous(), self.probability_flow, None, 1)
        return x_out, x_mean


class _EulerMaruyama:
    """x_mean = x + drift dt, x = x_mean + g sqrt(-dt) z (predictors.py:52-77, sde_lib.py:78-85)."""

    def _update(self, x, t, *cond):
        x = _prep(x)
        dt = -1.0 / self.rsde.N
        z = torch.randn_like(x)
        score = self.score_fn(x, *cond, t)
        d, g = self.sde.sde(torch.ones_like(t), t)
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        K.euler_maruyama_update(x, score.contiguous(), z, x_out, x_mean, None if _is_ve(self.sde) else d.contiguous(),
                                g.contiguous(), dt, self.probability_flow, None, 1)
        return x_out, x_mean


@register_predictor(name="euler_maruyama")
class EulerMaruyamaPredictor(_EulerMaruyama, Predictor):
    def update_fn(self, x, t):
        return self._update(x, t)


@register_predictor(name="conditional_euler_maruyama")
class conditionalEulerMaruyamaPredictor(_EulerMaruyama, Predictor):
    def update_fn(self, x, y, t):
        return self._update(x, t, y)


@register_predictor(name="reverse_diffusion")
class ReverseDiffusionPredictor(_ReverseDiffusion, Predictor):
    def update_fn(self, x, t):
        return self._update(x, t)


@register_predictor(name="conditional_reverse_diffusion")
class conditionalReverseDiffusionPredictor(_ReverseDiffusion, Predictor):
    def update_fn(self, x, y, t):
        return self._update(x, t, y)


class _Ancestral:
    """Ancestral sampling written as a reverse-diffusion-style update (predictors.py:105-144):
    VE: x_mean = x + (s^2 - s_adj^2) score, std = sqrt(s_adj^2 (s^2 - s_adj^2) / s^2);
    VP: x_mean = (x + beta score) / sqrt(1 - beta), std = sqrt(beta)."""

    def _update(self, x, t, *cond):
        x = _prep(x)
        sde = self.sde
        timestep = (t * (sde.N - 1) / sde.T).long()
        score = self.score_fn(x, *cond, t).contiguous()
        noise = torch.randn_like(x)
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        if _is_ve(sde):
            sig = sde._on("discrete_sigmas", t.device)
            sigma = sig[timestep]
            adj = torch.where(timestep == 0, torch.zeros_like(t), sig[timestep - 1])
            g = torch.sqrt(sigma ** 2 - adj ** 2)                       # x_mean = x + g^2 score
            std = torch.sqrt((adj ** 2 * (sigma ** 2 - adj ** 2)) / (sigma ** 2))
            # the kernel's probability-flow form applies g^2/2 and no noise: pass sqrt(2) g
            K.reverse_diffusion_update(x, score, None, x_out, x_mean, None, (g * 2 ** 0.5).contiguous(), True, None, 1)
            K.ve_perturb(x_mean, noise, x_out, std.contiguous(), None, 1)
        else:
            beta = sde._on("discrete_betas", t.device)[timestep]
            inv = 1.0 / torch.sqrt(1.0 - beta)
            # x_mean = inv*x + inv*beta*score = x - ((1 - inv) x - (sqrt(inv*beta))^2 score)
            K.reverse_diffusion_update(x, score, None, x_out, x_mean, (1.0 - inv).contiguous(),
                                       torch.sqrt(2.0 * inv * beta).contiguous(), True, None, 1)
            K.ve_perturb(x_mean, noise, x_out, torch.sqrt(beta).contiguous(), None, 1)
        return x_out, x_mean


@register_predictor(name="ancestral_sampling")
class AncestralSamplingPredictor(_Ancestral, Predictor):
    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__(sde, score_fn, probability_flow)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        assert not probability_flow, "Probability flow not supported by ancestral sampling"

    def update_fn(self, x, t):
        return self._update(x, t)


@register_predictor(name="conditional_ancestral_sampling")
class conditionalAncestralSamplingPredictor(_Ancestral, Predictor):
    """The reference's update_fn has the wrong arity and returns None (predictors.py:175-179,
    SURVEY.md §2); this one takes (x, y, t) like the other conditional predictors."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__(sde, score_fn, probability_flow)
        if not isinstance(sde, (sde_lib.cVESDE, sde_lib.cVPSDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        assert not probability_flow, "Probability flow not supported by ancestral sampling"

    def update_fn(self, x, y, t):
        return self._update(x, t, y)


@register_predictor(name="none")
class NonePredictor(Predictor):
    """An empty predictor that does nothing (predictors.py:182-190)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        pass

    def update_fn(self, x, t):
        return x, x


@register_predictor(name="conditional_none")
class NonePredictor(Predictor):  # noqa: F811 - the reference shadows the name the same way (predictors.py:192-200)
    def __init__(self, sde, score_fn, probability_flow=False):
        pass

    def update_fn(self, x, y, t):
        return x, x
