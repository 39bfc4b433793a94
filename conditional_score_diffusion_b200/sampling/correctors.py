"""Corrector classes and registry (reference: sampling/correctors.py:5-164).

Langevin: the two per-sample norms come from one reduction kernel, the batch means are taken inside
the update kernel, and x_mean / x are written in one pass (the reference: ~12 launches + 2 reductions).
"""
import abc

import torch

from .. import kernels as K
from .. import sde_lib
from .predictors import _prep

_CORRECTORS = {}


def register_corrector(cls=None, *, name=None):
    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _CORRECTORS:
            raise ValueError(f"Already registered model with name: {local_name}")
        _CORRECTORS[local_name] = cls
        return cls

    return _register if cls is None else _register(cls)


def get_corrector(name):
    return _CORRECTORS[name]


class Corrector(abc.ABC):
    """The abstract class for a corrector algorithm (correctors.py:29-49)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps

    @abc.abstractmethod
    def update_fn(self, x, t):
        pass


def _alpha(sde, t):
    if isinstance(sde, (sde_lib.VPSDE, sde_lib.cVPSDE, sde_lib.subVPSDE)):
        timestep = (t * (sde.N - 1) / sde.T).long()
        return sde._on("alphas", t.device)[timestep].contiguous()
    return None


class _Langevin:
    def _update(self, x, t, *cond):
        x = _prep(x)
        alpha = _alpha(self.sde, t)
        norms = torch.empty(2 * x.shape[0], device=x.device, dtype=torch.float32)
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, *cond, t).contiguous()
            noise = torch.randn_like(x)
            K.langevin_norms(grad, noise, norms)
            x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
            K.langevin_update(x, grad, noise, norms, x_out, x_mean, self.snr, alpha, None, 1)
            x = x_out
        return x, x_mean


@register_corrector(name="langevin")
class LangevinCorrector(_Langevin, Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE, sde_lib.subVPSDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t):
        return self._update(x, t)


@register_corrector(name="conditional_langevin")
class conditionalLangevinCorrector(_Langevin, Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, (sde_lib.cVESDE, sde_lib.cVPSDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, y, t):
        return self._update(x, t, y)


@register_corrector(name="ald")
class AnnealedLangevinDynamics(Corrector):
    """step = 2 alpha (snr * std(t))^2 (correctors.py:111-142): a Langevin update whose step does not
    depend on the batch norms; evaluated with the reverse-diffusion kernel (x + g^2 score + ... z)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE, sde_lib.subVPSDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t):
        x = _prep(x)
        alpha = _alpha(self.sde, t)
        alpha = torch.ones_like(t) if alpha is None else alpha
        std = self.sde.marginal_prob(x, t)[1]
        step = (self.snr * std) ** 2 * 2 * alpha
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t).contiguous()
            noise = torch.randn_like(x)
            x_tmp, x_mean, x_out = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
            # x_mean = x + step*grad  (g^2 = step, no noise)
            K.reverse_diffusion_update(x, grad, None, x_tmp, x_mean, None, torch.sqrt(2.0 * step).contiguous(), True,
                                       None, 1)
            K.ve_perturb(x_mean, noise, x_out, torch.sqrt(step * 2).contiguous(), None, 1)
            x = x_out
        return x, x_mean


@register_corrector(name="none")
class NoneCorrector(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        pass

    def update_fn(self, x, t):
        return x, x


@register_corrector(name="conditional_none")
class NoneCorrector(Corrector):  # noqa: F811 - same shadowing as the reference (correctors.py:145-164)
    def __init__(self, sde, score_fn, snr, n_steps):
        pass

    def update_fn(self, x, y, t):
        return x, x
