"""Conditional samplers (reference: sampling/conditional.py:8-254).

`get_pc_conditional_sampler` keeps the reference signature; the {'x': cVESDE, 'y': VESDE} pair with
use_path=False (the configuration every shipped config uses) runs on the fused CUDA-graph loop.
"""
import functools

import torch

from .. import kernels as K
from ..models import utils as mutils
from . import fused
from .correctors import NoneCorrector, get_corrector
from .predictors import NonePredictor, get_predictor
from .unconditional import fused_kinds

# use_path=True on the fused CUDA-graph loop (False keeps the per-step Python loop; A/B switch for the parity tests)
FUSED_PATH = True


def get_conditional_sampling_fn(config, sde, shape, eps, predictor="default", corrector="default", p_steps="default",
                                c_steps="default", snr="default", denoise="default", use_path="default"):
    """sampling/conditional.py:8-45."""
    predictor = get_predictor((config.sampling.predictor if predictor == "default" else predictor).lower())
    corrector = get_corrector((config.sampling.corrector if corrector == "default" else corrector).lower())
    p_steps = config.model.num_scales if p_steps == "default" else p_steps
    c_steps = config.sampling.n_steps_each if c_steps == "default" else c_steps
    snr = config.sampling.snr if snr == "default" else snr
    denoise = config.sampling.noise_removal if denoise == "default" else denoise
    use_path = False if use_path == "default" else use_path
    return get_pc_conditional_sampler(sde=sde, shape=shape, predictor=predictor, corrector=corrector, snr=snr,
                                      p_steps=p_steps, c_steps=c_steps,
                                      probability_flow=config.sampling.probability_flow,
                                      continuous=config.training.continuous, denoise=denoise, use_path=use_path,
                                      eps=eps)


def conditional_shared_predictor_update_fn(x, y, t, sde, model, predictor, probability_flow, continuous):
    """sampling/conditional.py:230-242."""
    score_fn = mutils.get_score_fn(sde, model, conditional=True, train=False, continuous=continuous)
    score_fn = mutils.get_conditional_score_fn(score_fn, target_domain="x")
    c_sde = sde["x"] if isinstance(sde, dict) else sde
    obj = NonePredictor(c_sde, score_fn, probability_flow) if predictor is None else predictor(c_sde, score_fn,
                                                                                              probability_flow)
    return obj.update_fn(x, y, t)


def conditional_shared_corrector_update_fn(x, y, t, sde, model, corrector, continuous, snr, n_steps):
    """sampling/conditional.py:244-255."""
    score_fn = mutils.get_score_fn(sde, model, conditional=True, train=False, continuous=continuous)
    score_fn = mutils.get_conditional_score_fn(score_fn, target_domain="x")
    c_sde = sde["x"] if isinstance(sde, dict) else sde
    obj = NoneCorrector(c_sde, score_fn, snr, n_steps) if corrector is None else corrector(c_sde, score_fn, snr, n_steps)
    return obj.update_fn(x, y, t)


def get_pc_conditional_sampler(sde, shape, predictor, corrector, snr, p_steps, c_steps=1, probability_flow=False,
                               continuous=False, denoise=True, use_path=False, eps=1e-5):
    """Create a conditional PC sampler (sampling/conditional.py:47-228).

    Returns f(model, y, show_evolution=False) -> (samples, info).
    """
    kinds = fused_kinds(predictor, corrector)
    pair = isinstance(sde, dict) and len(sde.keys()) == 2
    predictor_update_fn = functools.partial(conditional_shared_predictor_update_fn, sde=sde, predictor=predictor,
                                            probability_flow=probability_flow, continuous=continuous)
    corrector_update_fn = functools.partial(conditional_shared_corrector_update_fn, sde=sde, corrector=corrector,
                                            continuous=continuous, snr=snr, n_steps=c_steps)
    cache = {}

    def perturbed(y, vec_t):
        if not pair:
            return y
        std = sde["y"].marginal_prob(vec_t, vec_t)[1].contiguous()
        y = y.contiguous().float()
        return K.ve_perturb(y, torch.randn_like(y), torch.empty_like(y), std, None, 1)

    def backward_kernel_sample(y0, y_tpt, vec_t, vec_tau):
        """y_t ~ p(y_t | y_0, y_{t+tau}) (VESDE.compute_backward_kernel, sde_lib.py:323-339) in two fused passes:
        mean = a y_0 + b y_{t+tau}, then mean + std z."""
        s_t = sde["y"].marginal_prob(vec_t, vec_t)[1] ** 2
        s_tau = sde["y"].marginal_prob(vec_t, vec_t + vec_tau)[1] ** 2
        a = ((s_tau - s_t) / s_tau).float().contiguous()
        b = (s_t / s_tau).float().contiguous()
        std = torch.sqrt(s_t * (s_tau - s_t) / s_tau).float().contiguous()
        mean = K.sde_perturb(y0, y_tpt, torch.empty_like(y0), a, b)
        return K.sde_perturb(mean, torch.randn_like(y0), torch.empty_like(y0), None, std)

    def pc_conditional_sampler_path(model, y, show_evolution=False, x_init=None, noise_source=None):
        """use_path=True (sampling/conditional.py:87-94,124-176): y walks down ONE consistent path y_T -> y_0 through
        the backward kernel; predictor first, then the corrector on the same y_t."""
        if not pair:
            raise NotImplementedError("use_path=True needs the {'x', 'y'} SDE pair (sampling/conditional.py:86-87)")
        c_sde = sde["x"]
        if kinds is not None and hasattr(model, "_engine") and FUSED_PATH:
            fs = cache.get(("path", id(model)))
            if fs is None:
                fs = fused.FusedPCSampler(model, sde, shape, kinds[0], kinds[1], snr, p_steps, c_steps,
                                          probability_flow, continuous, denoise, eps, conditional=True, use_path=True)
                cache[("path", id(model))] = fs
            samples, evo = fs.sample(y=y, x_init=x_init, noise_source=noise_source, show_evolution=show_evolution)
            if show_evolution:
                return samples, {"evolution": {"x": torch.stack(evo["x"]), "y": torch.stack(evo["y"])}}
            return samples, {}
        with torch.no_grad():
            dev = model.device
            y = y.to(dev).contiguous().float()
            x = (c_sde.prior_sampling(shape) if x_init is None else x_init).to(dev)
            timesteps = torch.linspace(c_sde.T, eps, p_steps, device=dev)
            tau = timesteps[0] - timesteps[1]
            ones = torch.ones(x.shape[0], device=dev)
            std_T = sde["y"].marginal_prob(ones, ones * (timesteps[0] + tau))[1].float().contiguous()
            y_tpt = K.sde_perturb(y, torch.randn_like(y), torch.empty_like(y), None, std_T)
            evolution = {"x": [], "y": []}
            x_mean = x
            for i in range(p_steps):
                vec_t = ones * timesteps[i]
                y_tpt = backward_kernel_sample(y, y_tpt, vec_t, ones * tau)
                x, x_mean = predictor_update_fn(x=x, y=y_tpt, t=vec_t, model=model)
                x, x_mean = corrector_update_fn(x=x, y=y_tpt, t=vec_t, model=model)
                if show_evolution:
                    evolution["x"].append(x.cpu())
                    evolution["y"].append(y_tpt.cpu())
            if show_evolution:
                return (x_mean if denoise else x), {"evolution": {"x": torch.stack(evolution["x"]),
                                                                  "y": torch.stack(evolution["y"])}}
            return (x_mean if denoise else x), {}

    if use_path:
        return pc_conditional_sampler_path

    def pc_conditional_sampler(model, y, show_evolution=False, x_init=None, noise_source=None):
        c_sde = sde["x"] if isinstance(sde, dict) else sde
        if kinds is not None and pair and hasattr(model, "_engine"):
            fs = cache.get(id(model))
            if fs is None:
                fs = fused.FusedPCSampler(model, sde, shape, kinds[0], kinds[1], snr, p_steps, c_steps,
                                          probability_flow, continuous, denoise, eps, conditional=True)
                cache[id(model)] = fs
            samples, evo = fs.sample(y=y, x_init=x_init, noise_source=noise_source, show_evolution=show_evolution)
            if show_evolution:
                return samples, {"evolution": {"x": torch.stack(evo["x"]), "y": torch.stack(evo["y"])}}
            return samples, {}
        with torch.no_grad():
            x = (c_sde.prior_sampling(shape) if x_init is None else x_init).to(model.device)
            evolution = {"x": [], "y": []}
            timesteps = torch.linspace(c_sde.T, eps, p_steps, device=model.device)
            x_mean = x
            for i in range(p_steps):
                vec_t = torch.ones(x.shape[0], device=model.device) * timesteps[i]
                y_t = perturbed(y, vec_t)
                x, x_mean = corrector_update_fn(x=x, y=y_t, t=vec_t, model=model)
                y_t = perturbed(y, vec_t)
                x, x_mean = predictor_update_fn(x=x, y=y_t, t=vec_t, model=model)
                if show_evolution:
                    evolution["x"].append(x.cpu())
                    evolution["y"].append(y_t.cpu())
            if show_evolution:
                return (x_mean if denoise else x), {"evolution": {"x": torch.stack(evolution["x"]),
                                                                  "y": torch.stack(evolution["y"])}}
            return (x_mean if denoise else x), {}

    return pc_conditional_sampler
