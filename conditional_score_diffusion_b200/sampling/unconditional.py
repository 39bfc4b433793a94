"""Unconditional samplers (reference: sampling/unconditional.py:13-228).

`get_pc_sampler` keeps the reference signature and return values. With an engine-backed NCSN++ and a
(predictor, corrector) pair covered by the fused loop it runs `fused.FusedPCSampler` (one CUDA graph
per step); otherwise it runs the same corrector-then-predictor loop through the predictor /
corrector classes, whose updates are still single fused CUDA kernels.
"""
import functools

import torch

from ..models import utils as mutils
from . import fused
from .correctors import NoneCorrector, get_corrector, _CORRECTORS
from .predictors import NonePredictor, ReverseDiffusionPredictor, get_predictor, _PREDICTORS

_FUSED_PREDICTORS = {"reverse_diffusion": "reverse_diffusion", "euler_maruyama": "euler_maruyama", "none": "none",
                     "conditional_reverse_diffusion": "reverse_diffusion",
                     "conditional_euler_maruyama": "euler_maruyama", "conditional_none": "none"}
_FUSED_CORRECTORS = {"langevin": "langevin", "none": "none", "conditional_langevin": "langevin",
                     "conditional_none": "none"}


# get_pc_inpainter on the fused CUDA-graph loop (False keeps the per-step Python loop; A/B switch for the parity tests)
FUSED_INPAINT = True


def _registry_name(cls, registry):
    if cls is None:
        return "none"
    for k, v in registry.items():
        if v is cls:
            return k
    return None


def fused_kinds(predictor, corrector):
    """Map predictor / corrector classes to the fused loop's kinds, or None if not covered."""
    p = _FUSED_PREDICTORS.get(_registry_name(predictor, _PREDICTORS))
    c = _FUSED_CORRECTORS.get(_registry_name(corrector, _CORRECTORS))
    return (p, c) if p is not None and c is not None else None


def get_sampling_fn(config, sde, shape, eps, predictor="default", corrector="default", p_steps="default",
                    c_steps="default", snr="default", denoise="default"):
    """sampling/unconditional.py:13-75."""
    predictor = get_predictor((config.sampling.predictor if predictor == "default" else predictor).lower())
    corrector = get_corrector((config.sampling.corrector if corrector == "default" else corrector).lower())
    p_steps = config.model.num_scales if p_steps == "default" else p_steps
    c_steps = config.sampling.n_steps_each if c_steps == "default" else c_steps
    snr = config.sampling.snr if snr == "default" else snr
    denoise = config.sampling.noise_removal if denoise == "default" else denoise
    name = config.sampling.method.lower()
    if name == "pc":
        return get_pc_sampler(sde=sde, shape=shape, predictor=predictor, corrector=corrector, snr=snr, p_steps=p_steps,
                              c_steps=c_steps, probability_flow=config.sampling.probability_flow,
                              continuous=config.training.continuous, denoise=denoise, eps=eps)
    if name == "ode":
        return get_ode_sampler(sde=sde, shape=shape, denoise=denoise, eps=eps)
    raise ValueError(f"Sampler name {name} unknown.")


def get_ode_sampler(sde, shape, denoise=False, rtol=1e-5, atol=1e-5, method="RK45", eps=1e-3, device_integrator=True):
    """Probability-flow ODE sampler (sampling/unconditional.py:93-158).

    Returns ode_sampler(model, z=None) -> (samples, nfe). The right-hand side is one engine forward (CUDA graph replay)
    per call. With method='RK45' (the reference's default) on a CUDA model the integration runs in `ode.solve_rk45`, a
    restatement of scipy's RK45 whose state and stage derivatives stay in HBM (only the scalar error norm reaches the
    host per step); any other method, a CPU model, or device_integrator=False uses scipy.integrate.solve_ivp with the
    state crossing the host once per evaluation, as in the reference."""
    import numpy as np
    from scipy import integrate

    def denoise_update_fn(model, x):
        score_fn = mutils.get_score_fn(sde, model, conditional=False, train=False, continuous=True)
        predictor_obj = ReverseDiffusionPredictor(sde, score_fn, probability_flow=False)
        vec_eps = torch.ones(x.shape[0], device=x.device) * eps
        _, x = predictor_obj.update_fn(x, vec_eps)
        return x

    def drift_fn(model, x, t):
        score_fn = mutils.get_score_fn(sde, model, conditional=False, train=False, continuous=True)
        rsde = sde.reverse(score_fn, probability_flow=True)
        return rsde.sde(x, t)[0]

    def ode_sampler(model, z=None):
        with torch.no_grad():
            x = sde.prior_sampling(shape).to(model.device) if z is None else z

            def ode_func(t, x):
                x = mutils.from_flattened_numpy(x, shape).to(model.device).type(torch.float32)
                vec_t = torch.ones(shape[0], device=x.device) * t
                return mutils.to_flattened_numpy(drift_fn(model, x, vec_t))

            if device_integrator and method == "RK45" and x.is_cuda:
                from .. import ode

                def rhs(t, yflat):
                    vec_t = torch.ones(shape[0], device=yflat.device) * t
                    return drift_fn(model, yflat.view(shape), vec_t).reshape(-1)

                yT, nfe = ode.solve_rk45(rhs, float(sde.T), float(eps), x.reshape(-1).float().contiguous(), rtol, atol)
                x = yT.view(shape)
            else:
                solution = integrate.solve_ivp(ode_func, (sde.T, eps), mutils.to_flattened_numpy(x), rtol=rtol, atol=atol,
                                               method=method)
                nfe = solution.nfev
                x = torch.tensor(solution.y[:, -1]).reshape(shape).to(model.device).type(torch.float32)
            if denoise:
                x = denoise_update_fn(model, x)
            return x, nfe

    return ode_sampler


def get_pc_inpainter(sde, predictor, corrector, snr, n_steps=1, probability_flow=False, continuous=False, denoise=True,
                     eps=1e-5):
    """Image inpainting with PC samplers (sampling/unconditional.py:230-345): after every corrector / predictor update
    the known pixels (mask = 1) are replaced by the data perturbed to the current noise level - one fused kernel
    (csd_inpaint_merge_f32) instead of the reference's 8 elementwise ops. Returns pc_inpainter(model, data, mask,
    show_evolution=False) -> (samples, info)."""
    from .. import kernels as K
    predictor_update_fn = functools.partial(shared_predictor_update_fn, sde=sde, predictor=predictor,
                                            probability_flow=probability_flow, continuous=continuous)
    corrector_update_fn = functools.partial(shared_corrector_update_fn, sde=sde, corrector=corrector,
                                            continuous=continuous, snr=snr, n_steps=n_steps)

    def inpaint_update(update_fn, model, data, mask, x, vec_t):
        x, _ = update_fn(x, vec_t, model=model)
        mean_coef, std = sde.marginal_prob(torch.ones_like(vec_t), vec_t)
        z = torch.randn_like(x)
        return K.inpaint_merge(x.contiguous(), data, z, mask, torch.empty_like(x), torch.empty_like(x),
                               mean_coef.float().contiguous(), std.float().contiguous())

    kinds = fused_kinds(predictor, corrector)
    cache = {}

    def pc_inpainter(model, data, mask, show_evolution=False, x_init=None, noise_source=None):
        if kinds is not None and hasattr(model, "_engine") and FUSED_INPAINT:
            key = (id(model), tuple(data.shape))
            fs = cache.get(key)
            if fs is None:
                fs = fused.FusedPCSampler(model, sde, tuple(data.shape), kinds[0], kinds[1], snr, sde.N, n_steps,
                                          probability_flow, continuous, denoise, eps, conditional=False, inpaint=True)
                cache[key] = fs
            samples, evo = fs.sample(x_init=x_init, noise_source=noise_source, show_evolution=show_evolution, data=data,
                                     mask=mask)
            return samples, ({"evolution": torch.stack(evo["x"])} if show_evolution else {})
        with torch.no_grad():
            data = data.contiguous().float()
            mask = mask.expand_as(data).contiguous().float()
            prior = (sde.prior_sampling(data.shape) if x_init is None else x_init).type_as(data)
            # x = data * mask + prior * (1 - mask): the merge kernel with mean_coef = 1, std = 0
            zero = torch.zeros(data.shape[0], device=data.device)
            x, _ = K.inpaint_merge(prior.contiguous(), data, torch.zeros_like(data), mask, torch.empty_like(data),
                                   torch.empty_like(data), None, zero)
            evolution = [x.cpu()] if show_evolution else None
            timesteps = torch.linspace(sde.T, eps, sde.N)
            x_mean = x
            for i in range(sde.N):
                vec_t = torch.ones(data.shape[0], device=data.device) * timesteps[i]
                x, x_mean = inpaint_update(corrector_update_fn, model, data, mask, x, vec_t)
                x, x_mean = inpaint_update(predictor_update_fn, model, data, mask, x, vec_t)
                if show_evolution:
                    evolution.append(x.cpu())
            info = {"evolution": torch.stack(evolution)} if show_evolution else {}
            return (x_mean if denoise else x), info

    return pc_inpainter


def shared_predictor_update_fn(x, t, sde, model, predictor, probability_flow, continuous):
    """sampling/unconditional.py:347-356."""
    score_fn = mutils.get_score_fn(sde, model, conditional=False, train=False, continuous=continuous)
    obj = NonePredictor(sde, score_fn, probability_flow) if predictor is None else predictor(sde, score_fn, probability_flow)
    return obj.update_fn(x, t)


def shared_corrector_update_fn(x, t, sde, model, corrector, continuous, snr, n_steps):
    """sampling/unconditional.py:358-367."""
    score_fn = mutils.get_score_fn(sde, model, conditional=False, train=False, continuous=continuous)
    obj = NoneCorrector(sde, score_fn, snr, n_steps) if corrector is None else corrector(sde, score_fn, snr, n_steps)
    return obj.update_fn(x, t)


def get_pc_sampler(sde, shape, predictor, corrector, snr, p_steps, c_steps, probability_flow=False, continuous=False,
                   denoise=True, eps=1e-3):
    """Create a Predictor-Corrector sampler (sampling/unconditional.py:161-228).

    Returns pc_sampler(model, show_evolution=False) -> (samples, {'times', 'steps'[, 'evolution']}).
    """
    kinds = fused_kinds(predictor, corrector)
    predictor_update_fn = functools.partial(shared_predictor_update_fn, sde=sde, predictor=predictor,
                                            probability_flow=probability_flow, continuous=continuous)
    corrector_update_fn = functools.partial(shared_corrector_update_fn, sde=sde, corrector=corrector,
                                            continuous=continuous, snr=snr, n_steps=c_steps)
    cache = {}

    def pc_sampler(model, show_evolution=False, x_init=None, noise_source=None):
        steps = p_steps * (c_steps + 1)
        if kinds is not None and hasattr(model, "_engine"):
            fs = cache.get(id(model))
            if fs is None:
                fs = fused.FusedPCSampler(model, sde, shape, kinds[0], kinds[1], snr, p_steps, c_steps,
                                          probability_flow, continuous, denoise, eps, conditional=False)
                cache[id(model)] = fs
            samples, evo = fs.sample(x_init=x_init, noise_source=noise_source, show_evolution=show_evolution)
            info = {"times": fs.timesteps.to(model.device), "steps": steps}
            if show_evolution:
                info["evolution"] = torch.stack(evo["x"])
            return samples, info
        evolution = []
        with torch.no_grad():
            x = (sde.prior_sampling(shape) if x_init is None else x_init).to(model.device).type(torch.float32)
            timesteps = torch.linspace(sde.T, eps, p_steps, device=model.device)
            x_mean = x
            for i in range(p_steps):
                vec_t = torch.ones(shape[0], device=model.device) * timesteps[i]
                x, x_mean = corrector_update_fn(x, vec_t, model=model)
                x, x_mean = predictor_update_fn(x, vec_t, model=model)
                if show_evolution:
                    evolution.append(x.cpu())
            info = {"times": timesteps, "steps": steps}
            if show_evolution:
                info["evolution"] = torch.stack(evolution)
            return (x_mean if denoise else x), info

    return pc_sampler
