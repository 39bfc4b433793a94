"""Model registry and score-function wrappers (reference: models/utils.py:24-47, 50-74, 114-278).

Same registry keys and the same `get_score_fn` contract. When the model is one of this package's
engine-backed networks the division by sigma(t) (models/utils.py:50-74) is fused into the kernel that
writes the network output (csd_nhwc_bf16_to_nchw's row scale) instead of running as separate ops.
"""
import numpy as np
import torch

from .. import sde_lib

_MODELS = {}


def register_model(cls=None, *, name=None):
    """A decorator for registering model classes (models/utils.py:27-43)."""

    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _MODELS:
            raise ValueError(f"Already registered model with name: {local_name}")
        _MODELS[local_name] = cls
        return cls

    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]


def create_model(config):
    """models/utils.py:114-120."""
    return get_model(config.model.name)(config)


def get_sigmas(config):
    """models/utils.py:76-86."""
    return np.exp(np.linspace(np.log(config.model.sigma_max), np.log(config.model.sigma_min),
                              config.model.num_scales))


def _bc(v, like):
    return v[(...,) + (None,) * (like.ndim - 1)]


def _scaled_forward(model, x, labels, inv_std):
    """model(x, labels) / std with the division fused when the model supports it.

    inv_std: tensor [B] or dict of tensors [B] keyed like the model's output dict.
    """
    if hasattr(model, "forward_scaled"):
        return model.forward_scaled(x, labels, inv_std)
    out = model(x, labels)
    if isinstance(out, dict):
        return {k: v * _bc(inv_std[k], v) for k, v in out.items()}
    return out * _bc(inv_std, out)


def get_model_fn(model, train=False):
    """models/utils.py:123-152."""

    def model_fn(x, labels):
        if not train:
            model.eval()
        else:
            model.train()
        return model(x, labels)

    return model_fn


def divide_by_sigmas(h, labels, sde, continuous=False):
    """models/utils.py:50-74 (kept for callers that use it directly)."""
    inv = _inv_sigmas(h, labels, sde, continuous)
    if isinstance(h, dict):
        return {k: v * _bc(inv[k], v) for k, v in h.items()}
    return h * _bc(inv, h)


def _inv_sigmas(like, labels, sde, continuous):
    def one(s, ref):
        if continuous:
            return 1.0 / s.marginal_prob(torch.zeros(1, device=labels.device), labels)[1]
        return 1.0 / s.discrete_sigmas.to(labels.device)[labels]

    if isinstance(sde, dict):
        keys = like.keys() if isinstance(like, dict) else sde.keys()
        return {k: one(sde[k], None) for k in keys}
    return one(sde, None)


def get_score_fn(sde, model, conditional=False, train=False, continuous=False):
    """models/utils.py:156-267: network output -> score (divide by the perturbation std)."""

    def run(x, labels, inv_std):
        if not train:
            model.eval()
        else:
            model.train()
        return _scaled_forward(model, x, labels, inv_std)

    if conditional:
        if isinstance(sde, dict):
            if isinstance(sde["y"], (sde_lib.VPSDE, sde_lib.subVPSDE)):
                raise NotImplementedError("This combination of sdes is not supported for conditional SDEs yet.")
            if isinstance(sde["y"], sde_lib.VESDE) and isinstance(sde["x"], sde_lib.cVESDE) and len(sde) == 2:
                def score_fn(x, t):
                    if continuous:
                        labels = t * (sde["x"].N - 1)
                        inv = {k: 1.0 / sde[k].marginal_prob(t, t)[1] for k in ("x", "y")}
                    else:
                        labels = torch.round((t * (sde["x"].N - 1)).float()).long()
                        inv = {k: 1.0 / sde[k].discrete_sigmas.to(t.device)[labels] for k in ("x", "y")}
                    return run(x, labels, inv)
                return score_fn
            raise NotImplementedError("This combination of SDEs is not supported for conditional SDEs yet.")
        if isinstance(sde, sde_lib.cVPSDE):
            def score_fn(x, t):
                labels = t * (sde.N - 1)
                if continuous:
                    std = sde.marginal_prob(t, t)[1]
                else:
                    std = sde.sqrt_1m_alphas_cumprod.to(t.device)[labels.long()]
                return run(x, labels, 1.0 / std)
            return score_fn
        if isinstance(sde, (sde_lib.VESDE, sde_lib.cVESDE)):
            def score_fn(x, t):
                if continuous:
                    labels = t * (sde.N - 1)
                    inv = 1.0 / sde.marginal_prob(t, t)[1]
                else:
                    labels = torch.round((t * (sde.N - 1)).float()).long()
                    inv = 1.0 / sde.discrete_sigmas.to(t.device)[labels]
                return run(x, labels, inv)
            return score_fn
        raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
        def score_fn(x, t):
            labels = t * (sde.N - 1)
            if continuous or isinstance(sde, sde_lib.subVPSDE):
                std = sde.marginal_prob(t, t)[1]
            else:
                std = sde.sqrt_1m_alphas_cumprod.to(t.device)[labels.long()]
            return run(x, labels, 1.0 / std)
        return score_fn
    if isinstance(sde, (sde_lib.VESDE, sde_lib.cVESDE)):
        def score_fn(x, t):
            if continuous:
                std = sde.marginal_prob(t, t)[1]
                emb = torch.log(std) if model.embedding_type == "fourier" else std
            else:
                labels = torch.round(t * (sde.N - 1)).long()
                std = sde.discrete_sigmas.to(t.device)[labels]
                emb = std
            return run(x, emb, 1.0 / std)
        return score_fn
    raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")


def get_conditional_score_fn(score_fn, target_domain):
    """models/utils.py:270-278."""

    def conditional_score_fn(x, y, t):
        score = score_fn({"x": x, "y": y}, t)
        return score[target_domain] if isinstance(score, dict) else score

    return conditional_score_fn


def to_flattened_numpy(x):
    return x.detach().cpu().numpy().reshape((-1,))


def from_flattened_numpy(x, shape):
    return torch.from_numpy(x.reshape(shape))


def load_lightning_checkpoint(model, path_or_state, ema=None, strict=True):
    """Load a checkpoint written by the reference's Lightning modules into an engine-backed network.

    The reference saves `pl_module.state_dict()`, whose score-network entries are keyed `score_model.all_modules.<i>...`
    (lightning_modules/BaseSdeGenerativeModel.py:22-23), next to `hyper_parameters` (the whole config,
    BaseSdeGenerativeModel.py:17). `path_or_state` is a checkpoint path, the loaded checkpoint dict, or a bare state
    dict; the `score_model.` prefix is stripped and the rest loaded with the reference's own key names (the positional
    `all_modules` order is identical here). ema: an `models.ema.ExponentialMovingAverage` to restore from the
    checkpoint's `ema_state` entry (written by save_lightning_checkpoint; reference checkpoints have none -
    lightning_callbacks/callbacks.py:119-133 never saves the shadows - in which case the EMA is re-seeded from the loaded
    weights). Returns the (missing, unexpected) keys of load_state_dict."""
    ckpt = torch.load(path_or_state, map_location="cpu", weights_only=False) if isinstance(path_or_state, str) else path_or_state
    state = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    prefix = "score_model."
    if any(k.startswith(prefix) for k in state):
        state = {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
    res = model.load_state_dict(state, strict=strict)
    eng = getattr(model, "_engine", None)
    if eng is not None:
        eng.invalidate()
    if ema is not None and ema is not False:
        ema_state = ckpt.get("ema_state") if isinstance(ckpt, dict) else None
        if ema_state is not None:
            ema.load_state_dict(ema_state)
        else:
            with torch.no_grad():
                for s, p in zip(ema.shadow_params, [p for p in model.parameters() if p.requires_grad]):
                    s.copy_(p.detach().to(s.device))
    return res


def checkpoint_config(path_or_ckpt):
    """The config a reference checkpoint was trained with: `hyper_parameters['config']` (save_hyperparameters(),
    lightning_modules/BaseSdeGenerativeModel.py:17), or None for a bare state dict."""
    ckpt = torch.load(path_or_ckpt, map_location="cpu", weights_only=False) if isinstance(path_or_ckpt, str) else path_or_ckpt
    if isinstance(ckpt, dict):
        hp = ckpt.get("hyper_parameters")
        if isinstance(hp, dict):
            return hp.get("config")
    return None


def save_lightning_checkpoint(model, path=None, config=None, ema=None, extra=None):
    """Write the network back in the reference's Lightning checkpoint layout: `state_dict` with the `score_model.`
    prefix (so `load_from_checkpoint`, lightning_modules/utils.py:24-28, and `resume_from_checkpoint`, run_lib.py:63,
    read it), `hyper_parameters.config`, and - unlike the reference, whose EMACallback has no save hook - the EMA
    shadow parameters under `ema_state` (ignored by Lightning, restored by load_lightning_checkpoint). `extra`: more
    top-level entries (epoch, global_step, optimizer_states ...). Returns the checkpoint dict; writes it when `path`
    is given."""
    state = {"score_model." + k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ckpt = {"state_dict": state, "hyper_parameters": {"config": config if config is not None else getattr(model, "config", None)}}
    if ema is not None:
        es = ema.state_dict()
        ckpt["ema_state"] = {"decay": es["decay"], "num_updates": es["num_updates"],
                             "shadow_params": [t.detach().cpu().clone() for t in es["shadow_params"]],
                             "collected_params": [t.detach().cpu().clone() for t in es["collected_params"]]}
    if extra:
        ckpt.update(extra)
    if path is not None:
        torch.save(ckpt, path)
    return ckpt
