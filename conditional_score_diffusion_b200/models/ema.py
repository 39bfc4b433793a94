"""Exponential moving average of the score network's parameters (reference: models/ema.py:11-187, used by
lightning_callbacks/callbacks.py:119-133 as `ema.update` after every optimizer step and `store / copy_to / restore`
around validation and sampling).

Same public surface (`update`, `copy_to`, `store`, `restore`, `state_dict`, `load_state_dict`, the num_updates-warmed
decay of models/ema.py:80-85). Differences that matter on this engine:

* the reference updates 620 tensors with a Python loop of 3 tiny kernels each; here the shadow parameters live in ONE
  flat fp32 buffer (views per parameter) and `update` is two multi-tensor calls;
* the reference swaps weights through `param.data.copy_` (models/ema.py:111,149), which does not bump autograd's version
  counter - the engine's packed bf16/tf32 operands would go stale. `copy_to` / `restore` here copy in place under
  no_grad (version bump) and, when handed the owning module, also call `engine.invalidate()`;
* `state_dict()` carries the shadow parameters, so checkpoints written through `models.utils.save_lightning_checkpoint`
  keep the EMA weights (the reference's EMACallback drops them).

`optim.FusedAdamEMA` fuses this update into the optimizer kernel; this class is the drop-in for callers that keep the
reference's separate optimizer + EMA objects.
"""
import copy
import weakref

import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True, module=None):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        tracked = [p for p in parameters if p.requires_grad]
        self._flat = None
        self.shadow_params = self._make_shadow(tracked)
        self.collected_params = []
        self._params_refs = [weakref.ref(p) for p in parameters]
        self._module_ref = weakref.ref(module) if module is not None else None

    def _make_shadow(self, tracked):
        """One flat buffer per (device, dtype) group when all parameters share it, else per-tensor clones."""
        if tracked and all(p.device == tracked[0].device and p.dtype == tracked[0].dtype for p in tracked):
            flat = torch.empty(sum(p.numel() for p in tracked), device=tracked[0].device, dtype=tracked[0].dtype)
            views, off = [], 0
            for p in tracked:
                v = flat[off:off + p.numel()].view_as(p)
                v.copy_(p.detach())
                views.append(v)
                off += p.numel()
            self._flat = flat
            return views
        return [p.clone().detach() for p in tracked]

    def _get_parameters(self, parameters):
        if parameters is None:
            parameters = [p() for p in self._params_refs]
            if any(p is None for p in parameters):
                raise ValueError("(one of) the parameters this ExponentialMovingAverage was initialised with no longer "
                                 "exists; pass `parameters` explicitly or keep the model alive")
            return parameters
        return list(parameters)

    def _invalidate_engine(self):
        m = self._module_ref() if self._module_ref is not None else None
        eng = getattr(m, "_engine", None) if m is not None else None
        if eng is not None:
            eng.invalidate()

    def update(self, parameters=None):
        parameters = [p for p in self._get_parameters(parameters) if p.requires_grad]
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            # s <- s - (1 - decay) (s - p)  ==  lerp(s, p, 1 - decay): one multi-tensor launch group
            srcs = [p.detach().to(s.device) for s, p in zip(self.shadow_params, parameters)]
            torch._foreach_lerp_(self.shadow_params, srcs, 1.0 - decay)

    def copy_to(self, parameters=None):
        parameters = [p for p in self._get_parameters(parameters) if p.requires_grad]
        with torch.no_grad():
            for s, p in zip(self.shadow_params, parameters):
                p.copy_(s)          # in place under no_grad: bumps p._version, unlike the reference's p.data.copy_
        self._invalidate_engine()

    def store(self, parameters=None):
        parameters = self._get_parameters(parameters)
        self.collected_params = [p.detach().clone() for p in parameters if p.requires_grad]

    def restore(self, parameters=None):
        parameters = [p for p in self._get_parameters(parameters) if p.requires_grad]
        with torch.no_grad():
            for c, p in zip(self.collected_params, parameters):
                p.copy_(c)
        self._invalidate_engine()

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state_dict):
        state_dict = copy.deepcopy(state_dict)
        self.decay = state_dict["decay"]
        if self.decay < 0.0 or self.decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.num_updates = state_dict["num_updates"]
        if not (self.num_updates is None or isinstance(self.num_updates, int)):
            raise ValueError("invalid num_updates")
        shadow = state_dict["shadow_params"]
        if not (isinstance(shadow, list) and all(torch.is_tensor(t) for t in shadow)):
            raise ValueError("shadow_params must be a list of tensors")
        if len(shadow) != len(self.shadow_params):
            raise ValueError("shadow_params: wrong number of tensors")
        with torch.no_grad():
            for dst, src in zip(self.shadow_params, shadow):       # keep the flat buffer: copy values in
                dst.copy_(src.to(dst.device))
        collected = state_dict.get("collected_params", [])
        if not (isinstance(collected, list) and all(torch.is_tensor(t) for t in collected)):
            raise ValueError("collected_params must be a list of tensors")
        self.collected_params = collected
