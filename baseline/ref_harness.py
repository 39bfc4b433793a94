"""Drive the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py) for bench.py's reference arm
(CPU, `kind: "reference"`) and for the stock PyTorch + cuDNN leg on the B200 (the denominator of the north star's
">= 10x stock" target). Nothing here is imported by the product package.

The reference imports three things this image lacks or that must not run at import time; they are shimmed, the
reference's own files are untouched (sha256 in baseline/_ref/MANIFEST.json):
  * pytorch_lightning.LightningModule  -> nn.Module + .device (models/ncsnpp.py:23,40, models/ddpm.py:24,81)
  * ml_collections.ConfigDict          -> attribute dict (config objects)
  * torch.utils.cpp_extension.load     -> returns the pre-built extension from baseline/_ref/_ext/<name>/<name>.so
                                          (same sources and flags as the reference's own load(...) call,
                                          op/upfirdn2d.py:10-16); falls back to the real JIT build when CUDA is in
                                          use and no pre-built library is there; skipped on a CPU-only run, where the
                                          reference takes its own `upfirdn2d_native` branch (op/upfirdn2d.py:146-149).
"""
import importlib.util
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_STATE = {"installed": False, "mods": None}


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def available():
    return os.path.exists(os.path.join(REF_DIR, "sampling", "conditional.py"))


def _install_shims(need_cuda_ops):
    import torch.nn as nn
    import torch.utils.cpp_extension as ext
    if _STATE["installed"]:
        return
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            @property
            def device(self):
                return next(self.parameters()).device

            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl
    if "ml_collections" not in sys.modules:
        ml = types.ModuleType("ml_collections")
        ml.ConfigDict = ConfigDict
        sys.modules["ml_collections"] = ml
    real_load = ext.load

    def load(name, sources, **kw):
        so = os.path.join(REF_DIR, "_ext", name, name + ".so")
        if os.path.exists(so):
            spec = importlib.util.spec_from_file_location(name, so)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
        if not need_cuda_ops:
            return None
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
        bdir = os.path.join(REF_DIR, "_ext", name)
        os.makedirs(bdir, exist_ok=True)
        return real_load(name, sources=sources, build_directory=bdir, **kw)

    ext.load = load
    sys.path.insert(0, REF_DIR)
    _STATE["installed"] = True


def modules(need_cuda_ops=False):
    """Import the reference's modules (once). Returns a namespace with sde_lib, mutils, sampling pieces."""
    if _STATE["mods"] is not None:
        return _STATE["mods"]
    if not available():
        raise RuntimeError("baseline/_ref is not installed (run baseline/install_ref.py in the build container)")
    _install_shims(need_cuda_ops)
    import sde_lib  # noqa: E402  (the reference's, from baseline/_ref)
    from models import ncsnpp, ddpm, utils as mutils  # noqa: F401
    from sampling import predictors, correctors  # noqa: F401
    from sampling.conditional import get_pc_conditional_sampler
    from sampling.unconditional import get_pc_sampler
    import losses
    m = types.SimpleNamespace(sde_lib=sde_lib, mutils=mutils, predictors=predictors, correctors=correctors,
                              get_pc_conditional_sampler=get_pc_conditional_sampler, get_pc_sampler=get_pc_sampler,
                              losses=losses)
    assert os.path.abspath(sde_lib.__file__).startswith(REF_DIR), "a different sde_lib shadowed the reference's"
    _STATE["mods"] = m
    return m


def to_configdict(ns):
    """types.SimpleNamespace tree (bench.py's workload config) -> ConfigDict tree."""
    out = ConfigDict()
    for k, v in vars(ns).items():
        out[k] = to_configdict(v) if isinstance(v, types.SimpleNamespace) else v
    return out


def create_model(cfg_ns, state_dict=None, device="cpu"):
    import torch
    m = modules(need_cuda_ops=(str(device) != "cpu"))
    cfg = to_configdict(cfg_ns)
    model = m.mutils.get_model(cfg.model.name)(cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model.to(torch.device(device)).eval()


def conditional_pc_time(cfg_ns, state_dict, y, steps, warmup_steps, device, snr, eps, n_scales=1000):
    """Time the reference's own conditional PC sampler (sampling/conditional.py:47-228) for `steps` PC steps:
    one untimed call with p_steps=warmup_steps, then one timed call with p_steps=steps. Every PC step costs the same
    (2 network evaluations + updates), so s/step x 1000 is the PC-1000 time per batch. Returns a dict."""
    import torch
    m = modules(need_cuda_ops=(str(device) != "cpu"))
    dev = torch.device(device)
    model = create_model(cfg_ns, state_dict, dev)
    cm = cfg_ns.model
    sde = {"x": m.sde_lib.cVESDE(cm.sigma_min_x, cm.sigma_max_x, n_scales),
           "y": m.sde_lib.VESDE(cm.sigma_min_y, cm.sigma_max_y, n_scales)}
    if dev.type == "cuda":
        # sde_lib.py:359,415 index the CPU table `discrete_sigmas` with a CUDA index tensor - accepted by the torch 1.x the
        # reference was written for, rejected by torch 2.11 ("indices should be either on cpu or on the same device").
        # The table is moved to the device on the INSTANCES (the `.to(t.device)` the reference then does is a no-op);
        # the reference's files stay byte-identical.
        for s_ in sde.values():
            s_.discrete_sigmas = s_.discrete_sigmas.to(dev)
    shape = tuple(y.shape)
    pred = m.predictors.get_predictor("conditional_reverse_diffusion")
    corr = m.correctors.get_corrector("conditional_langevin")

    def sampler(p_steps):
        return m.get_pc_conditional_sampler(sde, shape, pred, corr, snr, p_steps, 1, probability_flow=False,
                                            continuous=True, denoise=True, use_path=False, eps=eps)

    y = y.to(dev)
    if warmup_steps > 0:
        sampler(warmup_steps)(model, y)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = sampler(steps)(model, y)
        e1.record()
        torch.cuda.synchronize(dev)
        s = e0.elapsed_time(e1) * 1e-3
    else:
        t0 = time.perf_counter()
        out, _ = sampler(steps)(model, y)
        s = time.perf_counter() - t0
    return {"s_per_step": s / steps, "steps": steps, "batch": shape[0], "finite": bool(torch.isfinite(out).all().item()),
            "out": out}
