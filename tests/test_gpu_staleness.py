"""Packed-operand staleness: the engine keeps bf16/tf32 K-major copies of the weights and captured CUDA
graphs; every way the reference's callers change weights must reach them.

* load_state_dict / optimizer steps bump autograd's version counters            -> detected by NetEngine.ensure_packed
* the reference's EMA swaps through `param.data.copy_` (models/ema.py:111,149),
  DDP-style broadcasts write through `.data` as well: no version bump          -> inference entry points re-pack always
* re-homed parameters (.to(), FusedAdamEMA's flat buffer) rebuild the plans    -> cached FusedPCSampler must re-plan
"""
import pytest
import torch

from golden_utils import golden, to_namespace

pytestmark = pytest.mark.gpu


def _model(seed=None):
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    f = golden()["ncsnpp_paired"]
    m = utils.create_model(to_namespace(f["config"]))
    m.load_state_dict(f["state_dict"], strict=True)
    if seed is not None:
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    return f, m.cuda().eval()


def _sampler(shape, p):
    from conditional_score_diffusion_b200 import sampling, sde_lib
    sde = {"x": sde_lib.cVESDE(p["sigma_min_x"], p["sigma_max_x"], p["N"]),
           "y": sde_lib.VESDE(p["sigma_min_y"], p["sigma_max_y"], p["N"])}
    return sampling.get_pc_conditional_sampler(sde, shape, sampling.get_predictor("conditional_reverse_diffusion"),
                                               sampling.get_corrector("conditional_langevin"), p["snr"], 3, 1,
                                               continuous=True, denoise=True, eps=p["eps"])


def _noise(shape, steps=3):
    g = torch.Generator().manual_seed(3)
    noise = {(n, i): torch.randn(*shape, generator=g) for i in range(steps) for n in ("y_c", "x_c", "y_p", "x_p")}
    x0 = torch.randn(*shape, generator=g) * 10.0
    return noise, x0


def test_cached_sampler_follows_load_state_dict_and_data_writes():
    p = golden()["pc_conditional"]
    shape = tuple(p["y"].shape)
    noise, x0 = _noise(shape)
    src = lambda n, i, k: noise[(n, i)]
    f, m = _model()
    sampler = _sampler(shape, p)
    a0, _ = sampler(m, p["y"].cuda(), x_init=x0, noise_source=src)
    # new weights through load_state_dict
    _, other = _model(seed=9)
    m.load_state_dict(other.state_dict())
    a1, _ = sampler(m, p["y"].cuda(), x_init=x0, noise_source=src)
    fresh, _ = _sampler(shape, p)(other, p["y"].cuda(), x_init=x0, noise_source=src)
    assert not torch.allclose(a0, a1)
    assert torch.allclose(a1, fresh, rtol=0, atol=1e-3 * fresh.abs().max().item())
    # back to the first weights through `.data` (no version bump), exactly as the reference's EMA does
    sd = f["state_dict"]
    for name, q in m.named_parameters():
        q.data.copy_(sd[name].cuda())
    a2, _ = sampler(m, p["y"].cuda(), x_init=x0, noise_source=src)
    assert torch.allclose(a2, a0, rtol=0, atol=1e-3 * a0.abs().max().item())


def test_cached_sampler_replans_after_parameters_move():
    p = golden()["pc_conditional"]
    shape = tuple(p["y"].shape)
    noise, x0 = _noise(shape)
    src = lambda n, i, k: noise[(n, i)]
    f, m = _model()
    sampler = _sampler(shape, p)
    a0, _ = sampler(m, p["y"].cuda(), x_init=x0, noise_source=src)
    # re-home every parameter (what optim.FusedAdamEMA / .to() do): new storage, the engine drops its plans
    with torch.no_grad():
        for q in m.parameters():
            q.data = q.data.clone()
    _, other = _model(seed=5)
    for (name, q), v in zip(m.named_parameters(), other.parameters()):
        q.data.copy_(v.data)
    a1, _ = sampler(m, p["y"].cuda(), x_init=x0, noise_source=src)
    fresh, _ = _sampler(shape, p)(other, p["y"].cuda(), x_init=x0, noise_source=src)
    assert not torch.allclose(a0, a1)
    assert torch.allclose(a1, fresh, rtol=0, atol=1e-3 * fresh.abs().max().item())


def test_reference_style_ema_swap_reaches_the_forward():
    """ema.store / copy_to / forward / restore / forward with BOTH EMA flavours: this package's (version-bumping) and
    a reference-style one writing through .data."""
    from conditional_score_diffusion_b200.models.ema import ExponentialMovingAverage
    f, m = _model()
    _, other = _model(seed=21)
    x, y, labels = f["x"].cuda(), f["y"].cuda(), f["labels"].cuda()
    with torch.no_grad():
        base = m({"x": x, "y": y}, labels)["x"]
        want = other({"x": x, "y": y}, labels)["x"]
    ema = ExponentialMovingAverage(m.parameters(), 0.999, module=m)
    with torch.no_grad():
        for s, q in zip(ema.shadow_params, other.parameters()):
            s.copy_(q)
    ema.store()
    ema.copy_to()
    with torch.no_grad():
        got = m({"x": x, "y": y}, labels)["x"]
    assert torch.allclose(got, want, rtol=0, atol=1e-3 * want.abs().max().item())
    ema.restore()
    with torch.no_grad():
        back = m({"x": x, "y": y}, labels)["x"]
    assert torch.allclose(back, base, rtol=0, atol=1e-3 * base.abs().max().item())
    # the reference's way: param.data.copy_ (models/ema.py:111)
    for s, q in zip(ema.shadow_params, m.parameters()):
        q.data.copy_(s.data)
    with torch.no_grad():
        got2 = m({"x": x, "y": y}, labels)["x"]
    assert torch.allclose(got2, want, rtol=0, atol=1e-3 * want.abs().max().item())
