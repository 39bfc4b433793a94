"""EMA (models/ema.py) and Lightning-checkpoint round trip - host logic, no GPU.

The EMA update is checked against the reference's formula restated inline (models/ema.py:64-93:
decay_t = min(decay, (1 + n) / (10 + n)); s <- s - (1 - decay_t)(s - p)) and, when baseline/_ref is present, against
the unmodified reference class itself."""
import os
import sys

import torch

from golden_utils import golden, to_namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    f = golden()["ncsnpp_paired"]
    m = utils.create_model(to_namespace(f["config"]))
    m.load_state_dict(f["state_dict"], strict=True)
    return m, f


def test_ema_update_matches_reference_formula_and_swaps_bump_versions():
    from conditional_score_diffusion_b200.models.ema import ExponentialMovingAverage
    m, _ = _model()
    params = [p for p in m.parameters() if p.requires_grad]
    ema = ExponentialMovingAverage(m.parameters(), 0.999, module=m)
    shadow_ref = [p.detach().clone() for p in params]
    g = torch.Generator().manual_seed(0)
    for n in range(1, 4):
        with torch.no_grad():
            for p in params:
                p.add_(0.01 * torch.randn(p.shape, generator=g))
        ema.update()
        d = min(0.999, (1 + n) / (10 + n))
        for s, p in zip(shadow_ref, params):
            s.sub_((1.0 - d) * (s - p.detach()))
    for s, r in zip(ema.shadow_params, shadow_ref):
        assert torch.allclose(s, r, rtol=1e-6, atol=1e-7)
    # store / copy_to / restore: values swap, and every write bumps the autograd version (the engine's staleness signal)
    live = [p.detach().clone() for p in params]
    v0 = [p._version for p in params]
    ema.store()
    ema.copy_to()
    assert all(p._version > v for p, v in zip(params, v0))
    for p, s in zip(params, ema.shadow_params):
        assert torch.equal(p.detach(), s)
    v1 = [p._version for p in params]
    ema.restore()
    assert all(p._version > v for p, v in zip(params, v1))
    for p, l in zip(params, live):
        assert torch.equal(p.detach(), l)
    assert m._engine.param_version is None      # invalidate() was called through the module handle


def test_ema_matches_unmodified_reference_class_when_installed():
    ref_dir = os.path.join(ROOT, "baseline", "_ref", "models", "ema.py")
    if not os.path.exists(ref_dir):
        import pytest
        pytest.skip("baseline/_ref not installed in this checkout")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ema", ref_dir)
    ref_ema = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ema)
    from conditional_score_diffusion_b200.models.ema import ExponentialMovingAverage
    m, _ = _model()
    params = [p for p in m.parameters() if p.requires_grad]
    a = ExponentialMovingAverage(m.parameters(), 0.9999)
    b = ref_ema.ExponentialMovingAverage(m.parameters(), 0.9999)
    g = torch.Generator().manual_seed(1)
    for _ in range(5):
        with torch.no_grad():
            for p in params:
                p.add_(0.02 * torch.randn(p.shape, generator=g))
        a.update()
        b.update()
    assert a.num_updates == b.num_updates
    for s, r in zip(a.shadow_params, b.shadow_params):
        assert torch.allclose(s, r, rtol=1e-6, atol=1e-7)
    # state dicts are interchangeable
    a.load_state_dict(b.state_dict())
    for s, r in zip(a.shadow_params, b.shadow_params):
        assert torch.equal(s, r)


def test_lightning_checkpoint_round_trip_keeps_ema_and_config(tmp_path):
    from conditional_score_diffusion_b200.models import utils
    from conditional_score_diffusion_b200.models.ema import ExponentialMovingAverage
    m, f = _model()
    ema = ExponentialMovingAverage(m.parameters(), 0.999, module=m)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
    ema.update()
    path = str(tmp_path / "ckpt.ckpt")
    ckpt = utils.save_lightning_checkpoint(m, path, config=f["config"], ema=ema, extra={"global_step": 7})
    assert all(k.startswith("score_model.all_modules.") for k in ckpt["state_dict"])
    assert ckpt["global_step"] == 7
    m2, _ = _model()
    ema2 = ExponentialMovingAverage(m2.parameters(), 0.5, module=m2)
    missing, unexpected = utils.load_lightning_checkpoint(m2, path, ema=ema2)
    assert not missing and not unexpected
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    assert ema2.decay == 0.999 and ema2.num_updates == 1
    for a, b in zip(ema.shadow_params, ema2.shadow_params):
        assert torch.equal(a.cpu(), b.cpu())
    assert utils.checkpoint_config(path) == f["config"]
    # a reference checkpoint (no ema_state): the EMA is re-seeded from the loaded weights
    del ckpt["ema_state"]
    ema3 = ExponentialMovingAverage(m2.parameters(), 0.999)
    utils.load_lightning_checkpoint(m2, ckpt, ema=ema3)
    for s, p in zip(ema3.shadow_params, [p for p in m2.parameters() if p.requires_grad]):
        assert torch.equal(s, p.detach())
