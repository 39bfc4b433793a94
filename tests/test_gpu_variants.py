"""NCSN++ construction variants on the engine: fir=False (DDPM++ configs), resblock_type='ddpm',
progressive_combine='cat', the 'residual' input pyramid without FIR - against the unmodified reference's outputs
(tests/golden/reference_vectors_variants.pt) - and the module-surface helpers naive_upsample_2d / naive_downsample_2d /
conv_downsample_2d. Variants the reference itself cannot run must raise here too."""
import pytest
import torch

from golden_utils import to_namespace
from test_oracle_variants import variants

pytestmark = pytest.mark.gpu


def _model(cfg, sd=None):
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    m = utils.create_model(cfg)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["fir_false_biggan", "ddpm_blocks_fir_noconv", "combine_cat", "residual_input_nofir"])
@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_variant_matches_reference(name, precision):
    f = variants()["nets"][name]
    m = _model(to_namespace(f["config"]), f["state_dict"])
    if precision == "tf32":
        m.set_precision("tf32")
    with torch.no_grad():
        out = m(f["x"].cuda(), f["labels"].cuda())
    err = (out.cpu() - f["out"]).abs().max().item() / f["out"].abs().max().item()
    print(f"[variants {precision}] {name}: rel={err:.3e}")
    assert err < (2e-2 if precision == "bf16" else 2e-3)


def test_variants_the_reference_cannot_run_raise():
    from conditional_score_diffusion_b200._lib import CsdError
    fx = variants()
    base = fx["nets"]["fir_false_biggan"]
    for name, over in [("fir_false_pyramids", dict(progressive="output_skip", progressive_input="input_skip")),
                       ("ddpm_blocks_nofir", dict(resblock_type="ddpm"))]:
        cfg = to_namespace(base["config"])
        for k, v in over.items():
            setattr(cfg.model, k, v)
        m = _model(cfg)
        with pytest.raises(CsdError):
            with torch.no_grad():
                m(base["x"].cuda(), base["labels"].cuda())


def test_resampling_helpers_match_reference():
    from conditional_score_diffusion_b200.models import up_or_down_sampling as uds
    h = variants()["helpers"]
    x, w = h["x"].cuda(), h["w"].cuda()
    assert torch.allclose(uds.naive_upsample_2d(x).cpu(), h["naive_up"], atol=1e-5)
    assert torch.allclose(uds.naive_downsample_2d(x).cpu(), h["naive_down"], atol=1e-5)
    got = uds.conv_downsample_2d(x, w, k=(1, 3, 3, 1)).cpu()
    err = (got - h["conv_down"]).abs().max().item() / h["conv_down"].abs().max().item()
    assert err < 1e-2, err                       # bf16 tensor-core operands
    with pytest.raises(NotImplementedError):
        uds.upsample_conv_2d(x, w)
    xg = x.clone().requires_grad_(True)
    uds.naive_upsample_2d(xg).sum().backward()   # the naive helpers stay differentiable (they are the CUDA op)
    assert torch.allclose(xg.grad, torch.full_like(xg, 4.0))
