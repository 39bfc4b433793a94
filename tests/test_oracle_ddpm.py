"""Pin the DDPM oracle (oracle/ddpm.py) against vectors produced by the real reference
(tests/golden/make_golden_ddpm.py). CPU only; 1e-5 relative to the tensor's max magnitude."""
import os

import torch

from golden_utils import to_namespace
from oracle import ddpm as o_ddpm

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_ddpm.pt")


def ddpm_golden():
    fx = torch.load(_PATH, map_location="cpu", weights_only=False)
    sd = {k: v.float() for k, v in fx["ddpm_paired"]["state_dict_bf16"].items()}
    last = max(int(k.split(".")[1]) for k in sd)
    sd_sr3 = {k: (v[:3] if k.startswith(f"all_modules.{last}.") else v) for k, v in sd.items()}
    return fx, sd, sd_sr3


def _close(a, b, rtol, what):
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def test_ddpm_paired_forward():
    fx, sd, _ = ddpm_golden()
    f = fx["ddpm_paired"]
    o = o_ddpm.model_options(to_namespace(f["config"]))
    out = o_ddpm.forward_paired(sd, o, f["x"], f["y"], f["labels"])
    _close(out["x"], f["out_x"], 1e-5, "ddpm_paired x")
    _close(out["y"], f["out_y"], 1e-5, "ddpm_paired y")


def test_ddpm_paired_sr3_forward():
    fx, _, sd3 = ddpm_golden()
    f = fx["ddpm_paired_SR3"]
    o = o_ddpm.model_options(to_namespace(f["config"]))
    out = o_ddpm.forward_paired_sr3(sd3, o, f["x"], f["y"], f["labels"])
    assert out.shape == f["out"].shape == (2, 3, 16, 16)
    _close(out, f["out"], 1e-5, "ddpm_paired_SR3")


def test_ddpm_module_surface():
    """Same registry names, constructor and state-dict keys as the reference (checkpoints load unchanged)."""
    fx, sd, sd3 = ddpm_golden()
    from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401
    for name, weights in (("ddpm_paired", sd), ("ddpm_paired_SR3", sd3)):
        model = utils.create_model(to_namespace(fx[name]["config"]))
        missing, unexpected = model.load_state_dict(weights, strict=True)
        assert not missing and not unexpected
        assert [k for k in model.state_dict()] == list(weights)


def _score_fns(fx, sd, sd3):
    from oracle import sde as o_sde
    cfg_p, cfg_s = to_namespace(fx["ddpm_paired"]["config"]), to_namespace(fx["ddpm_paired_SR3"]["config"])
    o_p, o_s = o_ddpm.model_options(cfg_p), o_ddpm.model_options(cfg_s)
    ls, lc = fx["loss_sr3"], fx["loss_cmde"]
    sx = o_sde.VE(ls["sigma_min"], ls["sigma_max"], 1000)
    sy = o_sde.VE(lc["sigma_min"], lc["sigma_max_y"], 1000)

    def sr3_score(d, t):      # get_score_fn, conditional cVESDE branch (models/utils.py:207-221): labels = t (N-1)
        out = o_ddpm.forward_paired_sr3(sd3, o_s, d["x"], d["y"], t * 999)
        return out / sx.sigma(t)[:, None, None, None]

    def pair_score(d, t):     # dict branch (models/utils.py:172-186)
        out = o_ddpm.forward_paired(sd, o_p, d["x"], d["y"], t * 999)
        return {"x": out["x"] / sx.sigma(t)[:, None, None, None], "y": out["y"] / sy.sigma(t)[:, None, None, None]}

    return sx, sy, sr3_score, pair_score


def test_loss_oracle_matches_reference():
    from oracle import losses as o_loss
    fx, sd, sd3 = ddpm_golden()
    sx, sy, sr3_score, pair_score = _score_fns(fx, sd, sd3)
    ls, lc = fx["loss_sr3"], fx["loss_cmde"]
    for lw in (True, False):
        got = o_loss.sr3_loss(sr3_score, sx, ls["y"], ls["x"], ls["t"], ls["z"], True, lw)
        ref = ls[f"loss_lw{int(lw)}"]
        assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()), (lw, got.item(), ref.item())
    for rm in (True, False):
        got = o_loss.cmde_loss(pair_score, sx, sy, lc["y"], lc["x"], lc["t"], lc["z_x"], lc["z_y"], rm)
        ref = lc[f"loss_rm{int(rm)}"]
        assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()), (rm, got.item(), ref.item())


def test_training_loss_has_no_cpu_fallback():
    """The training loss runs on the CUDA path only: CPU tensors raise instead of falling back to eager PyTorch."""
    import pytest
    from conditional_score_diffusion_b200 import losses, sde_lib
    from conditional_score_diffusion_b200._lib import CsdError
    fn = losses.get_general_sde_loss_fn(sde_lib.cVESDE(5e-3, 27.7, 1000), train=True, conditional=True)
    with pytest.raises(CsdError):
        fn(None, (torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4)))


def test_lightning_checkpoint_keys_load():
    """A Lightning checkpoint of the reference ({'state_dict': {'score_model.all_modules...': ...}}) loads unchanged."""
    fx, sd, _ = ddpm_golden()
    from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401
    model = utils.create_model(to_namespace(fx["ddpm_paired"]["config"]))
    ckpt = {"state_dict": {"score_model." + k: v for k, v in sd.items()}, "hyper_parameters": {"config": None},
            "epoch": 3}
    missing, unexpected = utils.load_lightning_checkpoint(model, ckpt)
    assert not missing and not unexpected
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k])
