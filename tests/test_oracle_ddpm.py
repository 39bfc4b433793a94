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
