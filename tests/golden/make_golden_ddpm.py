"""Golden vectors for the DDPM U-Net family (models/ddpm.py), generated from the UNMODIFIED reference.

Run in the BUILD container (reference mounted at /root/reference):  python tests/golden/make_golden_ddpm.py
Writes tests/golden/reference_vectors_ddpm.pt. Same shims as make_golden.py.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, ConfigDict, config_to_plain, install_shims  # noqa: E402


def small_ddpm_config(name, out_ch):
    """A shrunk celebA_ours_NDV_160.py as shipped (model.name='ddpm_paired') / edges2shoes_SR3.py ('ddpm_paired_SR3')."""
    c = ConfigDict()
    c.training = ConfigDict(continuous=True)
    c.data = ConfigDict(image_size=16, effective_image_size=16, num_channels=6, centered=False)
    c.model = ConfigDict(name=name, nf=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), dropout=0.1,
                         resamp_with_conv=True, conditional=True, nonlinearity="swish", input_channels=6,
                         output_channels=out_ch, num_scales=1000)
    return c


def main():
    install_shims()
    from models import ddpm, utils as mutils  # noqa: F401  (registers the models)
    torch.set_num_threads(4)
    fx = {}
    sd_paired = None
    for name, out_ch in (("ddpm_paired", 6), ("ddpm_paired_SR3", 3)):
        cfg = small_ddpm_config(name, out_ch)
        torch.manual_seed(5)
        model = mutils.create_model(cfg).eval()
        g = torch.Generator().manual_seed(6)
        with torch.no_grad():
            for pn, p in model.named_parameters():
                if pn.endswith("bias") or pn.endswith(".b"):
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
                elif "GroupNorm" in pn and pn.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                elif p.abs().max() < 1e-6:       # init_scale=0 layers (1e-10): give them real weights
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
                # bf16-representable parameters: the fixture stores them losslessly in half the bytes
                p.copy_(p.to(torch.bfloat16).to(torch.float32))
        if sd_paired is None:
            sd_paired = {k: v.detach().clone() for k, v in model.state_dict().items()}
        else:
            # the SR3 variant is the same network with a 3-channel head: reuse the paired weights (head sliced), so
            # the fixture carries one state dict
            last = max(int(k.split(".")[1]) for k in sd_paired)
            sd = {k: (v[:out_ch] if k.startswith(f"all_modules.{last}.") else v) for k, v in sd_paired.items()}
            model.load_state_dict(sd, strict=True)
        B, hw = 2, cfg.data.image_size
        x = torch.randn(B, 3, hw, hw, generator=g) * 2.0
        y = torch.rand(B, 3, hw, hw, generator=g)
        labels = torch.tensor([999.0 * 0.61, 999.0 * 0.07])
        with torch.no_grad():
            out = model({"x": x, "y": y}, labels)
        rec = {"config": config_to_plain(cfg), "x": x, "y": y, "labels": labels}
        if name == "ddpm_paired":
            rec["state_dict_bf16"] = {k: v.to(torch.bfloat16) for k, v in sd_paired.items()}
        if isinstance(out, dict):
            rec["out_x"], rec["out_y"] = out["x"].clone(), out["y"].clone()
        else:
            rec["out"] = out.clone()
        fx[name] = rec
    # ---- evaluation losses (losses.py:99-234) on those networks, RNG draws replayed and stored ----------------
    import numpy as np
    import losses as ref_losses
    import sde_lib
    hw, B = 16, 2
    smax = float(np.sqrt(3 * hw * hw))
    g = torch.Generator().manual_seed(8)
    xb = torch.rand(B, 3, hw, hw, generator=g)
    yb = torch.rand(B, 3, hw, hw, generator=g)
    models = {}
    for name, out_ch in (("ddpm_paired", 6), ("ddpm_paired_SR3", 3)):
        cfg = small_ddpm_config(name, out_ch)
        mdl = mutils.create_model(cfg).eval()
        last = max(int(k.split(".")[1]) for k in sd_paired)
        mdl.load_state_dict({k: (v[:out_ch] if k.startswith(f"all_modules.{last}.") else v)
                             for k, v in sd_paired.items()}, strict=True)
        models[name] = mdl
    eps = 1e-5
    # SR3 estimator: one conditional SDE, only x perturbed
    sde = sde_lib.cVESDE(sigma_min=5e-3, sigma_max=smax, N=1000)
    rec = {"x": xb, "y": yb, "sigma_min": 5e-3, "sigma_max": smax, "eps": eps}
    for lw in (True, False):
        fn = ref_losses.get_general_sde_loss_fn(sde, train=False, conditional=True, reduce_mean=True, continuous=True,
                                                likelihood_weighting=lw, eps=eps)
        torch.manual_seed(123)
        with torch.no_grad():
            rec[f"loss_lw{int(lw)}"] = fn(models["ddpm_paired_SR3"], (yb, xb)).clone()
    torch.manual_seed(123)
    rec["t"] = torch.rand(B) * (sde.T - eps) + eps
    rec["z"] = torch.randn_like(xb)
    fx["loss_sr3"] = rec
    # CMDE estimator: x and y SDEs
    sdes = {"x": sde_lib.cVESDE(sigma_min=5e-3, sigma_max=smax, N=1000), "y": sde_lib.VESDE(sigma_min=5e-3, sigma_max=0.5, N=1000)}
    rec = {"x": xb, "y": yb, "sigma_max_x": smax, "sigma_max_y": 0.5, "sigma_min": 5e-3, "eps": eps}
    for rm in (True, False):
        fn = ref_losses.get_general_sde_loss_fn(sdes, train=False, conditional=True, reduce_mean=rm, continuous=True,
                                                likelihood_weighting=True, eps=eps)
        torch.manual_seed(321)
        with torch.no_grad():
            rec[f"loss_rm{int(rm)}"] = fn(models["ddpm_paired"], (yb, xb)).clone()
    torch.manual_seed(321)
    rec["t"] = torch.rand(B) * (1 - eps) + eps
    rec["z_y"] = torch.randn_like(yb)
    rec["z_x"] = torch.randn_like(xb)
    fx["loss_cmde"] = rec

    path = os.path.join(OUT, "reference_vectors_ddpm.pt")
    torch.save(fx, path)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
