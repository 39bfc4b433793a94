"""Golden GRADIENTS of the training losses, generated from the UNMODIFIED reference with PyTorch autograd on CPU.

Run in the BUILD container (reference mounted at /root/reference):  python tests/golden/make_golden_grads.py
Reads the networks (configs + state dicts) already stored in reference_vectors.pt / reference_vectors_ddpm.pt, so the
fixture written here (tests/golden/reference_grads.pt) only carries the random draws, the loss values and the
parameter / input gradients (bf16: half the bytes; the parity tolerance is far above one bf16 ulp).

Cases (dropout = 0 so that train mode is deterministic; the reference's RNG draws are replayed and stored):
  cmde   ncsnpp_paired, two-SDE CMDE loss (losses.py:119-146), reduce_mean=True, likelihood weighting
  uncond ncsnpp (Fourier embedding, residual input pyramid), losses.py:207-232 without likelihood weighting, plus the
         input gradient d sum(score * eps) / dx of likelihood.get_div_fn (likelihood.py:26-37)
  sr3    ddpm_paired_SR3, SR3 loss (losses.py:185-206) with likelihood weighting
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, ConfigDict, install_shims  # noqa: E402


def to_config(d):
    c = ConfigDict()
    for k, v in d.items():
        setattr(c, k, to_config(v) if isinstance(v, dict) else v)
    return c


def grads_of(model):
    return {n: p.grad.detach().to(torch.bfloat16) for n, p in model.named_parameters() if p.grad is not None}


def main():
    install_shims()
    import losses as ref_losses
    import sde_lib
    from models import ddpm, ncsnpp, utils as mutils  # noqa: F401
    torch.set_num_threads(4)
    base = torch.load(os.path.join(OUT, "reference_vectors.pt"), weights_only=False)
    base_ddpm = torch.load(os.path.join(OUT, "reference_vectors_ddpm.pt"), weights_only=False)
    fx = {}
    eps = 1e-5
    hw, B = 16, 2
    smax = float(np.sqrt(3 * hw * hw))
    g = torch.Generator().manual_seed(21)
    xb = torch.rand(B, 3, hw, hw, generator=g)
    yb = torch.rand(B, 3, hw, hw, generator=g)

    def build(rec, sd, name=None, out_ch=None):
        cfg = to_config(rec["config"])
        cfg.model.dropout = 0.0
        if name is not None:
            cfg.model.name = name
            cfg.model.output_channels = out_ch
        model = mutils.create_model(cfg)
        model.load_state_dict(sd, strict=True)
        return model

    # ---- CMDE on ncsnpp_paired ----
    rec = base["ncsnpp_paired"]
    model = build(rec, rec["state_dict"])
    sdes = {"x": sde_lib.cVESDE(sigma_min=5e-3, sigma_max=smax, N=1000),
            "y": sde_lib.VESDE(sigma_min=5e-3, sigma_max=0.5, N=1000)}
    fn = ref_losses.get_general_sde_loss_fn(sdes, train=True, conditional=True, reduce_mean=True, continuous=True,
                                            likelihood_weighting=True, eps=eps)
    torch.manual_seed(501)
    loss = fn(model, (yb, xb))
    loss.backward()
    torch.manual_seed(501)
    t = torch.rand(B) * (1 - eps) + eps
    z_y = torch.randn_like(yb)
    z_x = torch.randn_like(xb)
    fx["cmde"] = {"x": xb, "y": yb, "t": t, "z_x": z_x, "z_y": z_y, "loss": loss.detach().clone(),
                  "grads": grads_of(model), "sigma_max_x": smax, "sigma_max_y": 0.5, "sigma_min": 5e-3, "eps": eps}

    # ---- unconditional on ncsnpp (cifar-like) + input gradient of the Hutchinson estimator ----
    rec = base["ncsnpp_cifar"]
    model = build(rec, rec["state_dict"])
    sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=1000)
    fn = ref_losses.get_sde_loss_fn(sde, train=True, reduce_mean=True, continuous=True, likelihood_weighting=False, eps=eps)
    torch.manual_seed(502)
    loss = fn(model, xb)
    loss.backward()
    torch.manual_seed(502)
    t = torch.rand(B) * (sde.T - eps) + eps
    z = torch.randn_like(xb)
    out = {"x": xb, "t": t, "z": z, "loss": loss.detach().clone(), "grads": grads_of(model), "sigma_min": 0.01,
           "sigma_max": 50.0, "eps": eps}
    model.zero_grad()
    score_fn = mutils.get_score_fn(sde, model, conditional=False, train=False, continuous=True)
    xs = (xb + 0.3 * torch.randn(xb.shape, generator=g)).requires_grad_(True)
    ts = torch.tensor([0.4, 0.9])
    probe = torch.randint(0, 2, xb.shape, generator=g).float() * 2 - 1
    with torch.enable_grad():
        s = score_fn(xs, ts)
        gx = torch.autograd.grad(torch.sum(s * probe), xs)[0]
    out.update(div_x=xs.detach().clone(), div_t=ts, div_eps=probe, div_score=s.detach().clone(), div_grad=gx.clone())
    fx["uncond"] = out

    # ---- likelihood (likelihood.py:40-113) and the probability-flow ODE sampler (sampling/unconditional.py:93-158) on
    #      the same network; loose solver tolerances keep the run short, the Hutchinson probe is stored ----
    import likelihood as ref_likelihood
    from sampling.unconditional import get_ode_sampler
    model.eval()
    sde_l = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=1000)
    like_fn = ref_likelihood.get_likelihood_fn(sde_l, lambda v: (v + 1.0) / 2.0, rtol=1e-3, atol=1e-3, eps=1e-5)
    torch.manual_seed(504)
    bpd, z_lat, nfe = like_fn(model, xb)
    torch.manual_seed(504)
    probe_l = torch.randint_like(xb, low=0, high=2).float() * 2 - 1.0
    fx["likelihood"] = {"x": xb, "epsilon": probe_l, "bpd": bpd.clone(), "z": z_lat.clone(), "nfe": nfe, "rtol": 1e-3,
                        "atol": 1e-3, "eps": 1e-5, "sigma_min": 0.01, "sigma_max": 50.0}
    ode = get_ode_sampler(sde_l, tuple(xb.shape), denoise=True, rtol=1e-3, atol=1e-3, eps=1e-3)
    z0 = torch.randn(xb.shape, generator=g) * 50.0
    xs_ode, nfe_ode = ode(model, z=z0.clone())
    fx["ode_sampler"] = {"z": z0, "samples": xs_ode.clone(), "nfe": nfe_ode, "rtol": 1e-3, "atol": 1e-3, "eps": 1e-3}

    # ---- inpainting (sampling/unconditional.py:230-345) on the unconditional network, 4 noise levels ----
    from sampling import predictors, correctors
    from sampling.unconditional import get_pc_inpainter
    from sampling.conditional import get_pc_conditional_sampler
    sde_i = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=4)
    inpainter = get_pc_inpainter(sde_i, predictors.get_predictor("reverse_diffusion"), correctors.get_corrector("langevin"),
                                 snr=0.16, n_steps=1, continuous=True, denoise=True, eps=1e-5)
    mask = torch.ones_like(xb)
    mask[:, :, 4:12, 4:12] = 0.0
    torch.manual_seed(505)
    xi, info_i = inpainter(model, xb, mask, show_evolution=True)
    fx["inpaint"] = {"data": xb, "mask": mask, "samples": xi.clone(), "evolution": info_i["evolution"].clone(), "seed": 505,
                     "snr": 0.16, "N": 4, "sigma_min": 0.01, "sigma_max": 50.0, "eps": 1e-5}
    # ---- use_path conditional sampler (sampling/conditional.py:87-94,124-176) on ncsnpp_paired ----
    recp = base["ncsnpp_paired"]
    model_p = build(recp, recp["state_dict"]).eval()
    sdes_p = {"x": sde_lib.cVESDE(sigma_min=5e-3, sigma_max=smax, N=1000),
              "y": sde_lib.VESDE(sigma_min=5e-3, sigma_max=0.5, N=1000)}
    sampler_p = get_pc_conditional_sampler(sdes_p, (B, 3, hw, hw), predictors.get_predictor("conditional_reverse_diffusion"),
                                           correctors.get_corrector("conditional_langevin"), snr=0.15, p_steps=3, c_steps=1,
                                           continuous=True, denoise=True, use_path=True, eps=1e-5)
    torch.manual_seed(506)
    xp, info_p = sampler_p(model_p, yb, show_evolution=True)
    fx["pc_use_path"] = {"y": yb, "samples": xp.clone(), "evolution_x": info_p["evolution"]["x"].clone(),
                         "evolution_y": info_p["evolution"]["y"].clone(), "seed": 506, "snr": 0.15, "p_steps": 3,
                         "sigma_min": 5e-3, "sigma_max_x": smax, "sigma_max_y": 0.5, "eps": 1e-5}

    # ---- legacy discrete losses (losses.py:55-85 SMLD, :320-340 DDPM), evaluation mode, on the unconditional network ----
    vesde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=1000)
    vpsde = sde_lib.VPSDE(beta_min=0.1, beta_max=20, N=1000)
    leg = {"x": xb}
    for rm in (False, True):
        torch.manual_seed(507)
        with torch.no_grad():
            leg[f"smld_rm{int(rm)}"] = ref_losses.get_smld_loss_fn(vesde, train=False, reduce_mean=rm)(model, xb).clone()
        torch.manual_seed(508)
        with torch.no_grad():
            leg[f"ddpm_rm{int(rm)}"] = ref_losses.get_ddpm_loss_fn(vpsde, train=False, reduce_mean=rm)(model, xb).clone()
    torch.manual_seed(507)
    leg["smld_labels"] = torch.randint(0, 1000, (B,))
    leg["smld_z"] = torch.randn_like(xb)
    torch.manual_seed(508)
    leg["ddpm_labels"] = torch.randint(0, 1000, (B,))
    leg["ddpm_z"] = torch.randn_like(xb)
    fx["legacy_losses"] = leg

    # ---- SR3 on ddpm_paired_SR3 ----
    rec = base_ddpm["ddpm_paired"]
    sd_paired = {k: v.float() for k, v in rec["state_dict_bf16"].items()}
    last = max(int(k.split(".")[1]) for k in sd_paired)
    sd = {k: (v[:3] if k.startswith(f"all_modules.{last}.") else v) for k, v in sd_paired.items()}
    model = build(base_ddpm["ddpm_paired_SR3"], sd)
    sde = sde_lib.cVESDE(sigma_min=5e-3, sigma_max=smax, N=1000)
    fn = ref_losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                            likelihood_weighting=True, eps=eps)
    torch.manual_seed(503)
    loss = fn(model, (yb, xb))
    loss.backward()
    torch.manual_seed(503)
    t = torch.rand(B) * (sde.T - eps) + eps
    z = torch.randn_like(xb)
    fx["sr3"] = {"x": xb, "y": yb, "t": t, "z": z, "loss": loss.detach().clone(), "grads": grads_of(model),
                 "sigma_max": smax, "sigma_min": 5e-3, "eps": eps}

    path = os.path.join(OUT, "reference_grads.pt")
    torch.save(fx, path)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB")
    print("likelihood bpd", fx["likelihood"]["bpd"], "nfe", fx["likelihood"]["nfe"], "ode nfe", fx["ode_sampler"]["nfe"])
    for k, v in fx.items():
        if "grads" not in v:
            continue
        gn = sum(float(gg.float().pow(2).sum()) for gg in v["grads"].values()) ** 0.5
        print(k, "loss", float(v["loss"]), "params with grad", len(v["grads"]), "grad norm", gn)


if __name__ == "__main__":
    main()
