"""DDPM U-Net family on the B200 engine: parity against the real reference's outputs (golden vectors), against
the CPU oracle on shapes that reach the transposed tensor-core kernel, and through the fused PC sampler.
Tolerance as for NCSN++ (tests/test_gpu_network.py): 2e-2 of the output maximum (bf16 operands)."""
import math

import pytest
import torch

from golden_utils import to_namespace
from oracle import ddpm as o_ddpm
from oracle import sampling as o_samp
from oracle import sde as o_sde
from test_gpu_network import _check
from test_oracle_ddpm import ddpm_golden

pytestmark = pytest.mark.gpu


def _model(cfg, sd):
    from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401
    m = utils.create_model(cfg)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def test_ddpm_paired_matches_reference_golden():
    fx, sd, _ = ddpm_golden()
    f = fx["ddpm_paired"]
    m = _model(to_namespace(f["config"]), sd)
    with torch.no_grad():
        out = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
        out2 = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())   # graph replay
    _check(out["x"], f["out_x"], "ddpm_paired x vs reference")
    _check(out["y"], f["out_y"], "ddpm_paired y vs reference")
    _check(out2["x"], f["out_x"], "ddpm_paired x (graph replay)")


def test_ddpm_paired_sr3_matches_reference_golden():
    fx, _, sd3 = ddpm_golden()
    f = fx["ddpm_paired_SR3"]
    m = _model(to_namespace(f["config"]), sd3)
    with torch.no_grad():
        out = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
    assert out.shape == (2, 3, 16, 16)
    _check(out, f["out"], "ddpm_paired_SR3 vs reference")


def _big_cfg():
    """64x64, nf 32: the 64 px and 32 px levels run in the persistent transposed kernel (fused GroupNorm(32)
    prologue, NIN shortcut / identity residual as K segments), the 16 px level in the per-tap kernel."""
    fx, _, _ = ddpm_golden()
    cfg = to_namespace(fx["ddpm_paired"]["config"])
    cfg.data.image_size = cfg.data.effective_image_size = 64
    cfg.model.ch_mult = (1, 2, 2)
    cfg.model.attn_resolutions = (16,)
    return cfg


def test_ddpm_transposed_levels_match_oracle():
    from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401
    cfg = _big_cfg()
    torch.manual_seed(21)
    m = utils.create_model(cfg)
    g = torch.Generator().manual_seed(22)
    with torch.no_grad():
        for pn, p in m.named_parameters():
            if pn.endswith("bias") or pn.endswith(".b"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif p.abs().max() < 1e-6:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o = o_ddpm.model_options(cfg)
    B = 3
    x = torch.randn(B, 3, 64, 64, generator=g) * 3
    y = torch.rand(B, 3, 64, 64, generator=g)
    labels = torch.rand(B, generator=g) * 999
    ref = o_ddpm.forward_paired(sd, o, x, y, labels)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m({"x": x.cuda(), "y": y.cuda()}, labels.cuda())
    _check(out["x"], ref["x"], "ddpm 64px x vs oracle")
    _check(out["y"], ref["y"], "ddpm 64px y vs oracle")


def test_ddpm_conditional_pc_sampler_matches_oracle():
    from conditional_score_diffusion_b200 import sampling, sde_lib
    fx, sd, _ = ddpm_golden()
    f = fx["ddpm_paired"]
    cfg = to_namespace(f["config"])
    m = _model(cfg, sd)
    o = o_ddpm.model_options(cfg)
    shape = tuple(f["y"].shape)
    smax = math.sqrt(3 * 16 * 16)
    g = torch.Generator().manual_seed(31)
    steps = 2
    noise = {(n, i): torch.randn(*shape, generator=g) for i in range(steps) for n in ("y_c", "x_c", "y_p", "x_p")}
    order = [(n, i) for i in range(steps) for n in ("y_c", "x_c", "y_p", "x_p")]
    pos = [0]

    def seq(like):
        t = noise[order[pos[0]]]
        pos[0] += 1
        return t

    x0 = torch.randn(*shape, generator=g) * smax
    sx, sy = o_sde.VE(5e-3, smax, 1000), o_sde.VE(5e-3, 0.5, 1000)
    model_fn = lambda d, l: o_ddpm.forward_paired(sd, o, d["x"], d["y"], l)
    ref, _ = o_samp.pc_conditional_sampler(o_sde.score_fn_conditional_pair(model_fn, sx, sy, True), sx, sy, f["y"],
                                           shape, 0.15, steps, 1, eps=1e-5, randn_like=seq, x_init=x0)
    sde = {"x": sde_lib.cVESDE(5e-3, smax, 1000), "y": sde_lib.VESDE(5e-3, 0.5, 1000)}
    sampler = sampling.get_pc_conditional_sampler(sde, shape, sampling.get_predictor("conditional_reverse_diffusion"),
                                                  sampling.get_corrector("conditional_langevin"), 0.15, steps, 1,
                                                  continuous=True, denoise=True, eps=1e-5)
    got, _ = sampler(m, f["y"].cuda(), x_init=x0, noise_source=lambda n, i, k: noise[(n, i)])
    _check(got, ref, "ddpm_paired 2-step conditional PC vs oracle")


def test_eval_losses_match_reference_golden():
    """losses.get_general_sde_loss_fn (train=False) on the CUDA path against the reference's own loss values, with
    the reference's random draws (t, z) injected. Loss values are sums of squares of ~1e3..1e4 terms: bf16 network
    error of ~1e-3 relative shows up as <= 1e-2 relative in the loss."""
    from conditional_score_diffusion_b200 import losses, sde_lib
    fx, sd, sd3 = ddpm_golden()
    ls, lc = fx["loss_sr3"], fx["loss_cmde"]
    m3 = _model(to_namespace(fx["ddpm_paired_SR3"]["config"]), sd3)
    mp = _model(to_namespace(fx["ddpm_paired"]["config"]), sd)
    sde = sde_lib.cVESDE(ls["sigma_min"], ls["sigma_max"], 1000)
    for lw in (True, False):
        fn = losses.get_general_sde_loss_fn(sde, train=False, conditional=True, reduce_mean=True, continuous=True,
                                            likelihood_weighting=lw, eps=ls["eps"])
        got = fn(m3, (ls["y"].cuda(), ls["x"].cuda()), noise={"t": ls["t"], "z": ls["z"]}).item()
        ref = ls[f"loss_lw{int(lw)}"].item()
        print(f"[loss] sr3 lw={lw}: got {got:.6e} ref {ref:.6e}")
        assert abs(got - ref) <= 1e-2 * abs(ref)
    sdes = {"x": sde_lib.cVESDE(lc["sigma_min"], lc["sigma_max_x"], 1000),
            "y": sde_lib.VESDE(lc["sigma_min"], lc["sigma_max_y"], 1000)}
    for rm in (True, False):
        fn = losses.get_general_sde_loss_fn(sdes, train=False, conditional=True, reduce_mean=rm, continuous=True,
                                            likelihood_weighting=True, eps=lc["eps"])
        got = fn(mp, (lc["y"].cuda(), lc["x"].cuda()), noise={"t": lc["t"], "z_x": lc["z_x"], "z_y": lc["z_y"]}).item()
        ref = lc[f"loss_rm{int(rm)}"].item()
        print(f"[loss] cmde reduce_mean={rm}: got {got:.6e} ref {ref:.6e}")
        assert abs(got - ref) <= 1e-2 * abs(ref)
    # without injected noise the draws come from torch's generators: finite, and different across calls
    a = fn(mp, (lc["y"].cuda(), lc["x"].cuda())).item()
    b = fn(mp, (lc["y"].cuda(), lc["x"].cuda())).item()
    assert math.isfinite(a) and math.isfinite(b) and a != b
