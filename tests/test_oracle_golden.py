"""Pin the oracle (oracle/*.py) against vectors produced by the real reference
(tests/golden/make_golden.py). CPU only. Tolerances: fp32 arithmetic in a different
association order -> 1e-5 relative to the tensor's max magnitude unless noted."""
import torch

from golden_utils import golden, to_namespace
from oracle import ncsnpp as o_net
from oracle import ops as o_ops
from oracle import sampling as o_samp
from oracle import sde as o_sde


def _close(a, b, rtol=1e-5, what=""):
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def test_upfirdn2d_cases():
    for i, c in enumerate(golden()["upfirdn2d"]):
        y = o_ops.upfirdn2d(c["x"], c["k"], up=c["up"], down=c["down"], pad=c["pad"])
        assert y.shape == c["y"].shape, (i, y.shape, c["y"].shape)
        _close(y, c["y"], what=f"upfirdn2d case {i}")


def test_fir_helpers():
    r = golden()["resample"]
    _close(o_ops.upsample_2d(r["x"]), r["up"], what="upsample_2d")
    _close(o_ops.downsample_2d(r["x"]), r["down"], what="downsample_2d")
    _close(o_ops.conv_downsample_2d(r["x"], r["conv_down_w"]), r["conv_down"], what="conv_downsample_2d")


def test_fused_leaky_relu():
    f = golden()["fused_leaky_relu"]
    _close(o_ops.fused_leaky_relu(f["x"], f["b"]), f["y"], what="fused_leaky_relu")


def test_sde_tables():
    s = golden()["sde"]
    t = s["t"]
    ve, vp = o_sde.VE(0.01, 50, 1000), o_sde.VP(0.1, 20, 1000)
    _close(ve.sigma(t), s["ve_std"], what="ve std")
    _close(ve.diffusion(t), s["ve_g"], what="ve g")
    _close(ve.discretize_g(t), s["ve_G"], what="ve G")
    mean_c, std = vp.marginal(t)
    _close(mean_c, s["vp_mean_coeff"], what="vp mean")
    _close(std, s["vp_std"], what="vp std")
    _close(vp.diffusion(t), s["vp_g"], what="vp g")
    fc, G = vp.discretize_fg(t)
    _close(fc, s["vp_f"], rtol=1e-4, what="vp f")
    _close(G, s["vp_G"], what="vp G")


def _net(name):
    f = golden()[f"ncsnpp_{name}"]
    o = o_net.model_options(to_namespace(f["config"]))
    return f, o


def test_ncsnpp_paired_forward():
    f, o = _net("paired")
    out = o_net.forward_paired(f["state_dict"], o, f["x"], f["y"], f["labels"])
    _close(out["x"], f["out_x"], rtol=2e-5, what="paired out x")
    _close(out["y"], f["out_y"], rtol=2e-5, what="paired out y")


def test_ncsnpp_cifar_forward():
    f, o = _net("cifar")
    out = o_net.forward(f["state_dict"], o, f["x"], f["labels"])
    _close(out, f["out"], rtol=2e-5, what="cifar out")


def test_spec_matches_state_dict():
    for name in ("paired", "cifar"):
        f, o = _net(name)
        spec = o_net.build_spec(o)
        idx = {int(k.split(".")[1]) for k in f["state_dict"]}
        assert max(idx) < len(spec)


def test_single_updates():
    f, o = _net("cifar")
    s = golden()["single_updates"]
    sde = o_sde.VE(0.01, 50, 10)
    score_fn = o_sde.score_fn_unconditional(lambda x, l: o_net.forward(f["state_dict"], o, x, l), sde, True, "fourier")
    score = score_fn(s["x"], s["t"])
    _close(score, s["score"], rtol=5e-5, what="score")
    torch.manual_seed(s["seed"])
    z = torch.randn_like(s["x"])
    x1, m1 = o_samp.reverse_diffusion_update(sde, score, s["x"], s["t"], z)
    _close(x1, s["rd_x"], what="rd x"); _close(m1, s["rd_mean"], what="rd mean")
    x2, m2 = o_samp.euler_maruyama_update(sde, score, s["x"], s["t"], z)
    _close(x2, s["em_x"], what="em x"); _close(m2, s["em_mean"], what="em mean")
    x3, m3 = o_samp.langevin_update(sde, score, s["x"], s["t"], z, 0.16)
    _close(x3, s["lc_x"], rtol=1e-4, what="lc x"); _close(m3, s["lc_mean"], rtol=1e-4, what="lc mean")


def test_pc_unconditional_trajectory():
    f, o = _net("cifar")
    p = golden()["pc_unconditional"]
    sde = o_sde.VE(p["sigma_min"], p["sigma_max"], p["N"])
    score_fn = o_sde.score_fn_unconditional(lambda x, l: o_net.forward(f["state_dict"], o, x, l), sde, True, "fourier")
    torch.manual_seed(p["seed"])
    rec = []
    samples, info = o_samp.pc_sampler(score_fn, sde, tuple(f["x"].shape), p["snr"], p["p_steps"], 1, eps=p["eps"],
                                      record=rec)
    _close(torch.stack(rec), p["evolution"], rtol=1e-3, what="evolution")
    _close(samples, p["samples"], rtol=1e-3, what="samples")
    assert info["steps"] == p["p_steps"] * 2


def test_pc_conditional_trajectory():
    f, o = _net("paired")
    p = golden()["pc_conditional"]
    sx = o_sde.VE(p["sigma_min_x"], p["sigma_max_x"], p["N"])
    sy = o_sde.VE(p["sigma_min_y"], p["sigma_max_y"], p["N"])
    model_fn = lambda d, l: o_net.forward_paired(f["state_dict"], o, d["x"], d["y"], l)
    score_fn = o_sde.score_fn_conditional_pair(model_fn, sx, sy, True)
    torch.manual_seed(p["seed"])
    rec = []
    samples, _ = o_samp.pc_conditional_sampler(score_fn, sx, sy, p["y"], tuple(p["y"].shape), p["snr"], p["p_steps"],
                                               1, eps=p["eps"], record=rec)
    _close(torch.stack(rec), p["evolution_x"], rtol=1e-3, what="evolution")
    _close(samples, p["samples"], rtol=1e-3, what="samples")
