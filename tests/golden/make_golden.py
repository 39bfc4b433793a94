"""Generate the golden vectors that pin the oracle (and, through it, the CUDA path).

Run in the BUILD container, where the reference is mounted read-only at /root/reference:

    python tests/golden/make_golden.py

It imports the UNMODIFIED reference modules (models.ncsnpp, sde_lib, sampling.*, op.upfirdn2d's
own CPU branch) behind three import shims for packages that are not installed here
(pytorch_lightning, ml_collections, and the JIT build that `import op` triggers), runs them on
seeded inputs and writes small fixtures next to this file. /root/reference does not exist on the
GPU box: tests only ever read the committed fixtures.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("CSD_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


# ---- shims ----------------------------------------------------------------------------------
class ConfigDict(dict):
    """Attribute-style dict standing in for ml_collections.ConfigDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def install_shims():
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
    ml = types.ModuleType("ml_collections")
    ml.ConfigDict = ConfigDict
    sys.modules["ml_collections"] = ml
    # `import op` JIT-compiles two CUDA extensions; on CPU only upfirdn2d_native / the F.leaky_relu
    # branch are used (op/upfirdn2d.py:146-149, op/fused_act.py:87-94), so the build is skipped.
    import torch.utils.cpp_extension as ext
    ext.load = lambda *a, **k: None
    sys.path.insert(0, REF)


def small_paired_config():
    """A shrunk celebA_ours_NDV_160.py in its NCSN++ form."""
    c = ConfigDict()
    c.training = ConfigDict(continuous=True)
    c.data = ConfigDict(image_size=16, effective_image_size=16, num_channels=6, centered=False)
    c.model = ConfigDict(
        name="ncsnpp_paired", nf=16, ch_mult=(1, 2, 2), num_res_blocks=1, attn_resolutions=(8, 4),
        dropout=0.1, resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1],
        skip_rescale=True, resblock_type="biggan", progressive="output_skip",
        progressive_input="input_skip", progressive_combine="sum", embedding_type="positional",
        init_scale=1.0, fourier_scale=16, nonlinearity="swish", num_scales=1000)
    return c


def small_cifar_config():
    """A shrunk ve/cifar10_ncsnpp_continuous.py (Fourier embedding, residual input pyramid)."""
    c = ConfigDict()
    c.training = ConfigDict(continuous=True)
    c.data = ConfigDict(image_size=16, effective_image_size=16, num_channels=3, centered=False)
    c.model = ConfigDict(
        name="ncsnpp", nf=16, ch_mult=(1, 2), num_res_blocks=2, attn_resolutions=(8,),
        dropout=0.1, resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1],
        skip_rescale=True, resblock_type="biggan", progressive="none", progressive_input="residual",
        progressive_combine="sum", embedding_type="fourier", init_scale=1.0, fourier_scale=16,
        nonlinearity="swish", num_scales=1000)
    return c


def config_to_plain(c):
    return {k: (config_to_plain(v) if isinstance(v, dict) else v) for k, v in c.items()}


def main():
    install_shims()
    import sde_lib
    from models import ncsnpp, utils as mutils  # noqa: F401  (registers the models)
    from models import up_or_down_sampling as uds
    from op import upfirdn2d, fused_leaky_relu
    from sampling import predictors, correctors
    from sampling.unconditional import get_pc_sampler
    from sampling.conditional import get_pc_conditional_sampler

    torch.set_num_threads(4)
    fx = {}

    # ---- upfirdn2d / FIR helpers -----------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    cases = []
    for (n, c, h, w, kh, up, down, pad) in [
        (2, 3, 8, 8, 4, 2, 1, (2, 1)),      # upsample_2d geometry (mode 3 of the reference kernel)
        (2, 3, 8, 8, 4, 1, 2, (1, 1)),      # downsample_2d geometry (mode 5)
        (1, 4, 9, 7, 4, 1, 1, (2, 2)),      # conv_downsample_2d pre-filter (mode 1), odd sizes
        (1, 2, 6, 5, 3, 2, 2, (1, 2)),      # up and down together, 3x3 kernel
        (1, 2, 5, 6, 2, 1, 1, (0, 0)),      # 2x2 kernel, no padding
        (1, 1, 7, 7, 4, 2, 1, (-1, 3)),     # negative padding crops
    ]:
        x = torch.randn(n, c, h, w, generator=g)
        k = torch.rand(kh, kh, generator=g)
        y = upfirdn2d(x, k, up=up, down=down, pad=pad)
        cases.append({"x": x, "k": k, "up": up, "down": down, "pad": pad, "y": y})
    fx["upfirdn2d"] = cases
    x = torch.randn(2, 4, 10, 10, generator=g)
    fx["resample"] = {"x": x, "up": uds.upsample_2d(x, [1, 3, 3, 1], factor=2),
                      "down": uds.downsample_2d(x, [1, 3, 3, 1], factor=2)}
    w = torch.randn(5, 4, 3, 3, generator=g)
    fx["resample"]["conv_down"] = uds.conv_downsample_2d(x, w, k=[1, 3, 3, 1])
    fx["resample"]["conv_down_w"] = w
    b = torch.randn(4, generator=g)
    fx["fused_leaky_relu"] = {"x": x, "b": b, "y": fused_leaky_relu(x, b)}

    # ---- NCSN++ forward ----------------------------------------------------------------------
    for name, cfg_fn in [("paired", small_paired_config), ("cifar", small_cifar_config)]:
        cfg = cfg_fn()
        torch.manual_seed(3)
        model = mutils.create_model(cfg).eval()
        # non-trivial GroupNorm affine / biases so that every parameter matters
        g2 = torch.Generator().manual_seed(4)
        with torch.no_grad():
            for pn, p in model.named_parameters():
                if pn.endswith("bias") or pn.endswith(".b"):
                    p.copy_(0.1 * torch.randn(p.shape, generator=g2))
                elif "GroupNorm" in pn and pn.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g2))
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        B = 2
        hw = cfg.data.image_size
        if name == "paired":
            x = torch.randn(B, 3, hw, hw, generator=g2) * 2.0
            y = torch.rand(B, 3, hw, hw, generator=g2)
            labels = torch.tensor([999.0 * 0.73, 999.0 * 0.11])
            with torch.no_grad():
                out = model({"x": x, "y": y}, labels)
            fx[f"ncsnpp_{name}"] = {"config": config_to_plain(cfg), "state_dict": sd, "x": x, "y": y,
                                    "labels": labels, "out_x": out["x"], "out_y": out["y"]}
        else:
            x = torch.rand(B, 3, hw, hw, generator=g2)
            labels = torch.log(torch.tensor([3.7, 0.05]))
            with torch.no_grad():
                out = model(x, labels)
            fx[f"ncsnpp_{name}"] = {"config": config_to_plain(cfg), "state_dict": sd, "x": x,
                                    "labels": labels, "out": out}

        # ---- samplers on that network --------------------------------------------------------
        if name == "cifar":
            sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=10)
            sampler = get_pc_sampler(sde, (B, 3, hw, hw), predictors.get_predictor("reverse_diffusion"),
                                     correctors.get_corrector("langevin"), snr=0.16, p_steps=4, c_steps=1,
                                     continuous=True, denoise=True, eps=1e-5)
            torch.manual_seed(1234)
            samples, info = sampler(model, show_evolution=True)
            fx["pc_unconditional"] = {"samples": samples, "evolution": info["evolution"],
                                      "sigma_min": 0.01, "sigma_max": 50.0, "N": 10, "snr": 0.16,
                                      "p_steps": 4, "eps": 1e-5, "seed": 1234}
            # single updates with injected tensors (noise is drawn inside: seed right before)
            t = torch.tensor([0.6, 0.6])
            xs = torch.randn(B, 3, hw, hw, generator=g2) * 5
            score_fn = mutils.get_score_fn(sde, model, conditional=False, train=False, continuous=True)
            with torch.no_grad():
                torch.manual_seed(77)
                px, pmean = predictors.ReverseDiffusionPredictor(sde, score_fn).update_fn(xs, t)
                torch.manual_seed(77)
                ex, emean = predictors.EulerMaruyamaPredictor(sde, score_fn).update_fn(xs, t)
                torch.manual_seed(77)
                cx, cmean = correctors.LangevinCorrector(sde, score_fn, 0.16, 1).update_fn(xs, t)
                score = score_fn(xs, t)
            fx["single_updates"] = {"x": xs, "t": t, "score": score, "seed": 77, "rd_x": px, "rd_mean": pmean,
                                    "em_x": ex, "em_mean": emean, "lc_x": cx, "lc_mean": cmean}
        else:
            sde = {"x": sde_lib.cVESDE(sigma_min=5e-3, sigma_max=float(np.sqrt(3 * hw * hw)), N=1000),
                   "y": sde_lib.VESDE(sigma_min=5e-3, sigma_max=0.5, N=1000)}
            sampler = get_pc_conditional_sampler(
                sde, (B, 3, hw, hw), predictors.get_predictor("conditional_reverse_diffusion"),
                correctors.get_corrector("conditional_langevin"), snr=0.15, p_steps=3, c_steps=1,
                continuous=True, denoise=True, use_path=False, eps=1e-5)
            torch.manual_seed(4321)
            samples, info = sampler(model, y, show_evolution=True)
            fx["pc_conditional"] = {"y": y, "samples": samples, "evolution_x": info["evolution"]["x"],
                                    "sigma_min_x": 5e-3, "sigma_max_x": float(np.sqrt(3 * hw * hw)),
                                    "sigma_min_y": 5e-3, "sigma_max_y": 0.5, "N": 1000, "snr": 0.15,
                                    "p_steps": 3, "eps": 1e-5, "seed": 4321}

    # ---- SDE tables -------------------------------------------------------------------------------
    ve = sde_lib.VESDE(0.01, 50, 1000)
    vp = sde_lib.VPSDE(0.1, 20, 1000)
    t = torch.tensor([1.0, 0.5, 1e-3, 1e-5])
    xz = torch.zeros(4, 1, 2, 2)
    fx["sde"] = {
        "t": t,
        "ve_std": ve.marginal_prob(xz, t)[1], "ve_g": ve.sde(xz, t)[1], "ve_G": ve.discretize(xz, t)[1],
        "vp_mean_coeff": vp.marginal_prob(torch.ones(4, 1, 2, 2), t)[0][:, 0, 0, 0],
        "vp_std": vp.marginal_prob(xz, t)[1], "vp_g": vp.sde(xz, t)[1],
        "vp_f": vp.discretize(torch.ones(4, 1, 2, 2), t)[0][:, 0, 0, 0], "vp_G": vp.discretize(xz, t)[1],
    }

    path = os.path.join(OUT, "reference_vectors.pt")
    torch.save(fx, path)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
