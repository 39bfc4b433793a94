"""Golden vectors for the NCSN++ construction variants (run in the build container, like make_golden.py):

    python tests/golden/make_golden_variants.py

Imports the UNMODIFIED reference (same shims as make_golden.py) and evaluates small NCSNpp networks for the options the
round-1 engine refused: fir=False (the DDPM++ configs, configs/vp/cifar10_ddpmpp_continuous.py:38-60),
resblock_type='ddpm' with and without resampling convolutions, progressive_combine='cat'; plus the module-surface helpers
naive_upsample_2d / naive_downsample_2d / conv_downsample_2d (models/up_or_down_sampling.py:59-69,144-178).
Writes tests/golden/reference_vectors_variants.pt.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def variant_config(**over):
    c = mg.small_paired_config()
    c.model.name = "ncsnpp"
    c.model.nf = 8                      # keeps the committed fixture small (4 state dicts)
    c.data.num_channels = 3
    for k, v in over.items():
        c.model[k] = v
    return c


VARIANTS = {
    "fir_false_biggan": dict(fir=False, progressive="none", progressive_input="none"),
    "fir_false_pyramids": dict(fir=False),
    "ddpm_blocks_fir_noconv": dict(resblock_type="ddpm", resamp_with_conv=False),
    "ddpm_blocks_nofir_conv": dict(resblock_type="ddpm", fir=False, resamp_with_conv=True),
    "ddpm_blocks_nofir_noconv": dict(resblock_type="ddpm", fir=False, resamp_with_conv=False, progressive="none",
                                     progressive_input="none"),
    "combine_cat": dict(progressive_combine="cat"),
    "residual_input_nofir": dict(fir=False, progressive_input="residual", progressive="none"),
}


def main():
    mg.install_shims()
    from models import ncsnpp, utils as mutils  # noqa: F401
    from models import up_or_down_sampling as uds
    torch.set_num_threads(4)
    fx = {"nets": {}, "broken": {}}
    for name, over in VARIANTS.items():
        cfg = variant_config(**over)
        torch.manual_seed(abs(hash(name)) % 1000)
        m = mutils.create_model(cfg) if hasattr(mutils, "create_model") else None
        m = mutils.get_model(cfg.model.name)(cfg).eval()
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for pn, p in m.named_parameters():
                if pn.endswith("bias") or pn.endswith(".b"):
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
        x = torch.randn(2, 3, 16, 16, generator=g) * 3
        labels = torch.tensor([800.0, 90.0])
        try:
            with torch.no_grad():
                out = m(x, labels)
        except Exception as e:  # noqa: BLE001 - record that the REFERENCE cannot run this variant
            fx["broken"][name] = f"{type(e).__name__}: {e}"
            print(name, "reference fails:", fx["broken"][name])
            continue
        fx["nets"][name] = {"config": mg.config_to_plain(cfg), "state_dict": {k: v.clone() for k, v in m.state_dict().items()},
                            "x": x, "labels": labels, "out": out}
        print(name, tuple(out.shape), float(out.abs().max()))
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 4, 10, 12, generator=g)
    w = torch.randn(6, 4, 3, 3, generator=g) * 0.2
    fx["helpers"] = {"x": x, "w": w, "naive_up": uds.naive_upsample_2d(x), "naive_down": uds.naive_downsample_2d(x),
                     "conv_down": uds.conv_downsample_2d(x, w, k=(1, 3, 3, 1))}
    torch.save(fx, os.path.join(HERE, "reference_vectors_variants.pt"))


if __name__ == "__main__":
    main()
