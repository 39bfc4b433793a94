"""The oracle against the reference's outputs for the NCSN++ construction variants (fir=False, resblock_type='ddpm',
progressive_combine='cat', 'residual' input pyramid without FIR) - tests/golden/reference_vectors_variants.pt, written by
tests/golden/make_golden_variants.py from the unmodified reference. Variants the REFERENCE itself cannot run are listed in
the fixture's `broken` table (layerspp.Upsample(fir=False) raises, models/layerspp.py:116-117)."""
import os

import torch

from golden_utils import to_namespace
from oracle import ncsnpp as o_net
from oracle import ops as o_ops

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_variants.pt")


def variants():
    return torch.load(_PATH, map_location="cpu", weights_only=False)


def test_oracle_matches_reference_on_every_runnable_variant():
    fx = variants()
    assert set(fx["nets"]) == {"fir_false_biggan", "ddpm_blocks_fir_noconv", "combine_cat", "residual_input_nofir"}
    for name, f in fx["nets"].items():
        cfg = to_namespace(f["config"])
        out = o_net.forward(f["state_dict"], o_net.model_options(cfg), f["x"], f["labels"])
        err = (out - f["out"]).abs().max().item() / f["out"].abs().max().item()
        assert err < 1e-5, f"{name}: {err:.3e}"


def test_reference_broken_variants_are_recorded():
    fx = variants()
    assert set(fx["broken"]) == {"fir_false_pyramids", "ddpm_blocks_nofir_conv", "ddpm_blocks_nofir_noconv"}
    assert all("scale_factor" in v for v in fx["broken"].values())


def test_oracle_helpers_match_reference():
    h = variants()["helpers"]
    x, w = h["x"], h["w"]
    assert torch.allclose(o_ops.upsample_2d(x, (1, 1)), h["naive_up"], atol=1e-6)
    assert torch.allclose(o_ops.downsample_2d(x, (1, 1)), h["naive_down"], atol=1e-6)
    assert torch.allclose(o_ops.conv_downsample_2d(x, w, (1, 3, 3, 1)), h["conv_down"], atol=1e-5)
