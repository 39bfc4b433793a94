"""Whole-network parity: the engine-backed NCSNpp against (a) outputs of the real reference
(golden vectors) and (b) the CPU oracle on fresh inputs.

Precision contract: the reference computes in fp32; this path stores activations and feeds the tensor
cores in bf16 (fp32 accumulation, fp32 statistics). Per stored tensor that is a relative rounding of
2^-9; through the ~30-layer golden nets the measured error stays below 1% of the output's max
magnitude, so the tolerance asserted here is 2e-2 for both max-abs / max-ref and relative L2.
"""
import pytest
import torch

from golden_utils import golden, to_namespace
from oracle import ncsnpp as o_net

pytestmark = pytest.mark.gpu

MAX_REL = 2e-2
L2_REL = 2e-2


def _model(name):
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    f = golden()[f"ncsnpp_{name}"]
    m = utils.create_model(to_namespace(f["config"]))
    m.load_state_dict(f["state_dict"], strict=True)
    return f, m.cuda().eval()


def _check(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    mx = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
    l2 = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
    print(f"[net] {what}: max_rel={mx:.3e} l2_rel={l2:.3e} ref_max={ref.abs().max().item():.3e}")
    assert mx < MAX_REL and l2 < L2_REL, f"{what}: max_rel={mx:.3e} l2_rel={l2:.3e}"


def test_paired_forward_matches_reference_golden():
    f, m = _model("paired")
    with torch.no_grad():
        out = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
    _check(out["x"], f["out_x"], "paired x vs reference")
    _check(out["y"], f["out_y"], "paired y vs reference")
    # second and third call: CUDA-graph replay path, fresh output tensors
    with torch.no_grad():
        out2 = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
        out3 = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
    assert out2["x"].data_ptr() != out3["x"].data_ptr()
    _check(out3["x"], f["out_x"], "paired x (graph replay)")
    # run-to-run: no float atomics on the path (fixed-order statistics) -> bitwise identical
    assert torch.equal(out2["x"], out3["x"])


def test_cifar_forward_matches_reference_golden():
    f, m = _model("cifar")
    with torch.no_grad():
        out = m(f["x"].cuda(), f["labels"].cuda())
    _check(out, f["out"], "cifar vs reference")


def test_paired_forward_other_batch_and_oracle():
    f, m = _model("paired")
    o = o_net.model_options(to_namespace(f["config"]))
    g = torch.Generator().manual_seed(99)
    B = 5
    x = torch.randn(B, 3, 16, 16, generator=g) * 10
    y = torch.rand(B, 3, 16, 16, generator=g)
    labels = torch.rand(B, generator=g) * 999
    ref = o_net.forward_paired(f["state_dict"], o, x, y, labels)
    with torch.no_grad():
        out = m({"x": x.cuda(), "y": y.cuda()}, labels.cuda())
    _check(out["x"], ref["x"], "paired x vs oracle B=5")
    _check(out["y"], ref["y"], "paired y vs oracle B=5")


def test_forward_scaled_fuses_sigma_division():
    f, m = _model("paired")
    inv = {"x": torch.tensor([0.5, 4.0]), "y": torch.tensor([2.0, 0.25])}
    with torch.no_grad():
        out = m.forward_scaled({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda(),
                               {k: v.cuda() for k, v in inv.items()})
    _check(out["x"], f["out_x"] * inv["x"].view(2, 1, 1, 1), "scaled x")
    _check(out["y"], f["out_y"] * inv["y"].view(2, 1, 1, 1), "scaled y")


def test_unsupported_modes_raise_loudly():
    f, m = _model("paired")
    with pytest.raises(NotImplementedError):       # train-mode dropout has no inference plan: no silent skip
        with torch.no_grad():
            m.train()({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
    m.eval()
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m.cpu()({"x": f["x"], "y": f["y"]}, f["labels"])


@pytest.mark.parametrize("head_mode,defer", [(0, True), (1, True), (2, True), (3, True), (0, False)])
def test_paired_forward_transposed_levels_vs_oracle(monkeypatch, head_mode, defer):
    """64x64, nf 32: the 64 px and 32 px levels run in the persistent transposed kernel (fused GroupNorm+SiLU
    prologue, skip / identity-residual K segments, epilogue GroupNorm sums); everything else as in the tiny nets.
    Parametrised over the engine's A/B switches: output heads in the per-tap kernel (0) or in the transposed kernel
    with the pyramid as identity segment (1) / added by its FIR pass (2) / tap-stacked 1x1 convolution + shifted sum
    (3, the default); per-tile statistics reduced by the
    coefficient kernel (defer) or by their own finalize launch."""
    from conditional_score_diffusion_b200 import engine
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    monkeypatch.setattr(engine, "HEAD_MODE", head_mode)
    monkeypatch.setattr(engine.BlockOps, "defer_finalize", defer)
    f = golden()["ncsnpp_paired"]
    cfg = to_namespace(f["config"])
    cfg.data.image_size = cfg.data.effective_image_size = 64
    cfg.model.nf = 32
    cfg.model.attn_resolutions = (16,)
    torch.manual_seed(41)
    m = utils.create_model(cfg)
    g = torch.Generator().manual_seed(42)
    with torch.no_grad():
        for pn, p in m.named_parameters():
            if pn.endswith("bias") or pn.endswith(".b"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o = o_net.model_options(cfg)
    B = 3
    x = torch.randn(B, 3, 64, 64, generator=g) * 5
    y = torch.rand(B, 3, 64, 64, generator=g)
    labels = torch.rand(B, generator=g) * 999
    ref = o_net.forward_paired(sd, o, x, y, labels)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m({"x": x.cuda(), "y": y.cuda()}, labels.cuda())
    _check(out["x"], ref["x"], "paired 64px x vs oracle")
    _check(out["y"], ref["y"], "paired 64px y vs oracle")


def test_sr_wrappers_squeeze_and_resize():
    """ncsnpp_2xSR / ncsnpp_KxSR (models/ncsnpp.py:403-449): the squeeze is pixel_unshuffle's permutation, the wrappers
    feed the same network as ncsnpp_paired (identical all_modules / state dict) with x squeezed resp. y resized."""
    import torch.nn.functional as F
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    from conditional_score_diffusion_b200.models.engine_net import SqueezeBlock
    z = torch.randn(2, 3, 8, 6)
    sq = SqueezeBlock()
    assert torch.equal(sq(z), F.pixel_unshuffle(z, 2))
    assert torch.equal(sq(sq(z), reverse=True), z)
    f = golden()["ncsnpp_paired"]
    cfg = to_namespace(f["config"])
    cfg.data.num_channels = 15               # 4*3 squeezed x channels + 3 y channels
    torch.manual_seed(7)
    cfg.model.name = "ncsnpp_2xSR"
    m2 = utils.create_model(cfg).cuda().eval()
    cfg.model.name = "ncsnpp_paired"
    mp = utils.create_model(cfg)
    mp.load_state_dict(m2.state_dict(), strict=True)
    mp = mp.cuda().eval()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 32, 32, generator=g).cuda()
    y = torch.rand(2, 3, 16, 16, generator=g).cuda()
    labels = torch.tensor([300.0, 40.0]).cuda()
    with torch.no_grad():
        o2 = m2({"x": x, "y": y}, labels)
        op = mp({"x": F.pixel_unshuffle(x, 2), "y": y}, labels)
    assert o2["x"].shape == x.shape and o2["y"].shape == y.shape
    assert torch.allclose(o2["x"], F.pixel_shuffle(op["x"], 2)) and torch.allclose(o2["y"], op["y"])
    # KxSR: y at low resolution, resized to the target before the network and back after it
    cfg = to_namespace(f["config"])
    cfg.data.target_resolution, cfg.data.scale = 16, 2
    cfg.model.name = "ncsnpp_KxSR"
    mk = utils.create_model(cfg)
    mk.load_state_dict(f["state_dict"], strict=True)
    mk = mk.cuda().eval()
    _, mref = _model("paired")
    ylo = torch.rand(2, 3, 8, 8, generator=g).cuda()
    with torch.no_grad():
        ok = mk({"x": f["x"].cuda(), "y": ylo}, f["labels"].cuda())
        oref = mref({"x": f["x"].cuda(), "y": mk.resize_to_GT(ylo)}, f["labels"].cuda())
    assert ok["y"].shape == ylo.shape
    assert torch.allclose(ok["x"], oref["x"]) and torch.allclose(ok["y"], mk.resize_to_LQ(oref["y"]))
