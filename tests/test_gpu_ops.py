"""Parity of the individual CUDA ops (through the C ABI) against the CPU oracle.

fp32 ops: tolerance 1e-5 relative to the tensor's max magnitude (different summation order).
bf16-storage ops: inputs are rounded to bf16 first; the output tolerance is one bf16 ulp (2^-8
relative to max magnitude) because the result is stored as bf16.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from golden_utils import golden
from oracle import ncsnpp as o_net
from oracle import ops as o_ops
from oracle import sampling as o_samp
from oracle import sde as o_sde

pytestmark = pytest.mark.gpu

F32 = 1e-5
BF16 = 2.0 ** -8


def K():
    from conditional_score_diffusion_b200 import kernels
    return kernels


def _close(got, ref, rtol, what):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    print(f"[ops] {what}: max_err={err:.3e} scale={scale:.3e} rel={err / scale:.2e}")
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def _upfirdn(x, k, up, down, pad):
    k_ = K()
    n, c, h, w = x.shape
    out = k_.upfirdn2d_planes(x.cuda().reshape(n * c, h, w), k.cuda(), up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.reshape(n, c, out.shape[-2], out.shape[-1])


def test_upfirdn2d_golden_cases():
    for i, c in enumerate(golden()["upfirdn2d"]):
        _close(_upfirdn(c["x"], c["k"], c["up"], c["down"], c["pad"]), c["y"], F32, f"upfirdn2d golden {i}")


@pytest.mark.parametrize("shape", [(2, 3, 160, 160), (1, 5, 80, 80), (3, 2, 20, 20), (2, 2, 10, 10), (1, 3, 5, 5),
                                   (1, 2, 37, 53), (1, 1, 130, 260)])
def test_upfirdn2d_hot_modes_vs_oracle(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    k_up = torch.tensor(o_ops.setup_kernel([1, 3, 3, 1]) * 4)
    k_dn = torch.tensor(o_ops.setup_kernel([1, 3, 3, 1]))
    k_rand = torch.rand(4, 4, generator=g)
    _close(_upfirdn(x, k_up, 2, 1, (2, 1)), o_ops.upfirdn2d(x, k_up, 2, 1, (2, 1)), F32, f"up2 {shape}")
    _close(_upfirdn(x, k_rand, 2, 1, (2, 1)), o_ops.upfirdn2d(x, k_rand, 2, 1, (2, 1)), F32, f"up2 rand-k {shape}")
    _close(_upfirdn(x, k_dn, 1, 2, (1, 1)), o_ops.upfirdn2d(x, k_dn, 1, 2, (1, 1)), F32, f"down2 {shape}")
    _close(_upfirdn(x, k_rand, 1, 2, (1, 1)), o_ops.upfirdn2d(x, k_rand, 1, 2, (1, 1)), F32, f"down2 rand-k {shape}")
    _close(_upfirdn(x, k_rand, 1, 1, (2, 2)), o_ops.upfirdn2d(x, k_rand, 1, 1, (2, 2)), F32, f"1:1 {shape}")
    # backward geometry of up2: up/down swapped, the FLIPPED kernel, g_pad = (1, 1) for k=4, pad=(2,1)
    # (op/upfirdn2d.py:25-44,110-113)
    k_flip = torch.flip(k_rand, [0, 1])
    _close(_upfirdn(x, k_flip, 1, 2, (1, 1)), o_ops.upfirdn2d(x, k_flip, 1, 2, (1, 1)), F32, f"bwd-of-up2 {shape}")


def test_upfirdn2d_generic_paths():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 11, 9, generator=g)
    for (kh, up, down, pad) in [(3, 2, 2, (1, 2)), (5, 3, 1, (2, 2)), (2, 1, 3, (0, 1)), (4, 2, 1, (-1, 3)), (1, 1, 1, (0, 0))]:
        k = torch.rand(kh, kh, generator=g)
        _close(_upfirdn(x, k, up, down, pad), o_ops.upfirdn2d(x, k, up, down, pad), F32, f"generic k{kh} up{up} down{down} pad{pad}")


def test_upfirdn2d_empty_and_errors():
    k_ = K()
    out = k_.upfirdn2d_planes(torch.empty(0, 8, 8, device="cuda"), torch.ones(4, 4, device="cuda"), 2, 2, 1, 1, 2, 1, 2, 1)
    assert out.shape == (0, 16, 16)
    from conditional_score_diffusion_b200._lib import CsdError
    with pytest.raises(CsdError):
        k_.upfirdn2d_planes(torch.ones(1, 2, 2, device="cuda"), torch.ones(9, 9, device="cuda"), 1, 1, 1, 1, 0, 0, 0, 0)


def test_fused_bias_act_vs_oracle():
    k_ = K()
    f = golden()["fused_leaky_relu"]
    out = k_.fused_bias_act(f["x"].cuda(), f["b"].cuda(), None, 3, 0, 0.2, 2 ** 0.5)
    _close(out, f["y"], F32, "fused_leaky_relu golden")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 7, 5, 4, generator=g)
    b = torch.randn(7, generator=g)
    r = torch.randn(3, 7, 5, 4, generator=g)
    for act in (1, 3):
        for grad in (0, 1, 2):
            ref = o_ops.fused_bias_act(x, b if grad == 0 else None, r, act, grad, 0.1, 1.3)
            got = k_.fused_bias_act(x.cuda(), b.cuda() if grad == 0 else None, r.cuda(), act, grad, 0.1, 1.3)
            _close(got, ref, F32, f"fused_bias_act act{act} grad{grad}")


def test_pc_update_kernels_vs_oracle():
    k_ = K()
    g = torch.Generator().manual_seed(8)
    B, shape = 4, (4, 3, 16, 16)
    x = torch.randn(*shape, generator=g) * 3
    score = torch.randn(*shape, generator=g) * 0.2
    z = torch.randn(*shape, generator=g)
    sde = o_sde.VE(0.01, 50.0, 1000)
    vp = o_sde.VP(0.1, 20.0, 1000)
    tvals = torch.linspace(1.0, 1e-5, 7)
    dev = "cuda"
    step = torch.tensor([3], dtype=torch.int32, device=dev)
    t = torch.ones(B) * tvals[3]
    xo, xm = torch.empty(shape, device=dev), torch.empty(shape, device=dev)
    # Langevin (VE: alpha = 1)
    norms = torch.empty(2 * B, device=dev)
    k_.langevin_norms(score.cuda(), z.cuda(), norms)
    ref_n = torch.cat([score.reshape(B, -1).norm(dim=-1), z.reshape(B, -1).norm(dim=-1)])
    _close(norms, ref_n, F32, "langevin norms")
    k_.langevin_update(x.cuda(), score.cuda(), z.cuda(), norms, xo, xm, 0.16, None, step)
    rx, rm = o_samp.langevin_update(sde, score, x, t, z, 0.16)
    _close(xo, rx, 2e-5, "langevin x"); _close(xm, rm, 2e-5, "langevin mean")
    # Langevin (VP: alpha table)
    alpha_tab = torch.stack([vp.alphas[(tv * (vp.N - 1)).long()] for tv in tvals]).cuda()
    k_.langevin_update(x.cuda(), score.cuda(), z.cuda(), norms, xo, xm, 0.16, alpha_tab, step)
    rx, rm = o_samp.langevin_update(vp, score, x, t, z, 0.16)
    _close(xo, rx, 2e-5, "langevin vp x")
    # reverse diffusion VE
    g_tab = torch.stack([sde.discretize_g(torch.ones(1) * tv)[0] for tv in tvals]).cuda()
    k_.reverse_diffusion_update(x.cuda(), score.cuda(), z.cuda(), xo, xm, None, g_tab, False, step)
    rx, rm = o_samp.reverse_diffusion_update(sde, score, x, t, z)
    _close(xo, rx, F32, "rd x"); _close(xm, rm, F32, "rd mean")
    k_.reverse_diffusion_update(x.cuda(), score.cuda(), None, xo, xm, None, g_tab, True, step)
    rx, rm = o_samp.reverse_diffusion_update(sde, score, x, t, z, probability_flow=True)
    _close(xo, rx, F32, "rd pf x")
    # reverse diffusion VP
    fg = [vp.discretize_fg(torch.ones(1) * tv) for tv in tvals]
    f_tab = torch.stack([a[0][0] for a in fg]).cuda()
    g_tab_vp = torch.stack([a[1][0] for a in fg]).cuda()
    k_.reverse_diffusion_update(x.cuda(), score.cuda(), z.cuda(), xo, xm, f_tab, g_tab_vp, False, step)
    rx, rm = o_samp.reverse_diffusion_update(vp, score, x, t, z)
    _close(xo, rx, F32, "rd vp x"); _close(xm, rm, F32, "rd vp mean")
    # Euler-Maruyama VE / VP
    gc = torch.stack([sde.diffusion(torch.ones(1) * tv)[0] for tv in tvals]).cuda()
    k_.euler_maruyama_update(x.cuda(), score.cuda(), z.cuda(), xo, xm, None, gc, -1.0 / sde.N, False, step)
    rx, rm = o_samp.euler_maruyama_update(sde, score, x, t, z)
    _close(xo, rx, F32, "em x"); _close(xm, rm, F32, "em mean")
    dc = torch.stack([-0.5 * vp.beta(tv) for tv in tvals]).float().cuda()
    gv = torch.stack([vp.diffusion(torch.ones(1) * tv)[0] for tv in tvals]).cuda()
    k_.euler_maruyama_update(x.cuda(), score.cuda(), z.cuda(), xo, xm, dc, gv, -1.0 / vp.N, False, step)
    rx, rm = o_samp.euler_maruyama_update(vp, score, x, t, z)
    _close(xo, rx, F32, "em vp x")
    # y perturbation + step counter
    sig_tab = torch.stack([sde.sigma(torch.ones(1) * tv)[0] for tv in tvals]).cuda()
    yp = torch.empty(shape, device=dev)
    k_.ve_perturb(x.cuda(), z.cuda(), yp, sig_tab, step)
    _close(yp, x + z * sde.sigma(t)[:, None, None, None], F32, "ve perturb")
    k_.step_advance(step)
    assert step.item() == 4
    # per-sample coefficient arrays (sample_stride = 1, no step index) and odd per-sample sizes
    xs = torch.randn(3, 3, 5, 5, generator=g)
    ss = torch.randn(3, 3, 5, 5, generator=g)
    zs = torch.randn(3, 3, 5, 5, generator=g)
    gv3 = torch.tensor([0.5, 1.5, 2.5])
    fv3 = torch.tensor([0.1, -0.2, 0.3])
    xo3, xm3 = torch.empty(3, 3, 5, 5, device=dev), torch.empty(3, 3, 5, 5, device=dev)
    k_.reverse_diffusion_update(xs.cuda(), ss.cuda(), zs.cuda(), xo3, xm3, fv3.cuda(), gv3.cuda(), False, None, 1)
    ref_m = xs - (fv3.view(3, 1, 1, 1) * xs - gv3.view(3, 1, 1, 1) ** 2 * ss)
    _close(xm3, ref_m, F32, "rd per-sample mean")
    _close(xo3, ref_m + gv3.view(3, 1, 1, 1) * zs, F32, "rd per-sample x")
    dst = torch.empty(5, device=dev)
    k_.broadcast_table(dst, torch.arange(10, dtype=torch.float32, device=dev), step, 0)
    assert (dst == 4).all()
    k_.broadcast_table(dst, torch.arange(40, dtype=torch.float32, device=dev), step, 1)
    assert dst.tolist() == [20.0, 21.0, 22.0, 23.0, 24.0]


@pytest.mark.parametrize("c0,c1,hw", [(96, 0, 24 * 24), (192, 96, 100), (288, 288, 25), (16, 0, 256), (288, 192, 400), (8, 0, 64),
                                      (288, 0, 25), (192, 0, 400), (192, 192, 400), (128, 0, 64), (576, 0, 100)])
def test_groupnorm_silu_vs_torch(c0, c1, hw):
    k_ = K()
    g = torch.Generator().manual_seed(c0 + c1)
    B = 3
    C = c0 + c1
    groups = min(C // 4, 32)
    a = (torch.randn(B, hw, c0, generator=g) * 2 + 0.5).to(torch.bfloat16)
    b = (torch.randn(B, hw, c1, generator=g) - 1.0).to(torch.bfloat16) if c1 else None
    gamma = 1 + 0.2 * torch.randn(C, generator=g)
    beta = 0.3 * torch.randn(C, generator=g)
    full = torch.cat([a, b], -1) if c1 else a
    ref_in = full.float().permute(0, 2, 1).reshape(B, C, hw, 1)
    ref = F.silu(F.group_norm(ref_in, groups, gamma, beta, eps=1e-6)).reshape(B, C, hw).permute(0, 2, 1)
    ag, bg = a.cuda(), (b.cuda() if c1 else None)
    sums0 = torch.zeros(B, c0, 2, device="cuda")
    k_.gn_chan_stats(ag, c0, sums0)
    _close(sums0[..., 0], a.float().sum(1), 1e-4, f"channel sums C0={c0}")
    _close(sums0[..., 1], (a.float() ** 2).sum(1), 1e-4, f"channel sums of squares C0={c0}")
    sums1 = None
    if c1:
        sums1 = torch.zeros(B, c1, 2, device="cuda")
        k_.gn_chan_stats(bg, c1, sums1)
    out = torch.empty(B, hw, C, device="cuda", dtype=torch.bfloat16)
    k_.gn_apply(ag, c0, sums0, bg, c1, sums1, gamma.cuda(), beta.cuda(), out, groups, 1e-6, True)
    _close(out, ref, BF16, f"gn+silu C={C} hw={hw}")
    k_.gn_apply(ag, c0, sums0, bg, c1, sums1, gamma.cuda(), beta.cuda(), out, groups, 1e-6, False)
    _close(out, F.group_norm(ref_in, groups, gamma, beta, eps=1e-6).reshape(B, C, hw).permute(0, 2, 1), BF16, "gn only")
    # one-launch form (statistics + apply) used at the small levels
    if k_.gn_fused_supported(c0, c1, hw, groups, B):
        for silu in (True, False):
            fo = torch.full((B, hw, C), float("nan"), device="cuda", dtype=torch.bfloat16)
            k_.gn_fused(ag, c0, bg, c1, gamma.cuda(), beta.cuda(), fo, groups, 1e-6, silu)
            gn = F.group_norm(ref_in, groups, gamma, beta, eps=1e-6).reshape(B, C, hw).permute(0, 2, 1)
            _close(fo, F.silu(gn) if silu else gn, BF16, f"fused gn C={C} hw={hw} silu={silu}")
    else:
        assert hw > 400, f"one-launch GroupNorm declined a small shape C={C} hw={hw}"
    # per-tile partials -> channel sums (the path fed by the transposed convolution's epilogue)
    tiles = 7
    parts = torch.randn(B * tiles, c0, 2, device="cuda")
    fin = torch.empty(B, c0, 2, device="cuda")
    k_.gn_finalize_partials(parts, fin, B, tiles, c0)
    _close(fin, parts.view(B, tiles, c0, 2).sum(1), 1e-5, "finalize partials")


@pytest.mark.parametrize("h,c", [(16, 96), (10, 288), (5, 8), (20, 192)])
def test_fir_nhwc_vs_oracle(h, c):
    k_ = K()
    g = torch.Generator().manual_seed(h * c)
    B = 2
    x = torch.randn(B, c, h, h, generator=g).to(torch.bfloat16)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    up = torch.empty(B, 2 * h, 2 * h, c, device="cuda", dtype=torch.bfloat16)
    k_.fir_resample(xn, up, "up", [1, 3, 3, 1])
    _close(up.permute(0, 3, 1, 2), o_ops.upsample_2d(x.float()), BF16, f"fir up {h} {c}")
    add = torch.randn(B, 2 * h, 2 * h, c, generator=g).to(torch.bfloat16)
    k_.fir_resample(xn, up, "up", [1, 3, 3, 1], add=add.cuda())
    _close(up.permute(0, 3, 1, 2), o_ops.upsample_2d(x.float()) + add.float().permute(0, 3, 1, 2), BF16, "fir up + add")
    if h % 2 == 0:
        dn = torch.empty(B, h // 2, h // 2, c, device="cuda", dtype=torch.bfloat16)
        k_.fir_resample(xn, dn, "down", [1, 3, 3, 1])
        _close(dn.permute(0, 3, 1, 2), o_ops.downsample_2d(x.float()), BF16, f"fir down {h} {c}")
        k_.fir_resample(xn, dn, "down", [1, 2, 3, 4])  # asymmetric taps exercise the flip
        _close(dn.permute(0, 3, 1, 2), o_ops.downsample_2d(x.float(), (1, 2, 3, 4)), BF16, "fir down asym")
        k_.fir_resample(xn, up, "up", [1, 2, 3, 4])
        _close(up.permute(0, 3, 1, 2), o_ops.upsample_2d(x.float(), (1, 2, 3, 4)), BF16, "fir up asym")


def test_fir_nhwc_more_tiles_than_resident_ctas():
    """FIR at a size with many more tiles than resident CTAs (B = 8, 80 px, 96 channels: the up x2 kernel takes whole
    96-channel pixel rows per CTA there, csrc/resample.cu launch_fir_tma), repeated launches, `add`, down x2, and the fp32
    instances on the same data."""
    k_ = K()
    g = torch.Generator().manual_seed(77)
    B, h, c = 8, 80, 96
    x = torch.randn(B, c, h, h, generator=g).to(torch.bfloat16)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    ref_up, ref_dn = o_ops.upsample_2d(x.float()), o_ops.downsample_2d(x.float())
    add = torch.randn(B, 2 * h, 2 * h, c, generator=g).to(torch.bfloat16)
    up = torch.empty(B, 2 * h, 2 * h, c, device="cuda", dtype=torch.bfloat16)
    for rep in range(2):                                        # (a second launch must not depend on stale barriers)
        k_.fir_resample(xn, up, "up", [1, 3, 3, 1])
        _close(up.permute(0, 3, 1, 2), ref_up, BF16, f"fir up many tiles, launch {rep}")
    k_.fir_resample(xn, up, "up", [1, 3, 3, 1], add=add.cuda())
    _close(up.permute(0, 3, 1, 2), ref_up + add.float().permute(0, 3, 1, 2), BF16, "fir up + add, many tiles")
    dn = torch.empty(B, h // 2, h // 2, c, device="cuda", dtype=torch.bfloat16)
    k_.fir_resample(xn, dn, "down", [1, 3, 3, 1])
    _close(dn.permute(0, 3, 1, 2), ref_dn, BF16, "fir down many tiles")
    xf = xn.float()
    upf = torch.empty(B, 2 * h, 2 * h, c, device="cuda", dtype=torch.float32)
    k_.fir_resample(xf, upf, "up", [1, 3, 3, 1])
    _close(upf.permute(0, 3, 1, 2), ref_up, 2e-5, "fir up many tiles fp32")
    dnf = torch.empty(B, h // 2, h // 2, c, device="cuda", dtype=torch.float32)
    k_.fir_resample(xf, dnf, "down", [1, 3, 3, 1])
    _close(dnf.permute(0, 3, 1, 2), ref_dn, 2e-5, "fir down many tiles fp32")


@pytest.mark.parametrize("cout,h,w,with_res", [(6, 40, 72, True), (6, 9, 5, False), (3, 33, 20, True), (8, 16, 64, True)])
def test_tap_shift_sum_vs_torch(cout, h, w, with_res):
    """csd_tap_shift_sum_bf16 (second half of the tap-stacked output heads, models/ncsnpp.py:337-352): out = bias + res +
    the nine per-tap partial maps shifted to their output pixel, zero outside the image. Ragged tiles, odd channel
    counts (two-byte tap offsets), with and without the pyramid residual."""
    k_ = K()
    g = torch.Generator().manual_seed(cout * 100 + h)
    B = 3
    pitch = -(-9 * cout // 8) * 8
    part = torch.zeros(B, h, w, pitch)
    part[..., :9 * cout] = torch.randn(B, h, w, 9 * cout, generator=g)
    part = part.to(torch.bfloat16)
    bias = torch.randn(cout, generator=g)
    res = torch.randn(B, h, w, 8, generator=g).to(torch.bfloat16) if with_res else None
    out = torch.full((B, h, w, 8), float("nan"), dtype=torch.bfloat16, device="cuda")
    k_.tap_shift_sum(part.cuda(), cout, bias.cuda(), res.cuda() if with_res else None, out)
    pf = F.pad(part.float(), (0, 0, 1, 1, 1, 1))               # [B, h + 2, w + 2, pitch]
    ref = bias.view(1, 1, 1, cout).expand(B, h, w, cout).clone()
    for t in range(9):
        ref = ref + pf[:, t // 3:t // 3 + h, t % 3:t % 3 + w, t * cout:(t + 1) * cout]
    if with_res:
        ref = ref + res.float()[..., :cout]
    _close(out[..., :cout], ref, BF16, f"tap_shift_sum cout={cout} {h}x{w} res={with_res}")
    assert (out[..., cout:].float() == 0).all(), "tap_shift_sum: padding channels must be zero"


@pytest.mark.parametrize("h,c,pitch", [(16, 96, 96), (40, 192, 192), (20, 40, 48), (6, 8, 8)])
def test_fir_with_fused_groupnorm_input_vs_oracle(h, c, pitch):
    """csd_fir_norm_resample_nhwc_*: FIR(SiLU(GroupNorm(x))) with the normalisation applied to the TMA-staged tile -
    act(GroupNorm_0(x)) -> upsample_2d / downsample_2d of ResnetBlockBigGANpp (models/layerspp.py:242-258) in one pass.
    bf16 and fp32 tensors; channel pitch > channels (padding channels stay zero); without SiLU as well."""
    k_ = K()
    g = torch.Generator().manual_seed(3 * h + c)
    B = 2
    groups = min(c // 4, 32)
    x = (torch.randn(B, c, h, h, generator=g) * 2 + 0.5).to(torch.bfloat16)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    for dtype, tol in ((torch.bfloat16, BF16), (torch.float32, 2e-5)):
        xin = x.float() if dtype == torch.float32 else x
        xn = torch.zeros(B, h, h, pitch, dtype=dtype)
        xn[..., :c] = xin.permute(0, 2, 3, 1)
        xn = xn.cuda()
        sums = torch.empty(B, c, 2, device="cuda")
        k_.gn_chan_stats(xn, c, sums)
        coef = torch.empty(B, c, 2, device="cuda")
        k_.gn_coeffs(sums, c, None, 0, gamma.cuda(), beta.cuda(), coef, None, h * h, groups, 1e-6)
        for silu in (True, False):
            gn = F.group_norm(xin.float(), groups, gamma, beta, eps=1e-6)
            act = F.silu(gn) if silu else gn
            if dtype == torch.bfloat16:
                act = act.to(torch.bfloat16).float()          # the staged tile holds the activation in bf16
            up = torch.full((B, 2 * h, 2 * h, pitch), float("nan"), device="cuda", dtype=dtype)
            k_.fir_resample(xn, up, "up", [1, 3, 3, 1], norm=coef, norm_silu=silu)
            _close(up[..., :c].permute(0, 3, 1, 2), o_ops.upsample_2d(act), tol, f"fir(norm) up {h} {c} {dtype} silu={silu}")
            assert (up[..., c:] == 0).all()
            dn = torch.full((B, h // 2, h // 2, pitch), float("nan"), device="cuda", dtype=dtype)
            k_.fir_resample(xn, dn, "down", [1, 3, 3, 1], norm=coef, norm_silu=silu)
            _close(dn[..., :c].permute(0, 3, 1, 2), o_ops.downsample_2d(act), tol, f"fir(norm) down {h} {c} {dtype}")
            assert (dn[..., c:] == 0).all()
    with pytest.raises(Exception):                            # pre-filter mode has no staged tile: refused, not ignored
        k_.fir_resample(xn, torch.empty(B, h + 1, h + 1, pitch, device="cuda"), "prefilter", [1, 3, 3, 1], norm=coef)


def test_layout_softmax_temb_dense():
    k_ = K()
    g = torch.Generator().manual_seed(21)
    B, H, W = 3, 12, 10
    x = torch.rand(B, 3, H, W, generator=g)
    y = torch.rand(B, 3, H, W, generator=g)
    out = torch.full((B, H, W, 8), 7.0, device="cuda", dtype=torch.bfloat16)
    k_.nchw_to_nhwc(x.cuda(), y.cuda(), out, 2.0, -1.0)
    ref = torch.zeros(B, H, W, 8)
    ref[..., :6] = (2 * torch.cat([x, y], 1) - 1).permute(0, 2, 3, 1)
    _close(out, ref, BF16, "nchw->nhwc")
    back = torch.empty(B, 3, H, W, device="cuda")
    rs = torch.tensor([0.5, 2.0, 3.0])
    k_.nhwc_to_nchw(out, 3, 3, back, rs.cuda())
    _close(back, out[..., 3:6].float().cpu().permute(0, 3, 1, 2) * rs.view(B, 1, 1, 1), F32, "nhwc->nchw")
    # softmax
    rows, cols, pitch = 37, 100, 104
    logits = torch.randn(rows, pitch, generator=g) * 4
    probs = torch.empty(rows, pitch, device="cuda", dtype=torch.bfloat16)
    k_.softmax_rows(logits.cuda(), probs, cols, 0.3)
    ref = torch.zeros(rows, pitch)
    ref[:, :cols] = torch.softmax(logits[:, :cols] * 0.3, -1)
    _close(probs, ref, BF16, "softmax")
    # time embedding (both kinds) + dense rows
    for emb_type, nf in (("positional", 96), ("fourier", 16), ("positional", 16)):
        embed = 2 * nf if emb_type == "fourier" else nf
        w0 = torch.randn(4 * nf, embed, generator=g) / math.sqrt(embed)
        b0 = torch.randn(4 * nf, generator=g) * 0.1
        w1 = torch.randn(4 * nf, 4 * nf, generator=g) / math.sqrt(4 * nf)
        b1 = torch.randn(4 * nf, generator=g) * 0.1
        fw = torch.randn(nf, generator=g) * 16
        labels = torch.tensor([999 * 0.73, 3.0, 0.0]) if emb_type == "positional" else torch.log(torch.tensor([3.7, 0.05, 50.0]))
        if emb_type == "fourier":
            proj = labels[:, None] * fw[None, :] * 2 * math.pi
            emb = torch.cat([torch.sin(proj), torch.cos(proj)], -1)
        else:
            emb = o_net.timestep_embedding(labels, nf)
        t = F.linear(F.silu(F.linear(emb, w0, b0)), w1, b1)
        ref = F.silu(t)
        out = torch.empty(3, 4 * nf, device="cuda")
        k_.time_embedding(labels.cuda(), nf, emb_type, fw.cuda(), w0.cuda(), b0.cuda(), w1.cuda(), b1.cuda(), out)
        # sin/cos of arguments up to ~1e3 (positional) / ~1e3 (fourier): fp32 argument rounding dominates
        _close(out, ref, 2e-3 if emb_type == "fourier" else 1e-4, f"time embedding {emb_type} nf={nf}")
        wd = torch.randn(500, 4 * nf, generator=g) / math.sqrt(4 * nf)
        bd = torch.randn(500, generator=g)
        od = torch.empty(3, 500, device="cuda")
        k_.dense_rows(ref.cuda(), wd.cuda(), bd.cuda(), od)
        _close(od, F.linear(ref, wd, bd), 1e-5, "dense rows")
    # ragged tile edges in both dimensions (tiles are 64 x 64, K chunks of 32)
    a70 = torch.randn(70, 100, generator=g)
    w70 = torch.randn(130, 100, generator=g) / 10
    o70 = torch.empty(70, 130, device="cuda")
    k_.dense_rows(a70.cuda(), w70.cuda(), None, o70)
    _close(o70, F.linear(a70, w70), 1e-5, "dense rows ragged, no bias")


# ---- module surface: the autograd Functions around the kernels (op/upfirdn2d.py:19-142, op/fused_act.py:20-97) ----
@pytest.mark.parametrize("up,down,pad", [(2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (2, 2))])
def test_op_upfirdn2d_autograd_grad_and_double_backward(up, down, pad):
    """`op.upfirdn2d` is differentiable with a defined double backward (UpFirDn2d / UpFirDn2dBackward): first and
    second derivatives against torch autograd through the oracle's plain-PyTorch restatement."""
    from conditional_score_diffusion_b200.op import upfirdn2d
    g = torch.Generator().manual_seed(11 * up + down)
    x = torch.randn(2, 3, 12, 10, generator=g)
    k = torch.rand(4, 4, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = o_ops.upfirdn2d(xr, k, up, down, pad)
    w = torch.randn(yr.shape, generator=g)
    v = torch.randn(x.shape, generator=g)
    (gr,) = torch.autograd.grad((yr * w).sum(), xr, create_graph=True)
    wr = w.clone().requires_grad_(True)
    yr2 = o_ops.upfirdn2d(xr, k, up, down, pad)
    (gr2,) = torch.autograd.grad((yr2 * wr).sum(), xr, create_graph=True)
    (ggr,) = torch.autograd.grad((gr2 * v).sum(), wr)            # d/dw <grad_x, v> = forward op applied to v

    xc = x.cuda().requires_grad_(True)
    wc = w.cuda().requires_grad_(True)
    yc = upfirdn2d(xc, k.cuda(), up=up, down=down, pad=pad)
    _close(yc.detach(), yr.detach(), F32, "op.upfirdn2d forward")
    (gc,) = torch.autograd.grad((yc * wc).sum(), xc, create_graph=True)
    _close(gc.detach(), gr.detach(), F32, "op.upfirdn2d grad")
    (ggc,) = torch.autograd.grad((gc * v.cuda()).sum(), wc)      # runs UpFirDn2dBackward.backward
    _close(ggc, ggr, F32, "op.upfirdn2d double backward")


def test_upsample_downsample_2d_module_surface():
    """models/up_or_down_sampling.py:195-257 (`upsample_2d`, `downsample_2d`) incl. their gradients."""
    from conditional_score_diffusion_b200.models import up_or_down_sampling as uds
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 10, 14, generator=g)
    for fn, ofn in ((uds.upsample_2d, o_ops.upsample_2d), (uds.downsample_2d, o_ops.downsample_2d)):
        xr = x.clone().requires_grad_(True)
        yr = ofn(xr, (1, 3, 3, 1), factor=2)
        w = torch.randn(yr.shape, generator=g)
        (gr,) = torch.autograd.grad((yr * w).sum(), xr)
        xc = x.cuda().requires_grad_(True)
        yc = fn(xc, (1, 3, 3, 1), factor=2)
        (gc,) = torch.autograd.grad((yc * w.cuda()).sum(), xc)
        _close(yc.detach(), yr.detach(), F32, f"{fn.__name__} forward")
        _close(gc, gr, F32, f"{fn.__name__} grad")


def test_fused_leaky_relu_module_autograd():
    """`FusedLeakyReLU` / `fused_leaky_relu` forward, gradients w.r.t. input and bias, and the double backward
    (op/fused_act.py:20-97) against torch autograd of leaky_relu(x + b) * scale."""
    import torch.nn.functional as F
    from conditional_score_diffusion_b200.op import FusedLeakyReLU, fused_leaky_relu
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 6, 5, 7, generator=g)
    b = torch.randn(6, generator=g)
    w = torch.randn(3, 6, 5, 7, generator=g)
    v = torch.randn(3, 6, 5, 7, generator=g)

    def ref(xx, bb):
        return F.leaky_relu(xx + bb.view(1, -1, 1, 1), 0.2) * (2 ** 0.5)

    xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = ref(xr, br)
    gxr, gbr = torch.autograd.grad((yr * wr).sum(), (xr, br), create_graph=True)
    (ggr,) = torch.autograd.grad((gxr * v).sum(), wr)

    xc, bc = x.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    wc = w.cuda().requires_grad_(True)
    yc = fused_leaky_relu(xc, bc, 0.2, 2 ** 0.5)
    _close(yc.detach(), yr.detach(), F32, "fused_leaky_relu forward")
    gxc, gbc = torch.autograd.grad((yc * wc).sum(), (xc, bc), create_graph=True)
    _close(gxc.detach(), gxr.detach(), F32, "fused_leaky_relu grad input")
    _close(gbc.detach(), gbr.detach(), F32, "fused_leaky_relu grad bias")
    (ggc,) = torch.autograd.grad((gxc * v.cuda()).sum(), wc)
    _close(ggc, ggr, F32, "fused_leaky_relu double backward")
    mod = FusedLeakyReLU(6).cuda()
    with torch.no_grad():
        mod.bias.copy_(b)
    _close(mod(x.cuda()).detach(), yr.detach(), F32, "FusedLeakyReLU module")
