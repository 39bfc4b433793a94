"""Parity of the backward (training) kernels, through the C ABI, against torch autograd of the same op in fp32.

The reference has no hand-written backward: every gradient comes from `loss.backward()` (losses.py:345-407), i.e.
from ATen/cuDNN autograd of nn.Conv2d, nn.GroupNorm, nn.SiLU, F.softmax and upfirdn2d's autograd.Function
(op/upfirdn2d.py:19-142). Each test feeds bf16-rounded inputs to both sides. Tolerances:
 - weight gradients (bf16 operands, fp32 accumulation over up to 1.6 M pixels): 2^-8 of the tensor's max;
 - activation gradients stored as bf16: 2^-7 of the tensor's max (one bf16 rounding of the result plus the bf16
   rounding of intermediate du inside the GroupNorm backward);
 - fp32 kernels: 1e-5 of the tensor's max.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def K():
    from conditional_score_diffusion_b200 import kernels
    return kernels


def _close(got, ref, rtol, what):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    print(f"[bwd] {what}: max_err={err:.3e} scale={scale:.3e} rel={err / scale:.2e}")
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def _nhwc(x, pitch=None):
    """NCHW fp32 (cpu) -> NHWC bf16 cuda with channel pitch."""
    b, c, h, w = x.shape
    pitch = pitch or (c + 7) // 8 * 8
    out = torch.zeros(b, h, w, pitch, dtype=BF, device="cuda")
    out[..., :c] = x.permute(0, 2, 3, 1).to(BF)
    return out


def _rt(x):
    return x.to(BF).float()


def _wgrad(a, g, taps, stride=1, pad=1, scale=1.0):
    """a NCHW fp32 [B,Cin,IH,IW], g NCHW fp32 [B,Cout,OH,OW] (both bf16-representable) -> dW [Cout,Cin,k,k] on GPU."""
    k = K()
    b, cin, ih, iw = a.shape
    cout = g.shape[1]
    geom = k.pixmajor_geometry(b, ih, iw)
    a_pm = k.pixmajor_alloc(geom, cin, 3 if taps == 9 else 1, "cuda")
    g_pm = k.pixmajor_alloc(geom, cout, 1, "cuda")
    k.nhwc_to_pixmajor(_nhwc(a), 0, cin, geom, a_pm)
    k.nhwc_to_pixmajor(_nhwc(g), 0, cout, geom, g_pm, stride=stride, offset=1 - pad if taps == 9 else 0)
    partial = torch.empty(geom.splits, taps, cout, cin, device="cuda", dtype=torch.float32)
    k.wgrad_gemm(g_pm, cout, a_pm, cin, taps, geom, partial)
    kk = 3 if taps == 9 else 1
    dw = torch.zeros(cout, cin, kk, kk, device="cuda", dtype=torch.float32)
    k.wgrad_reduce(partial, geom.splits, taps, cout, cin, scale, dw, cin * taps, taps, 1)
    return dw


def _wgrad_ref(a, g, ksz, stride, pad):
    w = torch.zeros(g.shape[1], a.shape[1], ksz, ksz, requires_grad=True)
    y = F.conv2d(a, w, stride=stride, padding=pad)
    assert y.shape == g.shape, (y.shape, g.shape)
    y.backward(g)
    return w.grad


@pytest.mark.parametrize("shape", [(2, 16, 24, 8, 8), (3, 96, 96, 16, 24), (5, 40, 72, 5, 5), (2, 288, 300, 10, 10),
                                   (2, 6, 128, 32, 32), (2, 96, 6, 32, 32)])
def test_wgrad_3x3(shape):
    b, cin, cout, h, w = shape
    g_ = torch.Generator().manual_seed(sum(shape))
    a = _rt(torch.randn(b, cin, h, w, generator=g_))
    g = _rt(torch.randn(b, cout, h, w, generator=g_))
    _close(_wgrad(a, g, 9), _wgrad_ref(a, g, 3, 1, 1), 2.0 ** -8, f"wgrad 3x3 {shape}")


def test_wgrad_1x1_and_scale():
    g_ = torch.Generator().manual_seed(5)
    a = _rt(torch.randn(3, 72, 12, 20, generator=g_))
    g = _rt(torch.randn(3, 40, 12, 20, generator=g_))
    _close(_wgrad(a, g, 1, scale=0.5), 0.5 * _wgrad_ref(a, g, 1, 1, 0), 2.0 ** -8, "wgrad 1x1")


def test_wgrad_stride2():
    """DDPM Downsample: F.pad(x, (0,1,0,1)) + 3x3 stride-2 VALID conv (models/layers.py:607-629), and the FIR
    conv_downsample_2d's stride-2 VALID conv on the (h+1) x (w+1) pre-filtered tensor (up_or_down_sampling.py:144-178)."""
    g_ = torch.Generator().manual_seed(6)
    a = _rt(torch.randn(2, 32, 16, 16, generator=g_))
    g = _rt(torch.randn(2, 48, 8, 8, generator=g_))
    ref = _wgrad_ref(F.pad(a, (0, 1, 0, 1)), g, 3, 2, 0)
    _close(_wgrad(a, g, 9, stride=2, pad=0), ref, 2.0 ** -8, "wgrad stride2 ddpm")
    a2 = _rt(torch.randn(2, 32, 17, 17, generator=g_))
    _close(_wgrad(a2, g, 9, stride=2, pad=0), _wgrad_ref(a2, g, 3, 2, 0), 2.0 ** -8, "wgrad stride2 fir")


def test_wgrad_full_size_linearity():
    """BASELINE size (64 x 96 x 160 x 160): wgrad(a, g1 + g2) == wgrad(a, g1) + wgrad(a, g2), and a known answer:
    with a = 1 everywhere and g = 1 everywhere the centre tap is B*H*W and a corner tap B*(H-1)*(W-1)."""
    b, c, h, w = 8, 96, 160, 160
    a = torch.ones(b, c, h, w)
    g = torch.ones(b, c, h, w)
    dw = _wgrad(a, g, 9).cpu()
    assert torch.all(dw[:, :, 1, 1] == b * h * w)
    assert torch.all(dw[:, :, 0, 0] == b * (h - 1) * (w - 1))
    assert torch.all(dw[:, :, 1, 2] == b * h * (w - 1))


def test_dgrad_via_conv_gemm():
    """Data gradient of a 3x3 conv = conv with flipped, transposed weights; stride 2 through zero stuffing."""
    k = K()
    g_ = torch.Generator().manual_seed(7)
    b, cin, cout, h, w = 2, 40, 64, 16, 16
    wt = _rt(torch.randn(cout, cin, 3, 3, generator=g_) * 0.1)
    g = _rt(torch.randn(b, cout, h, w, generator=g_))
    x = torch.zeros(b, cin, h, w, requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(g)
    wd = wt.flip(2, 3).transpose(0, 1).contiguous()       # [cin, cout, 3, 3]
    packed = k.pack_conv_weight(wd.cuda())
    out = torch.empty(b, h, w, cin, device="cuda", dtype=BF)
    gn = _nhwc(g)
    k.conv_gemm([(gn, gn.shape[-1], 0, cout, 9)], packed, cin, out, batch=b, h=h, w=w)
    _close(out.permute(0, 3, 1, 2), x.grad, 2.0 ** -8, "dgrad stride 1")
    # stride 2, pad 0 on an (2h+1) input (FIR conv_downsample) and on a 2h input padded after (DDPM)
    for ih in (2 * h + 1, 2 * h):
        x = torch.zeros(b, cin, ih, ih, requires_grad=True)
        xin = x if ih % 2 else F.pad(x, (0, 1, 0, 1))
        F.conv2d(xin, wt, stride=2).backward(g)
        gz = torch.empty(b, ih, ih, gn.shape[-1], device="cuda", dtype=BF)
        k.zero_stuff(gn, gz, 2, 1)
        out = torch.empty(b, ih, ih, cin, device="cuda", dtype=BF)
        k.conv_gemm([(gz, gz.shape[-1], 0, cout, 9)], packed, cin, out, batch=b, h=ih, w=ih)
        _close(out.permute(0, 3, 1, 2), x.grad, 2.0 ** -8, f"dgrad stride 2 ih={ih}")


@pytest.mark.parametrize("cfg", [(2, 32, 0, 12, 12, True), (3, 96, 96, 8, 8, True), (2, 192, 96, 5, 5, True),
                                 (2, 64, 0, 10, 10, False)])
def test_groupnorm_silu_backward(cfg):
    b, c0, c1, h, w, silu = cfg
    k = K()
    g_ = torch.Generator().manual_seed(sum(cfg[:5]))
    C = c0 + c1
    groups = min(C // 4, 32)
    x0 = _rt(torch.randn(b, c0, h, w, generator=g_) * 1.5 + 0.3)
    x1 = _rt(torch.randn(b, c1, h, w, generator=g_)) if c1 else None
    gamma = torch.rand(C, generator=g_) + 0.5
    beta = torch.randn(C, generator=g_) * 0.2
    dy = _rt(torch.randn(b, C, h, w, generator=g_))
    # reference
    xs = [x0.clone().requires_grad_(True)] + ([x1.clone().requires_grad_(True)] if c1 else [])
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(torch.cat(xs, 1), groups, gm, bt, eps=1e-6)
    if silu:
        y = F.silu(y)
    y.backward(dy)
    # ours
    n0 = _nhwc(x0)
    n1 = _nhwc(x1) if c1 else None
    dyn = _nhwc(dy)
    gm_d, bt_d = gamma.cuda(), beta.cuda()
    sums0 = torch.zeros(b, c0, 2, device="cuda")
    k.gn_chan_stats(n0, c0, sums0)
    sums1 = None
    if c1:
        sums1 = torch.zeros(b, c1, 2, device="cuda")
        k.gn_chan_stats(n1, c1, sums1)
    coef0 = torch.empty(b, c0, 2, device="cuda")
    coef1 = torch.empty(b, c1, 2, device="cuda") if c1 else None
    k.gn_coeffs(sums0, c0, sums1, c1, gm_d, bt_d, coef0, coef1, h * w, groups)
    s = torch.zeros(b, C, 2, device="cuda")
    k.gn_bwd_stats(n0, c0, dyn, 0, coef0, s, 0, silu)
    if c1:
        k.gn_bwd_stats(n1, c1, dyn, c0, coef1, s, c0, silu)
    bcoef = torch.empty(b, C, 4, device="cuda")
    dgamma = torch.zeros(C, device="cuda")
    dbeta = torch.zeros(C, device="cuda")
    k.gn_bwd_coeffs(sums0, c0, sums1, c1, gm_d, s, bcoef, dgamma, dbeta, h * w, groups)
    dx0 = torch.empty_like(n0)
    k.gn_bwd_apply(n0, c0, dyn, 0, coef0, bcoef, 0, dx0, silu, False)
    _close(dx0[..., :c0].permute(0, 3, 1, 2), xs[0].grad, 2.0 ** -7, f"gn bwd dx0 {cfg}")
    if c1:
        dx1 = torch.ones_like(n1)
        k.gn_bwd_apply(n1, c1, dyn, c0, coef1, bcoef, c0, dx1, silu, True)   # accumulate onto ones
        _close(dx1[..., :c1].permute(0, 3, 1, 2) - 1.0, xs[1].grad, 2.0 ** -6, f"gn bwd dx1 (accumulated) {cfg}")
    _close(dgamma, gm.grad, 1e-3, f"gn bwd dgamma {cfg}")
    _close(dbeta, bt.grad, 1e-3, f"gn bwd dbeta {cfg}")


@pytest.mark.parametrize("mode", ["up", "down", "prefilter"])
def test_fir_backward_is_adjoint(mode):
    """<fir(x), g> == <x, fir_bwd(g)> against the forward kernel, and equality with autograd of upfirdn2d_native's
    conv formulation for the [1,3,3,1] taps."""
    k = K()
    g_ = torch.Generator().manual_seed(11)
    b, c, h, w = 2, 16, 12, 20
    taps = (1.0, 3.0, 3.0, 1.0)
    x = _rt(torch.randn(b, c, h, w, generator=g_))
    oh, ow = {"up": (2 * h, 2 * w), "down": (h // 2, w // 2), "prefilter": (h + 1, w + 1)}[mode]
    g = _rt(torch.randn(b, c, oh, ow, generator=g_))
    xn, gn = _nhwc(x), _nhwc(g)
    y = torch.empty(b, oh, ow, c, device="cuda", dtype=BF)
    k.fir_resample(xn, y, mode, taps)
    din = torch.empty_like(xn)
    k.fir_resample_bwd(gn, din, mode, taps)
    # reference: autograd through the dense formulation
    k1 = torch.tensor(taps)
    k2 = torch.outer(k1, k1)
    k2 = k2 / k2.sum() * (4.0 if mode == "up" else 1.0)
    xr = x.clone().requires_grad_(True)
    if mode == "up":
        z = torch.zeros(b, c, 2 * h, 2 * w)
        z[:, :, ::2, ::2] = 1
        up = torch.zeros(b, c, 2 * h, 2 * w) + 0
        up = F.interpolate(xr, scale_factor=2, mode="nearest") * z
        yr = F.conv2d(F.pad(up, (2, 1, 2, 1)).reshape(b * c, 1, 2 * h + 3, 2 * w + 3), k2.flip(0, 1)[None, None])
    elif mode == "down":
        yr = F.conv2d(F.pad(xr, (1, 1, 1, 1)).reshape(b * c, 1, h + 2, w + 2), k2.flip(0, 1)[None, None], stride=2)
    else:
        yr = F.conv2d(F.pad(xr, (2, 2, 2, 2)).reshape(b * c, 1, h + 4, w + 4), k2.flip(0, 1)[None, None])
    yr = yr.reshape(b, c, oh, ow)
    _close(y.permute(0, 3, 1, 2), yr, 2.0 ** -8, f"fir fwd {mode} (reference formulation check)")
    yr.backward(g)
    _close(din.permute(0, 3, 1, 2), xr.grad, 2.0 ** -8, f"fir bwd {mode}")
    # accumulate flag
    din2 = torch.ones_like(xn)
    k.fir_resample_bwd(gn, din2, mode, taps, accumulate=True)
    _close(din2.permute(0, 3, 1, 2) - 1.0, xr.grad, 2.0 ** -6, f"fir bwd {mode} accumulate")


def test_softmax_backward_and_transpose():
    k = K()
    g_ = torch.Generator().manual_seed(12)
    z, L, c = 3, 100, 64
    lp = (L + 7) // 8 * 8
    scale = c ** -0.5
    logits = torch.randn(z, L, L, generator=g_) * 3
    dp = torch.randn(z, L, L, generator=g_)
    lg = logits.clone().requires_grad_(True)
    p_ref = F.softmax(lg * scale, dim=-1)
    s_d = torch.zeros(z, L, lp, device="cuda")
    s_d[..., :L] = logits.cuda()
    p_d = torch.empty(z, L, lp, device="cuda", dtype=BF)
    k.softmax_rows(s_d, p_d, L, scale)
    # backward reference with the bf16-rounded probabilities
    pr = p_d[..., :L].float().cpu()
    ds_ref = scale * pr * (dp - (pr * dp).sum(-1, keepdim=True))
    dp_d = torch.zeros(z, L, lp, device="cuda")
    dp_d[..., :L] = dp.cuda()
    ds = torch.empty(z, L, lp, device="cuda", dtype=BF)
    k.softmax_bwd(p_d, dp_d, ds, L, scale)
    _close(ds[..., :L], ds_ref, 2.0 ** -8, "softmax bwd")
    assert torch.all(ds[..., L:] == 0)
    p_ref.backward(dp)
    _close(ds[..., :L], lg.grad, 2.0 ** -6, "softmax bwd vs autograd")
    # transpose
    x = torch.randn(z, L, c, generator=g_).to(BF).cuda()
    xt = torch.full((z, c, lp), 7.0, device="cuda", dtype=BF)
    k.transpose(x, xt, L, c)
    assert torch.equal(xt[..., :L], x.transpose(1, 2))
    assert torch.all(xt[..., L:] == 0)


def test_elementwise_backward_helpers():
    k = K()
    g_ = torch.Generator().manual_seed(13)
    # axpy
    a = torch.randn(2, 6, 6, 16, generator=g_).to(BF).cuda()
    d = torch.randn(2, 6, 6, 16, generator=g_).to(BF).cuda()
    d0 = d.clone()
    k.axpy(a, d, 0.5, True)
    _close(d, d0.float() + 0.5 * a.float(), 2.0 ** -8, "axpy accumulate")
    k.axpy(a, d, 2.0, False)
    _close(d, 2.0 * a.float(), 2.0 ** -8, "axpy write")
    # gradient layout
    g0 = torch.randn(3, 3, 8, 8, generator=g_)
    g1 = torch.randn(3, 3, 8, 8, generator=g_)
    rs0, rs1 = torch.rand(3, generator=g_) + 0.5, torch.rand(3, generator=g_) + 0.5
    out = torch.empty(3, 8, 8, 8, device="cuda", dtype=BF)
    k.nchw_grad_to_nhwc(g0.cuda(), 3, rs0.cuda(), g1.cuda(), 3, rs1.cuda(), out)
    ref = torch.cat([g0 * rs0[:, None, None, None], g1 * rs1[:, None, None, None]], 1)
    _close(out[..., :6].permute(0, 3, 1, 2), ref, 2.0 ** -8, "nchw grad -> nhwc")
    assert torch.all(out[..., 6:] == 0)
    k.nchw_grad_to_nhwc(g0.cuda(), 3, None, None, 3, None, out)
    _close(out[..., :6].permute(0, 3, 1, 2), torch.cat([g0, torch.zeros_like(g1)], 1), 2.0 ** -8, "nchw grad, one group")
    # bias / temb projection gradient
    gy = torch.randn(4, 24, 6, 6, generator=g_).to(BF)
    sums = torch.zeros(4, 24, 2, device="cuda")
    k.gn_chan_stats(_nhwc(gy.float()), 24, sums)
    db = torch.zeros(24, device="cuda")
    dt = torch.zeros(4, 40, device="cuda")
    k.bias_temb_grad(sums, 24, 0.5, db, None, dt[:, 8:], 40)
    _close(db, 0.5 * gy.float().sum((0, 2, 3)), 1e-5, "bias grad")
    _close(dt[:, 8:32], 0.5 * gy.float().sum((2, 3)), 1e-5, "temb projection grad")
    # small GEMM, all transposition modes, with beta and bias
    m, n, kk = 7, 13, 29
    for ta in (0, 1):
        for tb in (0, 1):
            A = torch.randn((kk, m) if ta else (m, kk), generator=g_)
            B = torch.randn((n, kk) if tb else (kk, n), generator=g_)
            C0 = torch.randn(m, n, generator=g_)
            bias = torch.randn(n, generator=g_)
            C = C0.clone().cuda()
            k.sgemm_small(ta, tb, m, n, kk, A.cuda(), A.shape[1], B.cuda(), B.shape[1], C, n, alpha=0.7, beta=0.3,
                          bias=bias.cuda())
            ref = 0.7 * (A.t() if ta else A) @ (B.t() if tb else B) + 0.3 * C0 + bias
            _close(C, ref, 1e-5, f"sgemm ta={ta} tb={tb}")
    # silu forward / backward
    x = torch.randn(1000, generator=g_) * 3
    dy = torch.randn(1000, generator=g_)
    xr = x.clone().requires_grad_(True)
    F.silu(xr).backward(dy)
    y = torch.empty(1000, device="cuda")
    _close(k.silu_f32(x.cuda(), y), F.silu(x), 1e-5, "silu fwd")
    _close(k.silu_f32(x.cuda(), y, dy.cuda()), xr.grad, 1e-5, "silu bwd")
    # dsm loss backward
    b, n_ = 3, 3 * 8 * 8
    score = torch.randn(b, n_, generator=g_)
    z = torch.randn(b, n_, generator=g_)
    a_, c_, w_ = torch.rand(b, generator=g_) + 0.5, torch.rand(b, generator=g_) + 0.5, torch.rand(b, generator=g_)
    gl = torch.randn(b, generator=g_)
    sr = score.clone().requires_grad_(True)
    (w_[:, None] * (a_[:, None] * sr + c_[:, None] * z) ** 2).sum(1).backward(gl)
    ds = torch.empty(b, n_, device="cuda")
    k.dsm_loss_bwd(score.cuda(), z.cuda(), a_.cuda(), c_.cuda(), w_.cuda(), gl.cuda(), ds)
    _close(ds, sr.grad, 1e-5, "dsm loss bwd")


def test_time_features_match_fused_embedding():
    """csd_time_features_f32 + sgemm + silu == csd_time_embedding_f32 (the fused forward kernel)."""
    k = K()
    g_ = torch.Generator().manual_seed(14)
    for et, nf in (("fourier", 16), ("positional", 32)):
        b = 5
        embed = 2 * nf if et == "fourier" else nf
        labels = (torch.rand(b, generator=g_) * (3 if et == "fourier" else 999)).cuda()
        fw = (torch.randn(nf, generator=g_) * 16).cuda() if et == "fourier" else None
        w0, b0 = (torch.randn(4 * nf, embed, generator=g_) * 0.1).cuda(), torch.randn(4 * nf, generator=g_).cuda()
        w1, b1 = (torch.randn(4 * nf, 4 * nf, generator=g_) * 0.1).cuda(), torch.randn(4 * nf, generator=g_).cuda()
        fused = torch.empty(b, 4 * nf, device="cuda")
        k.time_embedding(labels, nf, et, fw, w0, b0, w1, b1, fused)
        emb = torch.empty(b, embed, device="cuda")
        k.time_features(labels, nf, et, fw, emb)
        h0 = torch.empty(b, 4 * nf, device="cuda")
        k.sgemm_small(0, 1, b, 4 * nf, embed, emb, embed, w0, embed, h0, 4 * nf, bias=b0)
        k.silu_f32(h0, h0)
        t = torch.empty(b, 4 * nf, device="cuda")
        k.sgemm_small(0, 1, b, 4 * nf, 4 * nf, h0, 4 * nf, w1, 4 * nf, t, 4 * nf, bias=b1)
        k.silu_f32(t, t)
        _close(t, fused, 1e-5, f"time embedding decomposition {et}")


def _wgrad_direct(a, g, taps, scale=1.0, a_pitch=None, g_pitch=None):
    """Direct (MN-major) path: a [B,Cin,H,W], g [B,Cout,H,W] on the same grid."""
    k = K()
    b, cin, h, w = a.shape
    cout = g.shape[1]
    an, gn = _nhwc(a, a_pitch), _nhwc(g, g_pitch)
    splits = k.wgrad_direct_splits(b, h, w, cout, cin, taps)
    partial = torch.full((splits, taps, cout, cin), float("nan"), device="cuda")
    k.wgrad_direct(gn, 0, cout, an, 0, cin, taps, partial, splits)
    kk = 3 if taps == 9 else 1
    dw = torch.zeros(cout, cin, kk, kk, device="cuda")
    k.wgrad_reduce(partial, splits, taps, cout, cin, scale, dw, cin * taps, taps, 1)
    return dw


@pytest.mark.parametrize("shape", [(2, 64, 64, 16, 16), (3, 96, 96, 16, 24), (2, 128, 128, 32, 32), (2, 192, 96, 40, 40),
                                   (2, 256, 288, 10, 12), (1, 72, 136, 9, 21), (2, 320, 64, 16, 16)])
def test_wgrad_direct_3x3(shape):
    b, cin, cout, h, w = shape
    g_ = torch.Generator().manual_seed(sum(shape) + 1)
    a = _rt(torch.randn(b, cin, h, w, generator=g_))
    g = _rt(torch.randn(b, cout, h, w, generator=g_))
    _close(_wgrad_direct(a, g, 9), _wgrad_ref(a, g, 3, 1, 1), 2.0 ** -8, f"wgrad direct 3x3 {shape}")


def test_wgrad_direct_1x1_pitch_and_known_answer():
    g_ = torch.Generator().manual_seed(9)
    a = _rt(torch.randn(3, 72, 12, 20, generator=g_))
    g = _rt(torch.randn(3, 104, 12, 20, generator=g_))
    _close(_wgrad_direct(a, g, 1, scale=0.5, a_pitch=80, g_pitch=112), 0.5 * _wgrad_ref(a, g, 1, 1, 0), 2.0 ** -8,
           "wgrad direct 1x1")
    b, c, h, w = 4, 96, 160, 160
    dw = _wgrad_direct(torch.ones(b, c, h, w), torch.ones(b, c, h, w), 9).cpu()
    assert torch.all(dw[:, :, 1, 1] == b * h * w)
    assert torch.all(dw[:, :, 0, 0] == b * (h - 1) * (w - 1))
    assert torch.all(dw[:, :, 2, 1] == b * (h - 1) * w)


def test_wgrad_direct_cluster_multicast_variant():
    """The opt-in 3-CTA cluster / TMA-multicast variant (CSD_WGRAD_CLUSTER=1) computes the same weight gradient."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, sys; sys.path.insert(0, 'tests'); import test_gpu_backward_ops as t;"
        "g_ = torch.Generator().manual_seed(3);"
        "a = t._rt(torch.randn(2, 128, 32, 32, generator=g_)); g = t._rt(torch.randn(2, 128, 32, 32, generator=g_));"
        "t._close(t._wgrad_direct(a, g, 9), t._wgrad_ref(a, g, 3, 1, 1), 2.0 ** -8, 'cluster variant'); print('CLUSTER_OK')")
    env = dict(os.environ, CSD_WGRAD_CLUSTER="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert "CLUSTER_OK" in r.stdout, r.stdout + r.stderr


def test_sgemm_long_k_variant():
    k = K()
    g_ = torch.Generator().manual_seed(15)
    m, n, kk = 50, 96, 3000
    A = torch.randn(m, kk, generator=g_)
    B = torch.randn(kk, n, generator=g_)
    C = torch.zeros(m, n, device="cuda")
    k.sgemm_small(0, 0, m, n, kk, A.cuda(), kk, B.cuda(), n, C, n)
    _close(C, A @ B, 1e-5, "sgemm long-K")
    Bt = torch.randn(n, kk, generator=g_)
    bias = torch.randn(n, generator=g_)
    k.sgemm_small(0, 1, m, n, kk, A.cuda(), kk, Bt.cuda(), kk, C, n, alpha=0.5, beta=1.0, bias=bias.cuda())
    _close(C, A @ B + 0.5 * A @ Bt.t() + bias, 1e-5, "sgemm long-K transposed, beta, bias")
