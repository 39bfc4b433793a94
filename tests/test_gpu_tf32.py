"""The reference-precision ("tf32") plan, kernel by kernel: fp32 NHWC activations in HBM, tcgen05 kind::tf32 operands.

Convolutions / GEMMs (csd_conv_gemm dtype = 1) against torch fp32 convolutions with TF32 switched OFF: the tensor core
reads the upper 19 bits of every fp32 operand (10 explicit mantissa bits), so each product carries a relative error
<= 2^-10 and a K-term dot product of random-sign terms lands at ~2^-11 * sqrt(K) / sqrt(K) of the output scale:
tolerance 1e-3 of the output maximum (the same class as cuDNN's TF32 convolutions the reference runs on under
PyTorch's defaults). The fp32-activation elementwise kernels (GroupNorm, FIR, layout, softmax) do fp32 arithmetic and
are held to 1e-5.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ops as o_ops

pytestmark = pytest.mark.gpu

TF32_RTOL = 1e-3          # single kernels
TF32_NET_RTOL = 2e-3      # whole networks (see tests/test_gpu_real_shapes.py for where the number comes from)
F32_RTOL = 1e-5


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _k():
    from conditional_score_diffusion_b200 import kernels
    return kernels


def _nhwc(x, pitch=None):
    b, c, h, w = x.shape
    pitch = pitch or (c + 7) // 8 * 8
    out = torch.zeros(b, h, w, pitch, device=x.device, dtype=x.dtype)
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out


def _rel(got, ref):
    return ((got.float() - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def _conv(name, B, H, W, cin, cout, taps=9, temb=False, res=False, scale=1.0, stride=1, seed=0):
    k = _k()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    ks = 3 if taps == 9 else 1
    x = torch.randn(B, cin, H, W, device=dev, generator=g)
    wgt = torch.randn(cout, cin, ks, ks, device=dev, generator=g) / math.sqrt(cin * ks * ks)
    a = _nhwc(x)
    n_store = k.ceil_to(cout, 8)
    wt = k.pack_conv_weight(wgt, n_pad=k.ceil_to(cout, 16) if cout <= 256 else 2 * k.ceil_to((cout + 1) // 2, 16),
                            dtype=torch.float32)
    npad = wt.shape[0]
    if stride == 1:
        ref = F.conv2d(x, wgt, padding=ks // 2)
        oh, ow, pad = H, W, 1
    else:   # DDPM Downsample: pad (0,1,0,1) + stride-2 VALID conv
        ref = F.conv2d(F.pad(x, (0, 1, 0, 1)), wgt, stride=2)
        oh, ow, pad = H // 2, W // 2, 0
    bias_t = torch.zeros(npad + 16, device=dev)
    bias_t[:cout] = torch.randn(cout, device=dev, generator=g)
    ref = ref + bias_t[:cout].view(1, -1, 1, 1)
    temb_t = res_t = None
    if temb:
        temb_t = torch.zeros(B, npad + 16, device=dev)
        temb_t[:, :cout] = torch.randn(B, cout, device=dev, generator=g)
        ref = ref + temb_t[:, :cout].reshape(B, cout, 1, 1)
    if res:
        r = torch.randn(B, cout, oh, ow, device=dev, generator=g)
        res_t = _nhwc(r, n_store)
        ref = ref + r
    ref = ref * scale
    out = torch.full((B, oh, ow, n_store), float("nan"), device=dev)
    k.conv_gemm([(a, a.shape[-1], 0, cin, taps)], wt, cout, out, batch=B, h=oh, w=ow, n_store=n_store, bias=bias_t,
                temb=temb_t, temb_pitch=npad + 16, res=res_t, res_pitch=n_store, scale=scale, stride=stride, pad=pad,
                in_h=H, in_w=W, transposed=False)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all(), f"{name}: non-finite output"
    if n_store > cout:
        assert (out[..., cout:] == 0).all()
    rel = _rel(out[..., :cout].permute(0, 3, 1, 2), ref)
    print(f"[tf32 conv] {name}: rel={rel:.3e}")
    assert rel < TF32_RTOL, f"{name}: rel {rel:.3e}"


def test_tf32_conv_shapes():
    _conv("1x1 64->96 16x16", 2, 16, 16, 64, 96, taps=1)
    _conv("3x3 96->96 16x16", 2, 16, 16, 96, 96)
    _conv("3x3 96->192 20x20 +temb+res*0.707", 3, 20, 20, 96, 192, temb=True, res=True, scale=1 / math.sqrt(2))
    _conv("3x3 6->96 32x32 (pitch 8)", 2, 32, 32, 6, 96)
    _conv("3x3 96->6 32x32", 2, 32, 32, 96, 6)
    _conv("3x3 192->288 10x10 (two accumulators)", 5, 10, 10, 192, 288)
    _conv("3x3 288->288 5x5 B=7 (multi-image tiles)", 7, 5, 5, 288, 288, temb=True)
    _conv("3x3 96->96 13x11 B=3 (ragged)", 3, 13, 11, 96, 96)
    _conv("3x3 stride 2 128->128 16x16", 2, 16, 16, 128, 128, stride=2)
    _conv("3x3 576->288 10x10 (deep K)", 2, 10, 10, 576, 288)
    _conv("3x3 96->96 160x160", 1, 160, 160, 96, 96, temb=True)


def test_tf32_two_segments_and_skip():
    """conv3x3 over cat([a1, a2]) plus a 1x1 skip conv over the raw input as extra K segments (fp32 operands)."""
    k = _k()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    B, H, W, c1, c2, cs, cout = 2, 20, 20, 96, 64, 160, 96
    a1 = torch.randn(B, c1, H, W, device=dev, generator=g)
    a2 = torch.randn(B, c2, H, W, device=dev, generator=g)
    xs = torch.randn(B, cs, H, W, device=dev, generator=g)
    w3 = torch.randn(cout, c1 + c2, 3, 3, device=dev, generator=g) / math.sqrt(9 * (c1 + c2))
    w1 = torch.randn(cout, cs, 1, 1, device=dev, generator=g) / math.sqrt(cs)
    ref = F.conv2d(torch.cat([a1, a2], 1), w3, padding=1) + F.conv2d(xs, w1)
    pk = lambda w: k.pack_conv_weight(w, dtype=torch.float32)
    wt = torch.cat([pk(w3[:, :c1]), pk(w3[:, c1:]), pk(w1)], dim=1).contiguous()
    out = torch.empty(B, H, W, cout, device=dev)
    k.conv_gemm([(_nhwc(a1), c1, 0, c1, 9), (_nhwc(a2), c2, 0, c2, 9), (_nhwc(xs), cs, 0, cs, 1)], wt, cout, out,
                batch=B, h=H, w=W, transposed=False)
    rel = _rel(out.permute(0, 3, 1, 2), ref)
    print(f"[tf32 conv] two segments + skip: rel={rel:.3e}")
    assert rel < TF32_RTOL


def test_tf32_attention_block_vs_oracle():
    """AttnBlockpp through the engine's tf32 plan pieces (z-batched GEMMs with fp32 logits and probabilities)."""
    from conditional_score_diffusion_b200 import engine as E
    from conditional_score_diffusion_b200.models import layerspp
    from oracle import ncsnpp as o_net
    torch.manual_seed(3)
    for (c, hw, b) in [(192, 20, 3), (288, 5, 4), (256, 16, 2)]:
        blk = layerspp.AttnBlockpp(c, skip_rescale=True, init_scale=1.0).cuda()
        with torch.no_grad():
            for p in blk.parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn_like(p))
        x = torch.randn(b, c, hw, hw, device="cuda")
        sd = {"all_modules.0." + k_: v.detach().cpu() for k_, v in blk.state_dict().items()}
        from types import SimpleNamespace
        ref = o_net.attn_block(sd, 0, x.cpu(), SimpleNamespace(skip_rescale=True))

        class _Net(torch.nn.Module):
            pass
        net = _Net()
        net.all_modules = torch.nn.ModuleList([blk])
        eng = E.NetEngine(net)
        eng.precision = "tf32"
        eng.device = torch.device("cuda")
        pk = eng._pack_attn(blk, eng.device)
        rec = E.Recorder()
        pool = E.BufferPool(eng.device, torch.float32)
        ops = E.BlockOps(eng.device, pool, rec, torch.zeros(1 << 20, device="cuda"))
        a = E.Act(_nhwc(x), c)
        out = ops.attention(pk, a, True)
        rec.run()
        torch.cuda.synchronize()
        rel = _rel(out.t[..., :c].permute(0, 3, 1, 2).cpu(), ref)
        print(f"[tf32 attn] C={c} L={hw * hw}: rel={rel:.3e}")
        assert rel < TF32_RTOL


@pytest.mark.parametrize("shape", [(3, 96, 20, 20), (2, 192, 40, 40), (2, 288, 5, 5), (1, 96, 160, 160)])
def test_f32_groupnorm_kernels_vs_torch(shape):
    k = _k()
    b, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(c + h)
    x = torch.randn(*shape, device="cuda", generator=g) * 3 + 0.5
    gamma = torch.randn(c, device="cuda", generator=g)
    beta = torch.randn(c, device="cuda", generator=g)
    groups = min(c // 4, 32)
    ref = F.silu(F.group_norm(x, groups, gamma, beta, eps=1e-6))
    a = _nhwc(x)
    sums = torch.full((b, c, 2), float("nan"), device="cuda")
    k.gn_chan_stats(a, c, sums)
    ref_s = torch.stack([x.sum((2, 3)), (x * x).sum((2, 3))], -1)
    assert _rel(sums, ref_s) < F32_RTOL
    sums2 = torch.empty_like(sums)
    k.gn_chan_stats(a, c, sums2)
    assert torch.equal(sums, sums2), "gn_chan_stats must be bitwise reproducible"
    out = torch.empty_like(a)
    k.gn_apply(a, c, sums, None, 0, None, gamma, beta, out, groups, 1e-6, True)
    assert _rel(out.permute(0, 3, 1, 2), ref) < 2e-5
    if k.gn_fused_supported(c, 0, h * w, groups, b, torch.float32):
        out2 = torch.empty_like(a)
        k.gn_fused(a, c, None, 0, gamma, beta, out2, groups, 1e-6, True)
        assert _rel(out2.permute(0, 3, 1, 2), ref) < 2e-5
    # bf16 statistics are deterministic too
    ab = a.to(torch.bfloat16)
    s1, s2 = torch.empty_like(sums), torch.empty_like(sums)
    k.gn_chan_stats(ab, c, s1)
    k.gn_chan_stats(ab, c, s2)
    assert torch.equal(s1, s2)
    assert _rel(s1, torch.stack([ab.float().sum((1, 2)), (ab.float() ** 2).sum((1, 2))], -1)) < F32_RTOL


def test_f32_groupnorm_concat_two_sources():
    k = _k()
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(2, 96, 20, 20, device="cuda", generator=g)
    x1 = torch.randn(2, 192, 20, 20, device="cuda", generator=g) * 2
    c = 288
    gamma = torch.randn(c, device="cuda", generator=g)
    beta = torch.randn(c, device="cuda", generator=g)
    ref = F.silu(F.group_norm(torch.cat([x0, x1], 1), 32, gamma, beta, eps=1e-6))
    a0, a1 = _nhwc(x0), _nhwc(x1)
    s0, s1 = torch.empty(2, 96, 2, device="cuda"), torch.empty(2, 192, 2, device="cuda")
    k.gn_chan_stats(a0, 96, s0)
    k.gn_chan_stats(a1, 192, s1)
    out = torch.empty(2, 20, 20, c, device="cuda")
    k.gn_apply(a0, 96, s0, a1, 192, s1, gamma, beta, out, 32, 1e-6, True)
    assert _rel(out.permute(0, 3, 1, 2), ref) < 2e-5
    out2 = torch.empty_like(out)
    assert k.gn_fused_supported(96, 192, 400, 32, 2, torch.float32)
    k.gn_fused(a0, 96, a1, 192, gamma, beta, out2, 32, 1e-6, True)
    assert _rel(out2.permute(0, 3, 1, 2), ref) < 2e-5


@pytest.mark.parametrize("shape", [(2, 96, 80, 80), (1, 192, 20, 20), (2, 8, 10, 10), (1, 288, 5, 5)])
def test_f32_fir_resample_vs_oracle(shape):
    k = _k()
    b, c, h, w = shape
    g = torch.Generator().manual_seed(h)
    x = torch.randn(*shape, generator=g)
    a = _nhwc(x.cuda())
    modes = [("up", o_ops.upsample_2d(x, (1, 3, 3, 1)))]
    if h % 2 == 0:
        modes.append(("down", o_ops.downsample_2d(x, (1, 3, 3, 1))))
    for mode, ref in modes:
        out = torch.empty(b, ref.shape[2], ref.shape[3], a.shape[-1], device="cuda")
        k.fir_resample(a, out, mode, [1, 3, 3, 1])
        assert _rel(out[..., :c].permute(0, 3, 1, 2).cpu(), ref) < F32_RTOL, mode
        # operand form: the same values rounded to tf32 (10 mantissa bits): relative error <= 2^-11 per element
        out_r = torch.empty_like(out)
        k.fir_resample(a, out_r, mode, [1, 3, 3, 1], round_out=True)
        assert ((out_r - out).abs() <= out.abs() * 2.0 ** -11 + 1e-30).all()
        assert (out_r.view(torch.int32) & 0x1FFF).eq(0).all(), "tf32-rounded words have 13 zero low bits"
    add = torch.randn(b, 2 * h, 2 * w, a.shape[-1], device="cuda")
    out = torch.empty_like(add)
    k.fir_resample(a, out, "up", [1, 3, 3, 1], add=add)
    ref = o_ops.upsample_2d(x, (1, 3, 3, 1)) + add[..., :c].permute(0, 3, 1, 2).cpu()
    assert _rel(out[..., :c].permute(0, 3, 1, 2).cpu(), ref) < F32_RTOL


def test_f32_layout_and_softmax():
    k = _k()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(2, 3, 12, 10, device="cuda", generator=g)
    y = torch.randn(2, 3, 12, 10, device="cuda", generator=g)
    out = torch.full((2, 12, 10, 8), float("nan"), device="cuda")
    k.nchw_to_nhwc(x, y, out, 2.0, -1.0)
    ref = torch.cat([x, y], 1) * 2 - 1
    got = out[..., :6].permute(0, 3, 1, 2)
    # the fp32 plan's network input is stored rounded to tf32 (it only feeds tensor-core operands)
    assert torch.equal(got, k.round_tf32(ref)) or ((got - ref).abs() <= ref.abs() * 2.0 ** -11 + 1e-7).all()
    assert (out.view(torch.int32) & 0x1FFF).eq(0).all()
    assert (out[..., 6:] == 0).all()
    back = torch.empty(2, 3, 12, 10, device="cuda")
    rs = torch.tensor([0.5, 3.0], device="cuda")
    k.nhwc_to_nchw(out, 3, 3, back, rs)
    assert torch.allclose(back, out[..., 3:6].permute(0, 3, 1, 2) * rs.view(2, 1, 1, 1), atol=1e-6)
    logits = torch.randn(6, 37, 40, device="cuda", generator=g) * 4
    probs = torch.empty(6, 37, 40, device="cuda")
    k.softmax_rows(logits, probs, 37, 0.3)
    ref = torch.softmax(logits[..., :37] * 0.3, -1)
    assert ((probs[..., :37] - ref).abs() <= ref * 2.0 ** -11 + 2e-6).all()     # stored rounded to tf32
    assert (probs[..., 37:] == 0).all()


def test_tf32_network_small_golden():
    """The tiny golden networks through the tf32 plan: 1e-3 against the real reference's outputs."""
    from golden_utils import golden, to_namespace
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    f = golden()["ncsnpp_paired"]
    m = utils.create_model(to_namespace(f["config"]))
    m.load_state_dict(f["state_dict"], strict=True)
    m = m.cuda().eval().set_precision("tf32")
    with torch.no_grad():
        out = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
        out2 = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
    for key, ref in (("x", f["out_x"]), ("y", f["out_y"])):
        rel = _rel(out[key].cpu(), ref)
        print(f"[tf32 net] golden ncsnpp_paired {key}: rel={rel:.3e}")
        assert rel < TF32_NET_RTOL
    assert torch.equal(out["x"], out2["x"]), "the tf32 plan has no atomics: two runs must agree bitwise"
    # switching back re-plans in bf16
    m.set_precision("bf16")
    with torch.no_grad():
        out3 = m({"x": f["x"].cuda(), "y": f["y"].cuda()}, f["labels"].cuda())
    assert _rel(out3["x"].cpu(), f["out_x"]) < 2e-2
    f2 = golden()["ncsnpp_cifar"]
    m2 = utils.create_model(to_namespace(f2["config"]))
    m2.load_state_dict(f2["state_dict"], strict=True)
    m2 = m2.cuda().eval().set_precision("tf32")
    with torch.no_grad():
        o2 = m2(f2["x"].cuda(), f2["labels"].cuda())
    rel = _rel(o2.cpu(), f2["out"])
    print(f"[tf32 net] golden ncsnpp cifar: rel={rel:.3e}")
    assert rel < TF32_NET_RTOL


# ---- the transposed kernel's kind::tf32 instance (fused GroupNorm prologue, epilogue statistics, fp32 residual) ------
def _tp_case(name, B, H, W, cins, cout, seed=0, silu=True, skip_c=0, norm=True, res=False, temb=False, scale=1.0):
    k = _k()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    xs, coefs, normed = [], [], []
    for c in cins:
        x = torch.randn(B, c, H, W, device=dev, generator=g)
        sc = torch.rand(B, c, device=dev, generator=g) + 0.5
        sh = torch.randn(B, c, device=dev, generator=g)
        y = x * sc[:, :, None, None] + sh[:, :, None, None] if norm else x
        if norm and silu:
            y = F.silu(y)
        xs.append(x)
        coefs.append(torch.stack([sc, sh], dim=-1).contiguous() if norm else None)
        normed.append(y)
    cin = sum(cins)
    wgt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)
    ref = F.conv2d(torch.cat(normed, 1), wgt, padding=1)
    parts, off = [], 0
    for c in cins:
        parts.append(k.pack_conv_weight(wgt[:, off:off + c], dtype=torch.float32))
        off += c
    segs = [(_nhwc(x), c, 0, c, 9, cf, silu) for x, c, cf in zip(xs, cins, coefs)]
    if skip_c:
        xr = torch.randn(B, skip_c, H, W, device=dev, generator=g)
        w1 = torch.randn(cout, skip_c, 1, 1, device=dev, generator=g) / math.sqrt(skip_c)
        ref = ref + F.conv2d(xr, w1)
        parts.append(k.pack_conv_weight(w1, dtype=torch.float32))
        segs.append((_nhwc(xr), skip_c, 0, skip_c, 1))
    wt = torch.cat(parts, dim=1).contiguous()
    npad = wt.shape[0]
    bias_t = torch.zeros(npad + 16, device=dev)
    bias_t[:cout] = torch.randn(cout, device=dev, generator=g)
    ref = ref + bias_t[:cout].view(1, -1, 1, 1)
    temb_t = res_t = None
    if temb:
        temb_t = torch.zeros(B, npad + 16, device=dev)
        temb_t[:, :cout] = torch.randn(B, cout, device=dev, generator=g)
        ref = ref + temb_t[:, :cout].reshape(B, cout, 1, 1)
    if res:
        r = torch.randn(B, cout, H, W, device=dev, generator=g) * 3
        res_t = _nhwc(r)
        ref = ref + r
    ref = ref * scale
    out = torch.full((B, H, W, cout), float("nan"), device=dev)
    tiles_img = math.ceil(H / k.transposed_tile_rows(H)) * math.ceil(W / 8) * 2
    partials = torch.full((B * tiles_img, cout, 2), float("nan"), device=dev)
    k.conv_gemm(segs, wt, cout, out, batch=B, h=H, w=W, bias=bias_t, temb=temb_t, temb_pitch=npad + 16, res=res_t,
                res_pitch=cout, scale=scale, transposed=True, stat_partials=partials)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all(), f"{name}: non-finite output"
    rel = _rel(out.permute(0, 3, 1, 2), ref)
    sums = torch.empty(B, cout, 2, device=dev)
    k.gn_finalize_partials(partials, sums, B, tiles_img, cout)
    o = out.permute(0, 3, 1, 2)
    ref_s = torch.stack([o.sum((2, 3)), (o * o).sum((2, 3))], -1)
    rel_s = _rel(sums, ref_s)
    print(f"[tf32 transposed] {name}: rel={rel:.3e} stats rel={rel_s:.3e}")
    assert rel < TF32_RTOL, f"{name}: rel {rel:.3e}"
    assert rel_s < 1e-4, f"{name}: epilogue statistics {rel_s:.3e}"


def test_tf32_transposed_kernel():
    _tp_case("96->96 32x32 plain", 2, 32, 32, [96], 96, norm=False)
    _tp_case("96->96 32x32 GN+SiLU +temb", 2, 32, 32, [96], 96, temb=True)
    _tp_case("96->96 80x80 (28-row tiles) +res*0.707", 2, 80, 80, [96], 96, seed=1, res=True, scale=1 / math.sqrt(2))
    _tp_case("64->96 24x24 padded cin", 2, 24, 24, [64], 96, seed=2)
    _tp_case("(96+96)->96 64x64 cat + skip conv", 2, 64, 64, [96, 96], 96, seed=3, skip_c=192, scale=1 / math.sqrt(2))
    _tp_case("(192+96)->192 32x32 cat (two 128-row blocks)", 2, 32, 32, [192, 96], 192, seed=4, skip_c=288)
    _tp_case("192->192 32x32 affine only +res", 2, 32, 32, [192], 192, seed=5, silu=False, res=True)
    _tp_case("(192+192)->192 40x40 20-row tiles", 2, 40, 40, [192, 192], 192, seed=6, skip_c=384)
    _tp_case("96->96 160x160", 1, 160, 160, [96], 96, seed=7, temb=True, res=True, scale=1 / math.sqrt(2))
    _tp_case("288->288 64x64 (three blocks, 64 + 32 channel tail)", 1, 64, 64, [288], 288, seed=8, res=True)


def test_tf32_network_transposed_vs_per_tap(monkeypatch):
    """The 64 px / nf 32 network of test_gpu_network through the tf32 plan with and without the transposed kernel: both
    within the tf32 tolerance of the oracle, and close to each other."""
    from golden_utils import golden, to_namespace
    from conditional_score_diffusion_b200 import engine
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    from oracle import ncsnpp as o_net
    f = golden()["ncsnpp_paired"]
    cfg = to_namespace(f["config"])
    cfg.data.image_size = cfg.data.effective_image_size = 64
    cfg.model.nf = 32
    cfg.model.attn_resolutions = (16,)
    torch.manual_seed(41)
    m0 = utils.create_model(cfg)
    g = torch.Generator().manual_seed(42)
    with torch.no_grad():
        for pn, p in m0.named_parameters():
            if pn.endswith("bias") or pn.endswith(".b"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    sd = {k_: v.detach().clone() for k_, v in m0.state_dict().items()}
    x = torch.randn(3, 3, 64, 64, generator=g) * 5
    y = torch.rand(3, 3, 64, 64, generator=g)
    labels = torch.rand(3, generator=g) * 999
    ref = o_net.forward_paired(sd, o_net.model_options(cfg), x, y, labels)
    outs = {}
    for tp in (True, False):
        monkeypatch.setattr(engine, "TF32_TRANSPOSED", tp)
        m = utils.create_model(cfg)
        m.load_state_dict(sd)
        m = m.cuda().eval().set_precision("tf32")
        with torch.no_grad():
            outs[tp] = m({"x": x.cuda(), "y": y.cuda()}, labels.cuda())
        names = [getattr(fn, "__name__", "") for fn, _, _ in next(iter(m._engine.plans.values())).rec.ops]
        assert ("gn_coeffs_partials" in names or "gn_coeffs" in names) == tp
        for key in ("x", "y"):
            rel = _rel(outs[tp][key].cpu(), ref[key])
            print(f"[tf32 net 64px] transposed={tp} {key}: rel={rel:.3e}")
            assert rel < TF32_NET_RTOL
