"""Parity of the tcgen05 implicit-GEMM conv / GEMM kernel (csd_conv_gemm) against torch fp32.

Inputs are rounded to bf16 first so that the only differences are fp32 accumulation order and the
bf16 rounding of the stored output (tolerances below say which).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# bf16 output: half-ulp relative 2^-9 on values up to ~max|ref|, plus accumulation-order noise.
BF16_RTOL = 2.0 ** -8
F32_RTOL = 2e-5


def _kern():
    from conditional_score_diffusion_b200 import kernels
    return kernels


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _report(name, got, ref, rtol):
    err = (got.float() - ref).abs()
    scale = ref.abs().max().item() + 1e-6
    rel = err.max().item() / scale
    print(f"[conv_gemm] {name}: max_abs_err={err.max().item():.4e} ref_max={scale:.4e} rel={rel:.3e} "
          f"mean_abs_err={err.mean().item():.3e}")
    if rel > rtol:
        bad = (err > rtol * scale)
        idx = bad.nonzero()[:8].tolist()
        print(f"   first bad indices: {idx}  bad_frac={bad.float().mean().item():.4f}")
    return rel


def _conv_case(name, B, H, W, cin, cout, taps=9, bias=True, temb=False, res=False, scale=1.0,
               c_pitch=None, out_f32=False, seed=0, tile=None, n_tile=None, k_splits=1):
    k = _kern()
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    c_pitch = c_pitch or k.ceil_to(cin, 8)
    x = torch.randn(B, cin, H, W, device=dev, generator=g).to(torch.bfloat16)
    ks = 3 if taps == 9 else 1
    wgt = (torch.randn(cout, cin, ks, ks, device=dev, generator=g) / math.sqrt(cin * ks * ks)).to(torch.bfloat16)
    a = torch.zeros(B, H, W, c_pitch, device=dev, dtype=torch.bfloat16)
    a[..., :cin] = _nhwc(x)
    n_store = k.ceil_to(cout, 8)
    out_pitch = n_store
    wt = k.pack_conv_weight(wgt, n_pad=k.ceil_to(cout, 16) if cout <= 256 else 2 * k.ceil_to((cout + 1) // 2, 16))
    npad = wt.shape[0]
    bias_t = temb_t = res_t = None
    ref = F.conv2d(x.float(), wgt.float(), padding=ks // 2)
    if bias:
        bias_t = torch.zeros(npad + 16, device=dev)
        bias_t[:cout] = torch.randn(cout, device=dev, generator=g)
        ref = ref + bias_t[:cout].view(1, -1, 1, 1)
    if temb:
        temb_t = torch.zeros(B, npad + 16, device=dev)
        temb_t[:, :cout] = torch.randn(B, cout, device=dev, generator=g)
        ref = ref + temb_t[:, :cout].reshape(B, cout, 1, 1)
    if res:
        r = torch.randn(B, cout, H, W, device=dev, generator=g).to(torch.bfloat16)
        res_t = torch.zeros(B, H, W, out_pitch, device=dev, dtype=torch.bfloat16)
        res_t[..., :cout] = _nhwc(r)
        ref = ref + r.float()
    ref = ref * scale
    out = torch.full((B, H, W, out_pitch), float("nan"), device=dev,
                     dtype=torch.float32 if out_f32 else torch.bfloat16)
    k.conv_gemm([(a, c_pitch, 0, cin, taps)], wt, cout, out, batch=B, h=H, w=W, n_store=n_store,
                bias=bias_t, temb=temb_t, temb_pitch=(npad + 16), res=res_t, res_pitch=out_pitch,
                scale=scale, tile=tile, n_tile=n_tile, transposed=False, k_splits=k_splits,
                splitk_ws=(torch.full((k_splits, B * H * W, n_store), float("nan"), device=dev)
                           if k_splits > 1 else None))
    torch.cuda.synchronize()
    got = out[..., :cout].permute(0, 3, 1, 2)
    assert torch.isfinite(out[..., :n_store].float()).all(), f"{name}: non-finite outputs"
    if n_store > cout:
        assert (out[..., cout:n_store].float() == 0).all(), f"{name}: padded output channels not zero"
    rel = _report(name, got, ref, F32_RTOL if out_f32 else BF16_RTOL)
    return rel


def test_gemm_1x1_small():
    assert _conv_case("1x1 64->96 16x16", 2, 16, 16, 64, 96, taps=1) < BF16_RTOL


def test_gemm_1x1_f32_out():
    assert _conv_case("1x1 64->96 16x16 f32", 2, 16, 16, 64, 96, taps=1, out_f32=True) < F32_RTOL


def test_conv3x3_basic():
    assert _conv_case("3x3 96->96 16x16", 2, 16, 16, 96, 96) < BF16_RTOL


def test_conv3x3_epilogue_all():
    assert _conv_case("3x3 96->192 +bias+temb+res*0.707", 3, 20, 20, 96, 192, temb=True, res=True,
                      scale=1 / math.sqrt(2)) < BF16_RTOL


def test_conv3x3_padded_input_channels():
    assert _conv_case("3x3 6->96 (pitch 8)", 2, 32, 32, 6, 96) < BF16_RTOL


def test_conv3x3_small_output_channels():
    assert _conv_case("3x3 96->6", 2, 32, 32, 96, 6) < BF16_RTOL


def test_conv3x3_n288_two_accumulators():
    assert _conv_case("3x3 192->288 10x10 n_tile=288", 5, 10, 10, 192, 288, n_tile=288) < BF16_RTOL
    assert _conv_case("3x3 192->288 10x10 n_tile=144", 5, 10, 10, 192, 288, n_tile=144) < BF16_RTOL


def test_conv3x3_ragged_tiles():
    # 5x5 images, batch not a multiple of the tile's batch extent, deep K
    assert _conv_case("3x3 288->288 5x5 B=7", 7, 5, 5, 288, 288) < BF16_RTOL
    assert _conv_case("3x3 96->96 13x11 B=3", 3, 13, 11, 96, 96) < BF16_RTOL


def test_conv3x3_explicit_tiles():
    for tile in [(16, 8, 1), (8, 16, 1), (8, 8, 2), (4, 4, 8), (32, 4, 1)]:
        assert _conv_case(f"3x3 96->96 32x32 tile={tile}", 8, 32, 32, 96, 96, tile=tile) < BF16_RTOL


@pytest.mark.parametrize("ks", [2, 3, 4, 7])
def test_conv3x3_split_k(ks):
    """Small levels: gridDim.z CTAs share one tile's K range and a second launch adds the fp32 partials in split
    order and applies the whole epilogue (bias, temb, residual, scale). Uneven splits (45 iterations / 7) included."""
    assert _conv_case(f"3x3 288->288 5x5 B=7 split-K {ks}", 7, 5, 5, 288, 288, temb=True, res=True,
                      scale=1 / math.sqrt(2), k_splits=ks) < BF16_RTOL
    assert _conv_case(f"3x3 192->288 10x10 split-K {ks}", 5, 10, 10, 192, 288, n_tile=144, k_splits=ks) < BF16_RTOL
    assert _conv_case(f"3x3 96->6 split-K {ks} f32", 2, 8, 8, 96, 6, out_f32=True, k_splits=ks) < F32_RTOL
    assert _conv_case(f"1x1 288->288 split-K {min(ks, 4)}", 3, 5, 5, 288, 288, taps=1, res=True,
                      k_splits=min(ks, 4)) < BF16_RTOL


def test_split_k_two_segments_and_determinism():
    """Split boundaries falling inside and between segments; run-to-run bitwise identical (fixed summation order)."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(15)
    B, H, W, c1, c2, cs, cout = 3, 10, 10, 96, 64, 160, 96
    a1 = torch.randn(B, c1, H, W, device=dev, generator=g).to(torch.bfloat16)
    a2 = torch.randn(B, c2, H, W, device=dev, generator=g).to(torch.bfloat16)
    xs = torch.randn(B, cs, H, W, device=dev, generator=g).to(torch.bfloat16)
    w3 = (torch.randn(cout, c1 + c2, 3, 3, device=dev, generator=g) / math.sqrt(9 * (c1 + c2))).to(torch.bfloat16)
    w1 = (torch.randn(cout, cs, 1, 1, device=dev, generator=g) / math.sqrt(cs)).to(torch.bfloat16)
    ref = F.conv2d(torch.cat([a1, a2], 1).float(), w3.float(), padding=1) + F.conv2d(xs.float(), w1.float())
    wt = torch.cat([k.pack_conv_weight(w3[:, :c1]), k.pack_conv_weight(w3[:, c1:]), k.pack_conv_weight(w1)], dim=1).contiguous()
    segs = [(_nhwc(a1), c1, 0, c1, 9), (_nhwc(a2), c2, 0, c2, 9), (_nhwc(xs), cs, 0, cs, 1)]
    assert k.pick_k_splits(segs, B, H, W, cout, 96) > 1
    outs = []
    for ks in (1, 2, 3, 5, 5):
        out = torch.empty(B, H, W, cout, device=dev, dtype=torch.bfloat16)
        ws = torch.empty(ks, B * H * W, cout, device=dev) if ks > 1 else None
        k.conv_gemm(segs, wt, cout, out, batch=B, h=H, w=W, k_splits=ks, splitk_ws=ws)
        torch.cuda.synchronize()
        assert _report(f"2 segments + skip, split-K {ks}", out.permute(0, 3, 1, 2), ref, BF16_RTOL) < BF16_RTOL
        outs.append(out)
    assert torch.equal(outs[-1], outs[-2])


def test_two_segments_concat_plus_skip():
    """conv3x3 over cat([a1, a2]) plus a 1x1 skip conv over raw x folded in as extra K."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    B, H, W, c1, c2, cs, cout = 2, 20, 20, 96, 64, 160, 96
    a1 = torch.randn(B, c1, H, W, device=dev, generator=g).to(torch.bfloat16)
    a2 = torch.randn(B, c2, H, W, device=dev, generator=g).to(torch.bfloat16)
    xs = torch.randn(B, cs, H, W, device=dev, generator=g).to(torch.bfloat16)
    w3 = (torch.randn(cout, c1 + c2, 3, 3, device=dev, generator=g) / math.sqrt(9 * (c1 + c2))).to(torch.bfloat16)
    w1 = (torch.randn(cout, cs, 1, 1, device=dev, generator=g) / math.sqrt(cs)).to(torch.bfloat16)
    ref = F.conv2d(torch.cat([a1, a2], 1).float(), w3.float(), padding=1) + F.conv2d(xs.float(), w1.float())
    wt = torch.cat([k.pack_conv_weight(w3[:, :c1]), k.pack_conv_weight(w3[:, c1:]), k.pack_conv_weight(w1)], dim=1).contiguous()
    out = torch.empty(B, H, W, cout, device=dev, dtype=torch.bfloat16)
    k.conv_gemm([(_nhwc(a1), c1, 0, c1, 9), (_nhwc(a2), c2, 0, c2, 9), (_nhwc(xs), cs, 0, cs, 1)], wt, cout, out,
                batch=B, h=H, w=W)
    torch.cuda.synchronize()
    assert _report("2 segments + skip", out.permute(0, 3, 1, 2), ref, BF16_RTOL) < BF16_RTOL


def test_batched_gemm_attention_shapes():
    """S[b] = Q[b] K[b]^T with Q,K slices of one [B, L, 2C] buffer; fp32 logits."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(7)
    for (B, L, C) in [(3, 400, 192), (2, 100, 288), (4, 25, 288)]:
        qk = torch.randn(B, L, 2 * C, device=dev, generator=g).to(torch.bfloat16)
        ref = torch.einsum("blc,bmc->blm", qk[..., :C].float(), qk[..., C:].float())
        lp = k.ceil_to(L, 8)
        out = torch.full((B, L, lp), float("nan"), device=dev, dtype=torch.float32)
        k.conv_gemm([(qk, 2 * C, 0, C, 1)], qk, L, out, batch=1, h=1, w=L, out_pitch=lp, n_store=L,
                    n_tile=208 if L > 256 else k.ceil_to(L, 16), z_batches=B, a_batch_step=1,
                    wt_batch_stride=L * 2 * C, wt_pitch=2 * C, wt_k_off=C, k_valid=C, wt_rows=L,
                    out_z_stride=L * lp)
        torch.cuda.synchronize()
        assert _report(f"QK^T B={B} L={L} C={C}", out[..., :L], ref, F32_RTOL) < F32_RTOL


def test_batched_gemm_weights_as_a_operand():
    """V^T[b] = Wv^T h[b]^T: the weight matrix is the (unbatched) A image, h[b] the batched B."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(9)
    B, L, C = 3, 100, 192
    hh = torch.randn(B, L, C, device=dev, generator=g).to(torch.bfloat16)
    wv = (torch.randn(C, C, device=dev, generator=g) / math.sqrt(C)).to(torch.bfloat16)  # [in, out]
    bv = torch.randn(C + 16, device=dev, generator=g)
    ref = torch.einsum("blc,cd->bdl", hh.float(), wv.float()) + bv[:C].view(1, C, 1)
    lp = k.ceil_to(L, 8)
    out = torch.zeros(B, C, lp, device=dev, dtype=torch.bfloat16)
    a = wv.t().contiguous()  # [out, in] rows = output channel
    k.conv_gemm([(a, C, 0, C, 1)], hh, L, out, batch=1, h=1, w=C, out_pitch=lp, n_store=L,
                n_tile=k.ceil_to(L, 16), z_batches=B, a_batch_step=0, wt_batch_stride=L * C, wt_rows=L,
                out_z_stride=C * lp, bias=bv, bias_per_row=True)
    torch.cuda.synchronize()
    assert _report("V^T", out[..., :L], ref, BF16_RTOL) < BF16_RTOL


def test_conv3x3_full_resolution_timing():
    """160x160, 96->96, B=8: correctness at the dominant shape plus a rough timing print."""
    k = _kern()
    rel = _conv_case("3x3 96->96 160x160 B=8", 8, 160, 160, 96, 96, temb=True, res=True, scale=0.5)
    assert rel < BF16_RTOL
    dev = "cuda"
    B, H, W, C = 64, 160, 160, 96
    a = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    wt = k.pack_conv_weight((torch.randn(C, C, 3, 3, device=dev) / 30).to(torch.bfloat16))
    out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        k.conv_gemm([(a, C, 0, C, 9)], wt, C, out, batch=B, h=H, w=W)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        k.conv_gemm([(a, C, 0, C, 9)], wt, C, out, batch=B, h=H, w=W)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = 2.0 * B * H * W * C * C * 9
    print(f"[conv_gemm] 3x3 96->96 160x160 B=64: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")


def test_conv3x3_stride2_valid():
    """3x3 stride-2 VALID conv over an odd-sized (pre-filtered) input: second half of
    conv_downsample_2d (models/up_or_down_sampling.py:144-178)."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(13)
    for (B, H, cin, cout) in [(3, 17, 8, 32), (2, 33, 96, 192), (5, 11, 64, 64)]:
        x = torch.randn(B, cin, H, H, device=dev, generator=g).to(torch.bfloat16)
        wgt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
        ref = F.conv2d(x.float(), wgt.float(), stride=2, padding=0)
        oh = ref.shape[-1]
        out = torch.empty(B, oh, oh, cout, device=dev, dtype=torch.bfloat16)
        k.conv_gemm([(_nhwc(x), cin, 0, cin, 9)], k.pack_conv_weight(wgt), cout, out, batch=B, h=oh, w=oh,
                    stride=2, pad=0, in_h=H, in_w=H)
        torch.cuda.synchronize()
        assert _report(f"3x3 s2 valid {cin}->{cout} {H}->{oh}", out.permute(0, 3, 1, 2), ref, BF16_RTOL) < BF16_RTOL


def test_conv3x3_stride2_ddpm_downsample():
    """F.pad(x, (0,1,0,1)) + 3x3 stride-2 conv (models/layers.py:607-629): the bottom/right padding is
    TMA's out-of-bounds zero fill."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(14)
    B, H, cin, cout = 2, 16, 64, 64
    x = torch.randn(B, cin, H, H, device=dev, generator=g).to(torch.bfloat16)
    wgt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), wgt.float(), stride=2, padding=0)
    out = torch.empty(B, H // 2, H // 2, cout, device=dev, dtype=torch.bfloat16)
    k.conv_gemm([(_nhwc(x), cin, 0, cin, 9)], k.pack_conv_weight(wgt), cout, out, batch=B, h=H // 2, w=H // 2,
                stride=2, pad=0, in_h=H, in_w=H)
    torch.cuda.synchronize()
    assert _report("ddpm downsample conv", out.permute(0, 3, 1, 2), ref, BF16_RTOL) < BF16_RTOL


def _halo_case(name, B, H, W, cin, cout, mt, seed=0, temb=True, res=True, n_tile=None):
    """Halo-mode kernel against the same fp32 reference (and implicitly against the per-tap kernel)."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, cin, H, W, device=dev, generator=g).to(torch.bfloat16)
    wgt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
    npad = k.ceil_to(cout, 16) if n_tile is None else k.ceil_to(cout, n_tile)
    wt = k.pack_conv_weight(wgt, n_pad=npad)
    bias = torch.zeros(npad + 16, device=dev)
    bias[:cout] = torch.randn(cout, device=dev, generator=g)
    ref = F.conv2d(x.float(), wgt.float(), padding=1) + bias[:cout].view(1, -1, 1, 1)
    temb_t = res_t = None
    if temb:
        temb_t = torch.zeros(B, npad + 16, device=dev)
        temb_t[:, :cout] = torch.randn(B, cout, device=dev, generator=g)
        ref = ref + temb_t[:, :cout].reshape(B, cout, 1, 1)
    if res:
        r = torch.randn(B, cout, H, W, device=dev, generator=g).to(torch.bfloat16)
        res_t = _nhwc(r)
        ref = ref + r.float()
    ref = ref * 0.5
    out = torch.full((B, H, W, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    k.conv_gemm([(_nhwc(x), cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=W, bias=bias, temb=temb_t,
                temb_pitch=npad + 16, res=res_t, res_pitch=cout, scale=0.5, halo=True, mt=mt, n_tile=n_tile)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), f"{name}: non-finite"
    return _report(name, out.permute(0, 3, 1, 2), ref, BF16_RTOL)


def test_halo_mode_basic():
    assert _halo_case("halo 96->96 32x32 mt=1", 2, 32, 32, 96, 96, 1) < BF16_RTOL
    assert _halo_case("halo 96->96 32x32 mt=2", 2, 32, 32, 96, 96, 2) < BF16_RTOL
    assert _halo_case("halo 96->96 64x64 mt=4", 2, 64, 64, 96, 96, 4) < BF16_RTOL


def test_halo_mode_ragged_and_wide():
    assert _halo_case("halo 192->192 40x40 mt=1 (ragged h)", 3, 40, 40, 192, 192, 1) < BF16_RTOL
    assert _halo_case("halo 96->192 80x80 mt=2 (ragged h)", 2, 80, 80, 96, 192, 2, n_tile=None) < BF16_RTOL
    assert _halo_case("halo 64->96 24x20 mt=1 (ragged w)", 2, 24, 20, 64, 96, 1) < BF16_RTOL
    assert _halo_case("halo 192->288 16x16 n_tile=144", 2, 16, 16, 192, 288, 1, n_tile=144) < BF16_RTOL
    assert _halo_case("halo 192->288 16x16 n_tile=288", 2, 16, 16, 192, 288, 1, n_tile=288) < BF16_RTOL
    assert _halo_case("halo 8->96 32x32 (padded cin)", 2, 32, 32, 8, 96, 2) < BF16_RTOL


def test_halo_mode_segments_and_skip():
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(25)
    B, H, W, c1, c2, cs, cout = 2, 32, 32, 96, 64, 160, 96
    a1 = torch.randn(B, c1, H, W, device=dev, generator=g).to(torch.bfloat16)
    a2 = torch.randn(B, c2, H, W, device=dev, generator=g).to(torch.bfloat16)
    xs = torch.randn(B, cs, H, W, device=dev, generator=g).to(torch.bfloat16)
    w3 = (torch.randn(cout, c1 + c2, 3, 3, device=dev, generator=g) / math.sqrt(9 * (c1 + c2))).to(torch.bfloat16)
    w1 = (torch.randn(cout, cs, 1, 1, device=dev, generator=g) / math.sqrt(cs)).to(torch.bfloat16)
    ref = F.conv2d(torch.cat([a1, a2], 1).float(), w3.float(), padding=1) + F.conv2d(xs.float(), w1.float())
    wt = torch.cat([k.pack_conv_weight(w3[:, :c1]), k.pack_conv_weight(w3[:, c1:]), k.pack_conv_weight(w1)], dim=1).contiguous()
    for mt in (1, 2):
        out = torch.empty(B, H, W, cout, device=dev, dtype=torch.bfloat16)
        k.conv_gemm([(_nhwc(a1), c1, 0, c1, 9), (_nhwc(a2), c2, 0, c2, 9), (_nhwc(xs), cs, 0, cs, 1)], wt, cout, out,
                    batch=B, h=H, w=W, halo=True, mt=mt)
        torch.cuda.synchronize()
        assert _report(f"halo 2 segments + skip mt={mt}", out.permute(0, 3, 1, 2), ref, BF16_RTOL) < BF16_RTOL


def test_halo_vs_tap_timing():
    k = _kern()
    dev = "cuda"
    for (B, H, cin, cout) in [(64, 160, 96, 96), (64, 160, 192, 96), (64, 80, 96, 96), (64, 80, 192, 192), (64, 40, 192, 192)]:
        a = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
        wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device=dev) / 30).to(torch.bfloat16))
        out = torch.empty(B, H, H, cout, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * B * H * H * cin * cout * 9
        for label, kw in [("tap", dict(halo=False, transposed=False)), ("halo mt=1", dict(halo=True, mt=1)),
                          ("halo mt=2", dict(halo=True, mt=2)), ("halo mt=4", dict(halo=True, mt=4))]:
            if kw.get("mt", 1) * k.ceil_to(cout, 16) > 512:
                continue
            for _ in range(2):
                k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"[conv timing] {H}x{H} {cin}->{cout} {label:10s}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s")


def _t_case(name, B, H, W, cin, cout, seed=0, temb=True, res=True, stats=True):
    """Transposed halo mode (weights = M operand, 256 pixels = N) against torch fp32."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, cin, H, W, device=dev, generator=g).to(torch.bfloat16)
    wgt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
    npad = k.ceil_to(cout, 16)
    wt = k.pack_conv_weight(wgt, n_pad=npad)
    bias = torch.zeros(npad + 16, device=dev)
    bias[:cout] = torch.randn(cout, device=dev, generator=g)
    ref = F.conv2d(x.float(), wgt.float(), padding=1) + bias[:cout].view(1, -1, 1, 1)
    temb_t = None
    segs = [(_nhwc(x), cin, 0, cin, 9)]
    if temb:
        temb_t = torch.zeros(B, npad + 16, device=dev)
        temb_t[:, :cout] = torch.randn(B, cout, device=dev, generator=g)
        ref = ref + temb_t[:, :cout].reshape(B, cout, 1, 1)
    if res:
        # the transposed kernel takes a residual as one more K segment with identity weights (exact in fp32)
        r = torch.randn(B, cout, H, W, device=dev, generator=g).to(torch.bfloat16)
        eye = torch.eye(cout, device=dev, dtype=torch.bfloat16).view(cout, cout, 1, 1)
        wt = torch.cat([wt, k.pack_conv_weight(eye, n_pad=npad)], dim=1).contiguous()
        segs.append((_nhwc(r), cout, 0, cout, 1))
        ref = ref + r.float()
    ref = ref * 0.5
    n_store = k.ceil_to(cout, 8)
    out = torch.full((B, H, W, n_store), float("nan"), device=dev, dtype=torch.bfloat16)
    tiles = B * math.ceil(H / k.transposed_tile_rows(H)) * math.ceil(W / 8) * 2     # partial sums per (tile, half)
    partials = torch.full((tiles, n_store, 2), float("nan"), device=dev) if stats else None
    k.conv_gemm(segs, wt, cout, out, batch=B, h=H, w=W, n_store=n_store, bias=bias, temb=temb_t,
                temb_pitch=npad + 16, scale=0.5, transposed=True, stat_partials=partials)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), f"{name}: non-finite"
    rel = _report(name, out[..., :cout].permute(0, 3, 1, 2), ref, T_RTOL)
    if stats:
        # the sums are taken from the fp32 values before their bf16 rounding: they differ from the sums of the
        # stored tensor by at most 2^-9 per element (random sign), bounded here by 1e-3 of the absolute sums
        tp = tiles // B
        ps = partials.view(B, tp, n_store, 2).sum(1)
        assert torch.isfinite(ps).all(), f"{name}: unwritten statistics partials"
        o32 = out.float().reshape(B, H * W, n_store)
        e1 = (ps[..., 0] - o32.sum(1)).abs() - 1e-3 * o32.abs().sum(1) - 1e-3
        e2 = (ps[..., 1] - (o32 * o32).sum(1)).abs() - 2e-3 * (o32 * o32).sum(1) - 1e-3
        assert e1.max().item() <= 0, f"{name}: channel sums off by {e1.max().item():.3e}"
        assert e2.max().item() <= 0, f"{name}: channel sums of squares off by {e2.max().item():.3e}"
        r32 = ref.reshape(B, cout, H * W).sum(2)
        assert (ps[..., :cout, 0] - r32).abs().max().item() <= 2e-3 * ref.abs().reshape(B, cout, -1).sum(2).max().item()
    return rel


# One bf16 rounding of the fp32 result (residuals ride on the tensor core as an identity K segment).
T_RTOL = 2.0 ** -8


def test_transposed_halo_mode():
    assert _t_case("T 96->96 32x32", 2, 32, 32, 96, 96) < T_RTOL
    assert _t_case("T 96->96 80x80 (ragged h)", 2, 80, 80, 96, 96) < T_RTOL
    assert _t_case("T 192->192 40x40 (2 channel blocks, 20-row tiles)", 3, 40, 40, 192, 192) < T_RTOL
    assert _t_case("T 96->96 20x24 (one 20-row tile)", 2, 20, 24, 96, 96, seed=7) < T_RTOL
    assert _t_case("T 64->96 24x20 (ragged w)", 2, 24, 20, 64, 96) < T_RTOL
    assert _t_case("T 96->6 32x32 (tiny cout)", 2, 32, 32, 96, 6, temb=False, res=False) < T_RTOL
    assert _t_case("T 8->96 32x32 (padded cin)", 2, 32, 32, 8, 96, res=False) < T_RTOL
    assert _t_case("T 288->288 16x16", 2, 16, 16, 288, 288) < T_RTOL


def test_transposed_segments_and_skip():
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(35)
    B, H, W, c1, c2, cs, cout = 2, 32, 32, 96, 64, 160, 96
    a1 = torch.randn(B, c1, H, W, device=dev, generator=g).to(torch.bfloat16)
    a2 = torch.randn(B, c2, H, W, device=dev, generator=g).to(torch.bfloat16)
    xs = torch.randn(B, cs, H, W, device=dev, generator=g).to(torch.bfloat16)
    w3 = (torch.randn(cout, c1 + c2, 3, 3, device=dev, generator=g) / math.sqrt(9 * (c1 + c2))).to(torch.bfloat16)
    w1 = (torch.randn(cout, cs, 1, 1, device=dev, generator=g) / math.sqrt(cs)).to(torch.bfloat16)
    ref = F.conv2d(torch.cat([a1, a2], 1).float(), w3.float(), padding=1) + F.conv2d(xs.float(), w1.float())
    wt = torch.cat([k.pack_conv_weight(w3[:, :c1]), k.pack_conv_weight(w3[:, c1:]), k.pack_conv_weight(w1)], dim=1).contiguous()
    out = torch.empty(B, H, W, cout, device=dev, dtype=torch.bfloat16)
    k.conv_gemm([(_nhwc(a1), c1, 0, c1, 9), (_nhwc(a2), c2, 0, c2, 9), (_nhwc(xs), cs, 0, cs, 1)], wt, cout, out,
                batch=B, h=H, w=W, transposed=True)
    torch.cuda.synchronize()
    assert _report("T 2 segments + skip", out.permute(0, 3, 1, 2), ref, T_RTOL) < T_RTOL


def test_transposed_timing():
    k = _kern()
    dev = "cuda"
    for (B, H, cin, cout) in [(64, 160, 96, 96), (64, 160, 192, 96), (64, 80, 96, 96), (64, 80, 192, 192), (64, 40, 192, 192),
                              (64, 160, 96, 6)]:
        a = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
        wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device=dev) / 30).to(torch.bfloat16))
        ns = k.ceil_to(cout, 8)
        out = torch.empty(B, H, H, ns, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * B * H * H * cin * cout * 9
        for label, kw in [("tap", dict(halo=False, transposed=False)), ("transposed", dict(transposed=True))]:
            for _ in range(2):
                k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, n_store=ns, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, n_store=ns, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"[conv timing] {H}x{H} {cin}->{cout} {label:10s}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s")


def _t_norm_case(name, B, H, W, cins, cout, seed=0, silu=True, skip_c=0):
    """Transposed mode with the fused GroupNorm(+SiLU) prologue: segments carry (scale, shift) tables and the
    kernel feeds act(x * scale + shift) to the MMA. Reference: the same map in fp32, rounded to bf16 (what
    gn_apply would have stored), then torch conv2d in fp32. Out-of-image taps must see zeros (padding of the
    normalised tensor), which a naive transform of the zero-filled halo would break (act(shift) != 0)."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    xs, coefs, normed = [], [], []
    for c in cins:
        x = torch.randn(B, c, H, W, device=dev, generator=g).to(torch.bfloat16)
        sc = torch.rand(B, c, device=dev, generator=g) + 0.5
        sh = torch.randn(B, c, device=dev, generator=g)
        y = x.float() * sc[:, :, None, None] + sh[:, :, None, None]
        if silu:
            y = F.silu(y)
        xs.append(x)
        coefs.append(torch.stack([sc, sh], dim=-1).contiguous())
        normed.append(y.to(torch.bfloat16).float())
    cin = sum(cins)
    wgt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
    ref = F.conv2d(torch.cat(normed, 1), wgt.float(), padding=1)
    parts, off = [], 0
    for c in cins:
        parts.append(k.pack_conv_weight(wgt[:, off:off + c]))
        off += c
    segs = [(_nhwc(x), c, 0, c, 9, cf, silu) for x, c, cf in zip(xs, cins, coefs)]
    if skip_c:
        xr = torch.randn(B, skip_c, H, W, device=dev, generator=g).to(torch.bfloat16)
        w1 = (torch.randn(cout, skip_c, 1, 1, device=dev, generator=g) / math.sqrt(skip_c)).to(torch.bfloat16)
        ref = ref + F.conv2d(xr.float(), w1.float())
        parts.append(k.pack_conv_weight(w1))
        segs.append((_nhwc(xr), skip_c, 0, skip_c, 1))
    wt = torch.cat(parts, dim=1).contiguous()
    out = torch.full((B, H, W, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    k.conv_gemm(segs, wt, cout, out, batch=B, h=H, w=W, transposed=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), f"{name}: non-finite"
    return _report(name, out.permute(0, 3, 1, 2), ref, T_NORM_RTOL)


# tanh.approx SiLU (rel. error 2^-11) can flip the bf16 rounding of an operand now and then: one extra bf16
# half-ulp on a few of the K products, far inside the output's own bf16 rounding.
T_NORM_RTOL = 2.0 ** -7


def test_transposed_fused_groupnorm_prologue():
    assert _t_norm_case("TN 96->96 32x32", 2, 32, 32, [96], 96) < T_NORM_RTOL
    assert _t_norm_case("TN 96->96 80x80 ragged", 2, 80, 80, [96], 96, seed=1) < T_NORM_RTOL
    assert _t_norm_case("TN 64->96 24x24 padded cin", 2, 24, 24, [64], 96, seed=2) < T_NORM_RTOL
    assert _t_norm_case("TN (96+96)->96 64x64 cat + skip", 2, 64, 64, [96, 96], 96, seed=3, skip_c=192) < T_NORM_RTOL
    assert _t_norm_case("TN (192+96)->192 32x32 cat", 2, 32, 32, [192, 96], 192, seed=4, skip_c=288) < T_NORM_RTOL
    assert _t_norm_case("TN 192->192 32x32 affine only", 2, 32, 32, [192], 192, seed=5, silu=False) < T_NORM_RTOL
    assert _t_norm_case("TN (192+192)->192 40x40 20-row tiles", 2, 40, 40, [192, 192], 192, seed=6, skip_c=384) < T_NORM_RTOL


def test_transposed_output_head():
    """The 6-channel output heads in the transposed kernel: 8 stored channels (pitch 8), fused GroupNorm+SiLU
    prologue, bias, and the upsampled pyramid as a 6-channel identity K segment (engine.BlockOps.head)."""
    k = _kern()
    dev = "cuda"
    for (B, H, W, with_res, seed) in ((2, 32, 32, True, 0), (2, 40, 40, True, 1), (3, 64, 64, False, 2)):
        g = torch.Generator(device=dev).manual_seed(seed)
        cin, cout = 96, 6
        x = torch.randn(B, cin, H, W, device=dev, generator=g).to(torch.bfloat16)
        sc = torch.rand(B, cin, device=dev, generator=g) + 0.5
        sh = torch.randn(B, cin, device=dev, generator=g)
        y = F.silu(x.float() * sc[:, :, None, None] + sh[:, :, None, None]).to(torch.bfloat16).float()
        wgt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
        bias = torch.zeros(32, device=dev)
        bias[:cout] = torch.randn(cout, device=dev, generator=g)
        ref = F.conv2d(y, wgt.float(), bias[:cout], padding=1)
        parts = [k.pack_conv_weight(wgt, n_pad=16)]
        segs = [(_nhwc(x), cin, 0, cin, 9, torch.stack([sc, sh], dim=-1).contiguous(), True)]
        if with_res:
            r = torch.randn(B, cout, H, W, device=dev, generator=g).to(torch.bfloat16)
            r8 = torch.zeros(B, H, W, 8, device=dev, dtype=torch.bfloat16)
            r8[..., :cout] = r.permute(0, 2, 3, 1)
            ref = ref + r.float()
            parts.append(k.pack_conv_weight(torch.eye(cout, device=dev).view(cout, cout, 1, 1).to(torch.bfloat16), n_pad=16))
            segs.append((r8, 8, 0, cout, 1))
        wt = torch.cat(parts, dim=1).contiguous()
        out = torch.full((B, H, W, 8), float("nan"), device=dev, dtype=torch.bfloat16)
        k.conv_gemm(segs, wt, cout, out, batch=B, h=H, w=W, n_store=8, n_tile=16, bias=bias, transposed=True)
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all(), "head: non-finite"
        assert (out[..., cout:].float() == 0).all(), "head: padding channels must be zero"
        assert _report(f"T head 96->6 {H}x{W} res={with_res}", out[..., :cout].permute(0, 3, 1, 2), ref, T_NORM_RTOL) < T_NORM_RTOL


def test_gn_coeffs_match_group_norm():
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(9)
    B, H, W, c0, c1 = 3, 16, 16, 192, 96
    C, groups = c0 + c1, 32
    a0 = torch.randn(B, H, W, c0, device=dev, generator=g).to(torch.bfloat16)
    a1 = (torch.randn(B, H, W, c1, device=dev, generator=g) * 2 + 1).to(torch.bfloat16)
    gamma, beta = torch.randn(C, device=dev, generator=g), torch.randn(C, device=dev, generator=g)
    s0, s1 = torch.zeros(B, c0, 2, device=dev), torch.zeros(B, c1, 2, device=dev)
    k.gn_chan_stats(a0, c0, s0)
    k.gn_chan_stats(a1, c1, s1)
    f0, f1 = torch.empty(B, c0, 2, device=dev), torch.empty(B, c1, 2, device=dev)
    k.gn_coeffs(s0, c0, s1, c1, gamma, beta, f0, f1, H * W, groups)
    torch.cuda.synchronize()
    x = torch.cat([a0, a1], dim=-1).float().permute(0, 3, 1, 2)
    ref = F.group_norm(x, groups, gamma, beta, eps=1e-6)
    coef = torch.cat([f0, f1], dim=1)
    got = x * coef[:, :, 0, None, None] + coef[:, :, 1, None, None]
    assert (got - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


@pytest.mark.parametrize("c0,c1,tiles", [(96, 0, 200), (192, 96, 50), (96, 288, 7), (288, 192, 20)])
def test_gn_coeffs_from_partials(c0, c1, tiles):
    """gn_coeffs_partials == gn_finalize_partials + gn_coeffs, for either source arriving as per-tile partials."""
    k = _kern()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(c0 + c1 + tiles)
    B, hw = 3, 1600
    C = c0 + c1
    groups = min(C // 4, 32)
    gamma, beta = torch.randn(C, device=dev, generator=g), torch.randn(C, device=dev, generator=g)

    def make(c):
        # partial sums of plausible data: sum ~ N(0, n), sumsq ~ n per tile
        p = torch.randn(B * tiles, c, 2, device=dev, generator=g) * 3
        p[..., 1] = p[..., 1].abs() * 4 + (hw / tiles)
        return p.contiguous()

    p0 = make(c0)
    s0 = p0.view(B, tiles, c0, 2).sum(1).contiguous()
    p1 = make(c1) if c1 else None
    s1 = p1.view(B, tiles, c1, 2).sum(1).contiguous() if c1 else None
    ref0, ref1 = torch.empty(B, c0, 2, device=dev), (torch.empty(B, c1, 2, device=dev) if c1 else None)
    k.gn_coeffs(s0, c0, s1, c1, gamma, beta, ref0, ref1, hw, groups)
    for mode in ((True, False), (False, True), (True, True)) if c1 else ((True, False),):
        f0 = torch.full((B, c0, 2), float("nan"), device=dev)
        f1 = torch.full((B, c1, 2), float("nan"), device=dev) if c1 else None
        o0 = torch.full((B, c0, 2), float("nan"), device=dev)
        o1 = torch.full((B, c1, 2), float("nan"), device=dev) if c1 else None
        a0 = (None, p0, tiles, o0, c0) if mode[0] else (s0, None, 0, None, c0)
        a1 = None
        if c1:
            a1 = (None, p1, tiles, o1, c1) if mode[1] else (s1, None, 0, None, c1)
        k.gn_coeffs_partials(a0, a1, gamma, beta, f0, f1, B, hw, groups)
        torch.cuda.synchronize()
        if mode[0]:
            assert (o0 - s0).abs().max().item() <= 1e-5 * s0.abs().max().item(), "reduced partials (source 0)"
        if c1 and mode[1]:
            assert (o1 - s1).abs().max().item() <= 1e-5 * s1.abs().max().item(), "reduced partials (source 1)"
        assert (f0 - ref0).abs().max().item() <= 2e-4 * ref0.abs().max().item(), f"coef0 mode={mode}"
        if c1:
            assert (f1 - ref1).abs().max().item() <= 2e-4 * ref1.abs().max().item(), f"coef1 mode={mode}"


# ---- fused attention core (csrc/attention.cu) --------------------------------------------------------------------
@pytest.mark.parametrize("c,hw,b,rescale", [(192, 20, 3, True), (288, 10, 5, True), (288, 5, 7, True), (256, 16, 2, True),
                                            (256, 8, 3, False), (288, 4, 2, False), (96, 20, 2, True), (128, 16, 1, False)])
def test_fused_attention_core_vs_oracle(c, hw, b, rescale):
    """AttnBlockpp / DDPM AttnBlock through BlockOps.attention (GroupNorm launch + one q|k|v GEMM + ONE fused
    QK^T / softmax / PV / projection / residual kernel) against the oracle's fp32 block, and against the separate-launch
    path it replaces (same bf16 roundings, so the two agree to ~1 bf16 ulp of the output)."""
    from types import SimpleNamespace
    from conditional_score_diffusion_b200 import engine as E
    from conditional_score_diffusion_b200.models import layerspp
    from oracle import ncsnpp as o_net
    k = _kern()
    L = hw * hw
    assert k.attn_core_supported(L, c), (L, c)
    torch.manual_seed(c + hw)
    blk = layerspp.AttnBlockpp(c, skip_rescale=rescale, init_scale=1.0).cuda()
    with torch.no_grad():
        for prm in blk.parameters():
            if prm.dim() == 1:
                prm.add_(0.1 * torch.randn_like(prm))
            else:
                prm.mul_(3.0)          # sharper softmax than the default init gives
    x = torch.randn(b, c, hw, hw, device="cuda").to(torch.bfloat16)
    sd = {"all_modules.0." + n: v.detach().cpu() for n, v in blk.state_dict().items()}
    ref = o_net.attn_block(sd, 0, x.float().cpu(), SimpleNamespace(skip_rescale=rescale))

    class _Net(torch.nn.Module):
        pass
    net = _Net()
    net.all_modules = torch.nn.ModuleList([blk])
    eng = E.NetEngine(net)
    eng.device = torch.device("cuda")
    pk = eng._pack_attn(blk, eng.device)
    outs = {}
    for fused in (True, False):
        rec = E.Recorder()
        pool = E.BufferPool(eng.device)
        ops = E.BlockOps(eng.device, pool, rec, torch.zeros(1 << 20, device="cuda"))
        ops.fused_attention = fused
        a = E.Act(_nhwc(x), c)
        out = ops.attention(pk, a, rescale)
        names = [getattr(fn, "__name__", "") for fn, _, _ in rec.ops]
        assert ("attn_core" in names) == fused
        rec.run()
        torch.cuda.synchronize()
        outs[fused] = out.t[..., :c].permute(0, 3, 1, 2).float().cpu()
        assert torch.isfinite(outs[fused]).all()
    scale = ref.abs().max().item()
    e_f = (outs[True] - ref).abs().max().item() / scale
    e_s = (outs[False] - ref).abs().max().item() / scale
    e_fs = (outs[True] - outs[False]).abs().max().item() / scale
    print(f"[attn] C={c} L={L} B={b}: fused vs oracle {e_f:.3e}, separate vs oracle {e_s:.3e}, fused vs separate {e_fs:.3e}")
    assert e_f < 1e-2 and e_fs < 1e-2
