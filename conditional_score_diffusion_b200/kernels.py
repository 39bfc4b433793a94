"""Torch-tensor front ends of the C ABI (pointer extraction, shape checks, stream).

torch is used here for device memory and streams only. Every function launches CUDA kernels from
libcsd_b200.so on torch's current stream and raises on any error; none has a PyTorch fallback.
"""
import ctypes
import math
import os

import torch

from . import _lib
from ._lib import ConvGemmDesc, check

_BF16 = torch.bfloat16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.CsdError("libcsd_b200 kernels need CUDA tensors; there is no CPU path")


def ceil_to(v, m):
    return (v + m - 1) // m * m


# --------------------------------------------------------------------------------------------
# weight packing for csd_conv_gemm
# --------------------------------------------------------------------------------------------
def pack_conv_weight(weight, n_pad=None, dtype=_BF16):
    """[Cout, Cin, kh, kw] conv weight -> Wt [n_pad, taps * ceil32(Cin)] K-major.

    K order is (tap = ky*kw_ + kx, channel), channels zero padded to a multiple of 32 per tap,
    matching the (segment, tap, chunk) loop of conv_gemm_kernel.
    """
    cout, cin, kh, kw = weight.shape
    cpad = ceil_to(cin, 32)
    n_pad = n_pad or ceil_to(cout, 16)
    w = weight.detach().to(torch.float32).permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    out = torch.zeros(n_pad, kh * kw, cpad, dtype=torch.float32, device=weight.device)
    out[:cout, :, :cin] = w
    out = out.reshape(n_pad, kh * kw * cpad)
    if dtype == torch.float32:
        return round_tf32(out).contiguous()      # tf32 plan: operands are stored rounded (see csrc/common.cuh round_tf32)
    return out.to(dtype).contiguous()


def round_tf32(t):
    """fp32 -> nearest tf32 value (10 mantissa bits, ties away from zero = PTX cvt.rna.tf32.f32), kept in fp32 words:
    what csd_pack_weights (kind 2) and the operand-producing kernels of the tf32 plan store."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


_TILE_CACHE = {}
# Halo mode (one activation load per 32-channel chunk feeds all 9 taps) is opt-in: measured on B200 it
# does not beat the per-tap kernel while pixels are the M operand (tcgen05.mma has a ~100-cycle floor per
# M=128 instruction, tools/mma_rate_probe.cu), see DESIGN.md.
HALO_DEFAULT = os.environ.get("CSD_HALO", "0") == "1"


def pick_tile(h, w, batch):
    """Pixel box (tile_w, tile_h, tile_b), product <= 128, for the 128-row MMA tile.

    Minimises wasted MMA rows, then the halo ratio (squarer boxes re-read less through L2),
    then prefers wider boxes (longer contiguous TMA rows).
    """
    key = (h, w, batch)
    if key in _TILE_CACHE:
        return _TILE_CACHE[key]
    best = None
    for tw in range(1, min(w, 128) + 1):
        for th in range(1, min(h, 128 // tw) + 1):
            tb = max(1, min(batch, 128 // (tw * th)))
            tiles = math.ceil(w / tw) * math.ceil(h / th) * math.ceil(batch / tb)
            waste = tiles * 128 - w * h * batch
            halo = (tw + 2) * (th + 2) * tb
            cand = ((waste, halo, -tw), (tw, th, tb))
            if best is None or cand[0] < best[0]:
                best = cand
    _TILE_CACHE[key] = best[1]
    return best[1]


TRANSPOSED_DEFAULT = os.environ.get("CSD_NO_TRANSPOSED", "0") != "1"
# GroupNorm+SiLU applied inside the transposed convolution (no normalised copy in HBM); CSD_NO_FUSE_GN=1 keeps
# the separate gn_apply pass (A/B measurements).
FUSE_GN_DEFAULT = os.environ.get("CSD_NO_FUSE_GN", "0") != "1"


# Transposed kernel on images whose width is not a multiple of the 8-pixel tile width (20 px level: 3 tile columns, the
# last one half empty - TMA clips the loads and the store, the epilogue's ragged statistics path counts valid pixels only).
# bf16 instance only (the fp32 epilogue has no ragged path). Parity-tested, but measured slower than the per-tap kernel +
# one-launch GroupNorm at that level (29.15 vs 28.96 ms per PC step: 3 tile columns for 2.5, 192 channels on 256 M rows,
# 384 tiles on 148 SMs), so it is off unless CSD_T_RAGGED=1.
T_RAGGED = os.environ.get("CSD_T_RAGGED", "0") == "1"


def transposed_eligible(segments, h, w, stride=1, pad=1, z_batches=1, allow_1tap=False, ragged=False):
    """3x3 stride-1 convs whose image tiles well into 32x8-pixel macro tiles run in the transposed halo
    mode (output channels on M, 256 pixels on N): measured faster on B200 whenever the 32-row tiling
    wastes < ~20% of the rows (tests/test_gpu_conv_gemm.py::test_transposed_timing)."""
    return ((segments[0][4] == 9 or allow_1tap) and stride == 1 and pad == 1 and z_batches == 1
            and transposed_shape_ok(h, w, ragged and segments[0][0].dtype == _BF16))


def transposed_shape_ok(h, w, ragged=False):
    w_ok = w % 8 == 0 or (ragged and T_RAGGED and w >= 16)
    return w_ok and (h % 32 == 0 or h >= 64 or h % 20 == 0)


def transposed_tile_rows(h):
    """Rows of the transposed kernel's macro tile: mirrors pick_t_rows() in csrc/conv_gemm.cu (tiles x cycles per
    instruction, cycles ~ max(100, N / 2) with N = 8 * rows): 160 -> 32, 80 -> 28, 40 -> 20."""
    if os.environ.get("CSD_TROWS_LEGACY", "0") != "0":
        return 20 if (h % 32 != 0 and h < 64 and h % 20 == 0) else 32
    best, best_cost = 32, None
    for t in (32, 28, 24, 20):
        cost = math.ceil(h / t) * max(100, t * 4)
        if best_cost is None or cost <= best_cost:
            best, best_cost = t, cost
    return best


def conv_gemm(segments, wt, n, out, *, batch, h, w, out_pitch=None, n_store=None, n_tile=None,
              tile=None, bias=None, bias_per_row=False, temb=None, temb_pitch=0, res=None,
              res_pitch=0, scale=1.0, out_f32=None, z_batches=1, a_batch_step=0, wt_batch_stride=0,
              out_z_stride=0, res_z_stride=0, wt_pitch=0, wt_k_off=0, k_valid=0, wt_rows=None,
              stride=1, pad=1, in_h=0, in_w=0, halo=None, mt=None, transposed=None, stat_partials=None,
              round_out=False, k_splits=1, splitk_ws=None):
    """Launch csd_conv_gemm. segments: list of (tensor, pitch, c_off, c_cnt, taps[, norm, norm_silu]): `norm` is the
    [batch, c_cnt, 2] (scale, shift) table of gn_coeffs for the fused GroupNorm(+SiLU) prologue (transposed mode)."""
    _require_cuda(wt, out, bias, temb, res, *[s[0] for s in segments])
    segments = [tuple(s) + (None, True)[len(s) - 5:] for s in segments]
    d = ConvGemmDesc()
    d.batch, d.h, d.w = batch, h, w
    d.in_h, d.in_w, d.stride, d.pad = in_h, in_w, stride, pad
    tw, th, tb = tile if tile is not None else pick_tile(h, w, batch)
    d.tile_w, d.tile_h, d.tile_b = tw, th, tb
    n_store_ = n_store if n_store is not None else n
    nt_ = n_tile if n_tile is not None else None
    d.nseg = len(segments)
    k_total = 0
    tf32 = segments[0][0].dtype == torch.float32     # fp32 activations + weights: the kind::tf32 per-tap kernel
    for i, (a, pitch, c_off, c_cnt, taps, norm, norm_silu) in enumerate(segments):
        assert a.dtype == (torch.float32 if tf32 else _BF16), "conv_gemm: mixed activation precisions"
        d.seg[i].a = a.data_ptr()
        d.seg[i].pitch = pitch
        d.seg[i].c_off = c_off
        d.seg[i].c_cnt = c_cnt
        d.seg[i].taps = taps
        if norm is not None:
            assert norm.dtype == torch.float32 and norm.is_cuda and norm.numel() == batch * c_cnt * 2
            d.seg[i].norm = norm.data_ptr()
            d.seg[i].norm_silu = int(bool(norm_silu))
        k_total += taps * ceil_to(c_cnt, 32)
    d.n = n
    d.n_store = n_store if n_store is not None else n
    if n_tile is None:
        n16 = ceil_to(d.n_store, 16)
        n_tile = n16 if n16 <= 256 else (ceil_to(n16 // 2, 16) if n16 <= 512 else 256)
    d.n_tile = n_tile
    if tf32:
        if halo:
            raise _lib.CsdError("the fp32 / tf32 plan runs in the per-tap and the transposed kernels only")
        assert wt.dtype == torch.float32 and out.dtype == torch.float32 and (res is None or res.dtype == torch.float32)
        halo = False
        transposed = bool(transposed)
        d.dtype = 1
        d.out_round_tf32 = int(bool(round_out))
    if halo is None:
        halo = (HALO_DEFAULT and tile is None and segments[0][4] == 9 and stride == 1 and pad == 1 and z_batches == 1
                and w % 8 == 0 and h >= 16 and n_tile <= 512)
    if transposed is None:
        transposed = (TRANSPOSED_DEFAULT and not halo and tile is None and n >= 32
                      and transposed_eligible(segments, h, w, stride, pad, z_batches)
                      and out.dtype == _BF16 and not bias_per_row)
    if transposed:
        d.mode, d.mt = 2, 2
        d.stat_partials = stat_partials.data_ptr() if stat_partials is not None else None
    elif halo:
        if mt is None:
            mt = 2 if (2 * n_tile <= 256 and h % 32 == 0) else 1
        d.mode, d.mt = 1, mt
    d.wt = wt.data_ptr()
    d.wt_rows = wt_rows if wt_rows is not None else wt.shape[-2]
    d.k_total = k_total
    d.wt_pitch = wt_pitch
    d.wt_k_off = wt_k_off
    d.k_valid = k_valid
    d.wt_batch_stride = wt_batch_stride
    d.z_batches = z_batches
    d.a_batch_step = a_batch_step
    d.out = out.data_ptr()
    d.out_pitch = out_pitch if out_pitch is not None else out.shape[-1]
    d.out_f32 = int(out.dtype == torch.float32) if out_f32 is None else int(out_f32)
    d.out_z_stride = out_z_stride
    d.bias = bias.data_ptr() if bias is not None else None
    d.bias_per_row = int(bias_per_row)
    d.temb = temb.data_ptr() if temb is not None else None
    d.temb_pitch = temb_pitch
    d.res = res.data_ptr() if res is not None else None
    d.res_pitch = res_pitch
    d.res_z_stride = res_z_stride
    d.scale = scale
    if k_splits > 1:
        assert (splitk_ws is not None and splitk_ws.dtype == torch.float32 and splitk_ws.is_cuda
                and splitk_ws.numel() >= k_splits * batch * h * w * ceil_to(d.n_store, 8)), "split-K workspace"
        d.k_splits = k_splits
        d.splitk_ws = splitk_ws.data_ptr()
    check(_lib.lib().csd_conv_gemm(ctypes.byref(d), _stream()))
    return out


# Split-K for the small levels: at most this many CTAs share one tile's K range (0 / 1 = off; A/B switch).
SPLITK_MAX = int(os.environ.get("CSD_SPLITK_MAX", "4"))


def pick_k_splits(segments, batch, h, w, n_store, n_tile, tile=None, sms=148):
    """How many ways csd_conv_gemm's per-tap kernel should split K for this launch. A level with fewer 128-pixel
    tiles than half the SMs (5 / 10 px at batch 64) leaves most of the GPU idle while each CTA walks K = 9 C alone,
    bound by its own TMA ring; sharing K between idle SMs shortens that walk. At least 4 ring iterations per split."""
    if SPLITK_MAX <= 1:
        return 1
    tw, th, tb = tile if tile is not None else pick_tile(h, w, batch)
    tiles = -(-w // tw) * -(-h // th) * -(-batch // tb)
    ctas = tiles * -(-n_store // n_tile)
    if ctas * 2 > sms:
        return 1
    f32 = segments[0][0].dtype == torch.float32
    chunk = 64 if (not f32 and any(sg[3] > 32 for sg in segments)) else 32
    iters = sum(sg[4] * -(-sg[3] // chunk) for sg in segments)
    return max(1, min(SPLITK_MAX, sms // ctas, iters // 4))


# --------------------------------------------------------------------------------------------
# module-surface ops (NCHW fp32)
# --------------------------------------------------------------------------------------------
def upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    oh, ow = ctypes.c_int(0), ctypes.c_int(0)
    check(_lib.lib().csd_upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, px0, px1, py0, py1,
                                            ctypes.byref(oh), ctypes.byref(ow)))
    return oh.value, ow.value


def upfirdn2d_planes(x, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """x [planes, H, W] fp32 contiguous CUDA, kernel [kh, kw] fp32 CUDA -> [planes, OH, OW]."""
    _require_cuda(x, kernel)
    assert x.dtype == torch.float32 and kernel.dtype == torch.float32
    x = x.contiguous()
    kernel = kernel.contiguous()
    planes, in_h, in_w = x.shape
    kh, kw = kernel.shape
    oh, ow = upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, px0, px1, py0, py1)
    out = torch.empty(planes, oh, ow, device=x.device, dtype=torch.float32)
    check(_lib.lib().csd_upfirdn2d_f32(_ptr(x), _ptr(kernel), _ptr(out), planes, in_h, in_w, kh, kw, up_x, up_y,
                                       down_x, down_y, px0, px1, py0, py1, _stream()))
    return out


def fused_bias_act(x, bias, refer, act, grad, alpha, scale):
    """op/fused_bias_act semantics over a contiguous fp32 tensor; bias broadcasts over dim 1."""
    _require_cuda(x, bias, refer)
    x = x.contiguous()
    out = torch.empty_like(x)
    size_b, step_b = 0, 1
    if bias is not None and bias.numel() > 0:
        size_b = bias.numel()
        step_b = 1
        for d in x.shape[2:]:
            step_b *= d
    else:
        bias = None
    if refer is not None and refer.numel() == 0:
        refer = None
    check(_lib.lib().csd_fused_bias_act_f32(_ptr(x), _ptr(bias), _ptr(refer.contiguous() if refer is not None else None),
                                            _ptr(out), x.numel(), size_b, step_b, act, grad, alpha, scale, _stream()))
    return out


# --------------------------------------------------------------------------------------------
# PC update kernels (fp32, contiguous, [batch, ...])
# --------------------------------------------------------------------------------------------
def _per_sample(t):
    b = t.shape[0]
    return b, t.numel() // b


def ve_perturb(y, z, out, sigma_tab, step_idx=None, sample_stride=0):
    _require_cuda(y, z, out, sigma_tab, step_idx)
    b, ps = _per_sample(y)
    check(_lib.lib().csd_ve_perturb_f32(_ptr(y), _ptr(z), _ptr(out), b, ps, _ptr(sigma_tab), _ptr(step_idx),
                                        sample_stride, _stream()))
    return out


def sde_perturb(x, z, out, mean_coef, std):
    """out = mean_coef[b] * x + std[b] * z (mean_coef None = 1)."""
    _require_cuda(x, z, out, mean_coef, std)
    b, ps = _per_sample(x)
    check(_lib.lib().csd_sde_perturb_f32(_ptr(x), _ptr(z), _ptr(out), b, ps, _ptr(mean_coef), _ptr(std), _stream()))
    return out


def inpaint_merge(x, data, z, mask, x_out, x_mean, mean_coef, std):
    _require_cuda(x, data, z, mask, x_out, x_mean, mean_coef, std)
    b, ps = _per_sample(x)
    check(_lib.lib().csd_inpaint_merge_f32(_ptr(x), _ptr(data), _ptr(z), _ptr(mask), _ptr(x_out), _ptr(x_mean), b, ps,
                                           _ptr(mean_coef), _ptr(std), _stream()))
    return x_out, x_mean


def dsm_loss(score, z, a, c, w, losses):
    """losses[b] += w[b] * sum((a[b] * score + c[b] * z)^2); `losses` is pre-zeroed by the caller."""
    _require_cuda(score, z, a, c, w, losses)
    b, ps = _per_sample(score)
    check(_lib.lib().csd_dsm_loss_f32(_ptr(score), _ptr(z), _ptr(a), _ptr(c), _ptr(w), _ptr(losses), b, ps, _stream()))
    return losses


def langevin_norms(grad, noise, norms):
    b, ps = _per_sample(grad)
    check(_lib.lib().csd_langevin_norms_f32(_ptr(grad), _ptr(noise), _ptr(norms), b, ps, _stream()))
    return norms


def langevin_update(x, grad, noise, norms, x_out, x_mean, snr, alpha_tab=None, step_idx=None, sample_stride=0):
    b, ps = _per_sample(x)
    check(_lib.lib().csd_langevin_update_f32(_ptr(x), _ptr(grad), _ptr(noise), _ptr(norms), _ptr(x_out), _ptr(x_mean),
                                             b, ps, float(snr), _ptr(alpha_tab), _ptr(step_idx), sample_stride,
                                             _stream()))


def reverse_diffusion_update(x, score, noise, x_out, x_mean, f_tab, g_tab, probability_flow=False, step_idx=None,
                             sample_stride=0):
    b, ps = _per_sample(x)
    check(_lib.lib().csd_reverse_diffusion_update_f32(_ptr(x), _ptr(score), _ptr(noise), _ptr(x_out), _ptr(x_mean),
                                                      b, ps, _ptr(f_tab), _ptr(g_tab), int(probability_flow),
                                                      _ptr(step_idx), sample_stride, _stream()))


def euler_maruyama_update(x, score, noise, x_out, x_mean, d_tab, g_tab, dt, probability_flow=False, step_idx=None,
                          sample_stride=0):
    b, ps = _per_sample(x)
    check(_lib.lib().csd_euler_maruyama_update_f32(_ptr(x), _ptr(score), _ptr(noise), _ptr(x_out), _ptr(x_mean),
                                                   b, ps, _ptr(d_tab), _ptr(g_tab), float(dt),
                                                   int(probability_flow), _ptr(step_idx), sample_stride, _stream()))


def broadcast_table(dst, tab, step_idx=None, sample_stride=0):
    check(_lib.lib().csd_broadcast_table_f32(_ptr(dst), dst.numel(), _ptr(tab), _ptr(step_idx), sample_stride,
                                             _stream()))
    return dst


def step_advance(step_idx):
    check(_lib.lib().csd_step_advance(_ptr(step_idx), _stream()))


# --------------------------------------------------------------------------------------------
# score-network building blocks (NHWC bf16)
# --------------------------------------------------------------------------------------------
def nchw_to_nhwc(src0, src1, out, scale=1.0, shift=0.0):
    b, c0, h, w = src0.shape
    c1 = src1.shape[1] if src1 is not None else 0
    fn = _lib.lib().csd_nchw_to_nhwc_f32 if out.dtype == torch.float32 else _lib.lib().csd_nchw_to_nhwc_bf16
    check(fn(_ptr(src0), c0, _ptr(src1), c1, _ptr(out), out.shape[-1], b, h, w, float(scale), float(shift), _stream()))
    return out


def nhwc_to_nchw(src, c_off, c_cnt, dst, row_scale=None):
    b, h, w, pitch = src.shape
    fn = _lib.lib().csd_nhwc_f32_to_nchw if src.dtype == torch.float32 else _lib.lib().csd_nhwc_bf16_to_nchw
    check(fn(_ptr(src), pitch, c_off, c_cnt, _ptr(dst), b, h, w, _ptr(row_scale), _stream()))
    return dst


def gn_chan_stats(src, c, chan_sums):
    """src [B, H, W, pitch] (or [B, HW, pitch]) bf16 or fp32; chan_sums [B, c, 2] fp32 (stored, deterministic)."""
    b = src.shape[0]
    hw = src.numel() // (b * src.shape[-1])
    fn = _lib.lib().csd_gn_chan_stats_f32 if src.dtype == torch.float32 else _lib.lib().csd_gn_chan_stats_bf16
    check(fn(_ptr(src), c, src.shape[-1], _ptr(chan_sums), b, hw, _stream()))
    return chan_sums


def gn_finalize_partials(partials, chan_sums, batch, tiles_per_img, c):
    check(_lib.lib().csd_gn_finalize_partials_f32(_ptr(partials), _ptr(chan_sums), batch, tiles_per_img, c, _stream()))
    return chan_sums


def gn_coeffs(sums0, c0, sums1, c1, gamma, beta, coef0, coef1, hw, groups, eps=1e-6):
    """(scale, shift) tables of GroupNorm over cat(src0, src1) for the fused conv prologue: coef_i [B, c_i, 2]."""
    b = sums0.shape[0]
    check(_lib.lib().csd_gn_coeffs_f32(_ptr(sums0), c0, _ptr(sums1), c1, _ptr(gamma), _ptr(beta), _ptr(coef0),
                                       _ptr(coef1), b, hw, groups, float(eps), _stream()))
    return coef0, coef1


def _gn_flags(silu, round_out, dtype):
    """apply_silu argument: bit 0 = SiLU, bit 1 (fp32 tensors) = round the result to tf32 (it feeds an MMA operand)."""
    return int(bool(silu)) | (2 if (round_out and dtype == torch.float32) else 0)


def gn_apply(src0, c0, sums0, src1, c1, sums1, gamma, beta, out, groups, eps=1e-6, silu=True, round_out=False):
    b = src0.shape[0]
    hw = src0.numel() // (b * src0.shape[-1])
    fn = _lib.lib().csd_gn_apply_f32 if src0.dtype == torch.float32 else _lib.lib().csd_gn_apply_bf16
    check(fn(_ptr(src0), c0, src0.shape[-1], _ptr(sums0), _ptr(src1), c1,
             src1.shape[-1] if src1 is not None else 0, _ptr(sums1), _ptr(gamma), _ptr(beta),
             _ptr(out), out.shape[-1], b, hw, groups, float(eps), _gn_flags(silu, round_out, src0.dtype), _stream()))
    return out


def gn_coeffs_partials(src0, src1, gamma, beta, coef0, coef1, batch, hw, groups, eps=1e-6):
    """srcN = (sums, partials, tiles, sums_out, c) with exactly one of sums / partials set (src1 may be None)."""
    s1 = src1 if src1 is not None else (None, None, 0, None, 0)
    check(_lib.lib().csd_gn_coeffs_partials_f32(_ptr(src0[0]), _ptr(src0[1]), int(src0[2]), _ptr(src0[3]), int(src0[4]),
                                                _ptr(s1[0]), _ptr(s1[1]), int(s1[2]), _ptr(s1[3]), int(s1[4]),
                                                _ptr(gamma), _ptr(beta), _ptr(coef0), _ptr(coef1), batch, hw, groups,
                                                float(eps), _stream()))
    return coef0


def gn_fused_supported(c0, c1, hw, groups, batch, dtype=_BF16):
    """Host-side test: does the one-launch GroupNorm (statistics + apply) take this shape?"""
    fn = _lib.lib().csd_gn_fused_supported_f32 if dtype == torch.float32 else _lib.lib().csd_gn_fused_supported
    return bool(fn(int(c0), int(c1), int(hw), int(groups), int(batch)))


def gn_fused(src0, c0, src1, c1, gamma, beta, out, groups, eps=1e-6, silu=True, round_out=False):
    b = src0.shape[0]
    hw = src0.numel() // (b * src0.shape[-1])
    fn = _lib.lib().csd_gn_fused_f32 if src0.dtype == torch.float32 else _lib.lib().csd_gn_fused_bf16
    check(fn(_ptr(src0), c0, src0.shape[-1], _ptr(src1), c1,
             src1.shape[-1] if src1 is not None else 0, _ptr(gamma), _ptr(beta), _ptr(out),
             out.shape[-1], b, hw, groups, float(eps), _gn_flags(silu, round_out, src0.dtype), _stream()))
    return out


def fir_resample(src, out, mode, taps, add=None, round_out=False, norm=None, norm_silu=True):
    """mode 'up' | 'down' | 'prefilter'; src/out NHWC bf16 (fp32 in the tf32 plan). norm: [batch, c, 2] fp32 (scale, shift)
    table of gn_coeffs - GroupNorm(+SiLU) is then applied to the input on the fly (up / down only)."""
    b, h, w, pitch = src.shape
    arr = (ctypes.c_float * 4)(*[float(t) for t in taps])
    f32 = src.dtype == torch.float32
    m = {"up": 1, "down": 2, "prefilter": 3}[mode] | (0x10 if (round_out and f32) else 0)
    if norm is not None:
        assert norm.dtype == torch.float32 and norm.is_cuda and norm.shape[0] == b and norm.shape[2] == 2
        fn = _lib.lib().csd_fir_norm_resample_nhwc_f32 if f32 else _lib.lib().csd_fir_norm_resample_nhwc_bf16
        check(fn(_ptr(src), _ptr(out), _ptr(add), _ptr(norm), norm.shape[1], int(bool(norm_silu)), b, h, w, pitch, m, arr,
                 _stream()))
        return out
    fn = _lib.lib().csd_fir_resample_nhwc_f32 if f32 else _lib.lib().csd_fir_resample_nhwc_bf16
    check(fn(_ptr(src), _ptr(out), _ptr(add), b, h, w, pitch, m, arr, _stream()))
    return out


# GroupNorm_0 + SiLU of a resampling residual block applied inside the FIR kernel (no normalised tensor in HBM) when the
# block input's channel sums are already known. Parity-tested, but the FIR kernel is instruction-issue bound and the
# extra transform costs more than the saved gn_apply pass: 28.48 vs 28.34 ms per PC step (bf16), 52.31 vs 52.10 (tf32).
# Off unless CSD_FIR_NORM=1.
FIR_NORM_DEFAULT = os.environ.get("CSD_FIR_NORM", "0") == "1" and os.environ.get("CSD_FIR_PER_OUTPUT") is None


def tap_shift_sum(partial, cout, bias, res, out):
    """out[b,y,x,co] = bias[co] + res[...] + sum_t partial[b, y+dy_t, x+dx_t, t*cout + co] (tap-stacked output heads)."""
    b, h, w, pitch = partial.shape
    check(_lib.lib().csd_tap_shift_sum_bf16(_ptr(partial), pitch, cout, _ptr(bias), _ptr(res),
                                            res.shape[-1] if res is not None else 0, _ptr(out), out.shape[-1], b, h, w,
                                            _stream()))
    return out


def softmax_rows(logits, probs, cols, scale):
    rows = logits.numel() // logits.shape[-1]
    fn = _lib.lib().csd_softmax_rows_f32_f32 if probs.dtype == torch.float32 else _lib.lib().csd_softmax_rows_f32_bf16
    check(fn(_ptr(logits), logits.shape[-1], _ptr(probs), probs.shape[-1], rows, cols, float(scale), _stream()))
    return probs


# Fused attention core (csd_attn_core_bf16); CSD_NO_FUSED_ATTN=1 keeps the separate GEMM / softmax launches (A/B).
FUSED_ATTN_DEFAULT = os.environ.get("CSD_NO_FUSED_ATTN", "0") != "1"


def attn_core_supported(L, c):
    return FUSED_ATTN_DEFAULT and bool(_lib.lib().csd_attn_core_supported(int(L), int(c)))


def attn_core(qkv, wo, bo, res, out, batch, L, c, out_scale):
    """qkv [B, L, >=3c] bf16 (q | k | v), wo packed NIN_3 weights [rows >= c, pitch] bf16, bo fp32, res / out
    [B, L, pitch] bf16: out = (res + softmax(q k^T c^-0.5) v Wo + bo) * out_scale, one launch."""
    _require_cuda(qkv, wo, bo, res, out)
    assert qkv.dtype == _BF16 and wo.dtype == _BF16 and res.dtype == _BF16 and out.dtype == _BF16
    check(_lib.lib().csd_attn_core_bf16(_ptr(qkv), qkv.shape[-1], _ptr(wo), wo.shape[-1], wo.shape[-2], _ptr(bo),
                                        _ptr(res), res.shape[-1], _ptr(out), out.shape[-1], batch, L, c,
                                        float(out_scale), _stream()))
    return out


def time_embedding(labels, nf, embedding_type, fourier_w, w0, b0, w1, b1, out):
    check(_lib.lib().csd_time_embedding_f32(_ptr(labels), labels.shape[0], nf, 1 if embedding_type == "fourier" else 0,
                                            _ptr(fourier_w), _ptr(w0), _ptr(b0), _ptr(w1), _ptr(b1), _ptr(out), _stream()))
    return out


def dense_rows(act, w, bias, out, total_out=None):
    b, in_dim = act.shape
    total_out = total_out if total_out is not None else w.shape[0]
    check(_lib.lib().csd_dense_rows_f32(_ptr(act), _ptr(w), _ptr(bias), _ptr(out), b, in_dim, total_out, _stream()))
    return out


def rk_combine(y, k_stack, stages, coefs, h, out):
    _require_cuda(y, k_stack, out)
    arr = (ctypes.c_float * stages)(*[float(c) for c in coefs[:stages]])
    check(_lib.lib().csd_rk_combine_f32(_ptr(y), _ptr(k_stack), y.numel(), stages, arr, float(h), _ptr(out), _stream()))
    return out


REDUCE_WS = 1024   # CSD_REDUCE_WS_FLOATS: workspace of the deterministic device-wide sums (result in [0])


def reduce_workspace(device):
    """Zeroed workspace for rk_error_sumsq / sumsq (allocate once, reuse: the kernels keep its ticket counter at zero)."""
    return torch.zeros(REDUCE_WS, device=device, dtype=torch.float32)


def rk_error_sumsq(k_stack, stages, e, h, y, y2, atol, rtol, out):
    _require_cuda(y, y2, k_stack, out)
    assert out.numel() >= REDUCE_WS, "rk_error_sumsq: `out` is a kernels.reduce_workspace()"
    arr = (ctypes.c_float * stages)(*[float(c) for c in e[:stages]])
    check(_lib.lib().csd_rk_error_sumsq_f32(_ptr(k_stack), y.numel(), stages, arr, float(h), _ptr(y), _ptr(y2), float(atol),
                                            float(rtol), _ptr(out), _stream()))
    return out


def sumsq(x, out):
    assert out.numel() >= REDUCE_WS, "sumsq: `out` is a kernels.reduce_workspace()"
    check(_lib.lib().csd_sumsq_f32(_ptr(x), x.numel(), _ptr(out), _stream()))
    return out


def fused_adam_ema(p, g, m, v, ema, lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, max_norm, gnorm_sq,
                   ema_decay):
    _require_cuda(p, g, m, v, ema, gnorm_sq)
    check(_lib.lib().csd_fused_adam_ema_f32(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(ema), p.numel(), float(lr), float(beta1),
                                            float(beta2), float(eps), float(weight_decay), float(bias_corr1),
                                            float(bias_corr2), float(max_norm), _ptr(gnorm_sq), float(ema_decay), _stream()))


class PackTable:
    """Device-resident job table of csd_pack_weights (built once per packed network; one launch per refresh)."""

    def __init__(self, jobs, device):
        import numpy as np
        self.n = len(jobs)
        self.max_elems = max([1] + [j.rows * j.cols * j.taps if j.kind == 0 else j.rows for j in jobs])
        arr = (_lib.PackJob * self.n)(*jobs)
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.dev = torch.from_numpy(raw).to(device)

    def run(self):
        if self.n:
            check(_lib.lib().csd_pack_weights(_ptr(self.dev), self.n, self.max_elems, _stream()))


def pack_job(src, dst, rows, cols, taps, k_pad, dst_pitch, s_row, s_col, s_tap, flip=False, scale=1.0, src_off=0, dst_off=0):
    """kind-0 / kind-2 job; src fp32 tensor (+ element offset), dst bf16 or fp32 tensor (+ element offset)."""
    j = _lib.PackJob()
    j.src = src.data_ptr() + 4 * src_off
    j.src2 = None
    j.dst = dst.data_ptr() + dst.element_size() * dst_off
    j.s_row, j.s_col, j.s_tap = s_row, s_col, s_tap
    j.rows, j.cols, j.taps, j.k_pad, j.dst_pitch = rows, cols, taps, k_pad, dst_pitch
    j.flip, j.kind, j.scale = int(flip), (2 if dst.dtype == torch.float32 else 0), float(scale)
    return j


def bias_job(dst, dst_off, src, src2=None):
    j = _lib.PackJob()
    j.src, j.src2 = src.data_ptr(), (src2.data_ptr() if src2 is not None else None)
    j.dst = dst.data_ptr() + 4 * dst_off
    j.rows, j.cols, j.taps, j.kind, j.scale = src.numel(), 1, 1, 1, 1.0
    return j


# --------------------------------------------------------------------------------------------
# training backward (adjoints of the kernels above)
# --------------------------------------------------------------------------------------------
def pixmajor_geometry(batch, grid_h, grid_w):
    g = _lib.PixMajorGeom()
    check(_lib.lib().csd_pixmajor_geometry(batch, grid_h, grid_w, ctypes.byref(g)))
    return g


def pixmajor_alloc(geom, channels, ncopies, device):
    """Zero-filled pixel-major buffer [copies, splits, channels, row_pitch] bf16 (the zeros are the grid padding;
    csd_nhwc_to_pixmajor_bf16 never writes them, so the buffer is zeroed once and reused for the same pattern)."""
    return torch.zeros(ncopies, geom.splits, channels, geom.row_pitch, device=device, dtype=_BF16)


def nhwc_to_pixmajor(src, c_off, c_cnt, geom, out, stride=1, offset=0):
    b, h, w, pitch = src.shape
    check(_lib.lib().csd_nhwc_to_pixmajor_bf16(_ptr(src), pitch, c_off, c_cnt, b, h, w, stride, offset,
                                               ctypes.byref(geom), out.shape[0], _ptr(out), _stream()))
    return out


def wgrad_gemm(g_pm, cout, a_pm, cin, taps, geom, partial):
    check(_lib.lib().csd_wgrad_gemm_bf16(_ptr(g_pm), cout, _ptr(a_pm), cin, taps, ctypes.byref(geom), _ptr(partial),
                                         _stream()))
    return partial


# MN-major direct wgrad kernel (csd_wgrad_direct_bf16); CSD_NO_WGRAD_DIRECT=1 keeps the pixel-major GEMM path everywhere
WGRAD_DIRECT = os.environ.get("CSD_NO_WGRAD_DIRECT", "0") != "1"


def wgrad_direct_splits(batch, h, w, cout, cin, taps):
    s = ctypes.c_int(0)
    check(_lib.lib().csd_wgrad_direct_splits(batch, h, w, cout, cin, taps, ctypes.byref(s)))
    return s.value


def wgrad_direct(g, g_c_off, cout, a, a_c_off, cin, taps, partial, splits):
    """g, a: NHWC bf16 [B, h, w, pitch] on the same (activation) grid."""
    b, h, w, _ = a.shape
    assert g.shape[:3] == a.shape[:3]
    check(_lib.lib().csd_wgrad_direct_bf16(_ptr(g), g.shape[-1], g_c_off, cout, _ptr(a), a.shape[-1], a_c_off, cin, taps,
                                           b, h, w, _ptr(partial), splits, _stream()))
    return partial


def wgrad_reduce(partial, splits, taps, cout, cin, scale, dw, stride_co, stride_ci, stride_tap, ci_off=0,
                 accumulate=False):
    check(_lib.lib().csd_wgrad_reduce_f32(_ptr(partial), splits, taps, cout, cin, float(scale), _ptr(dw), stride_co,
                                          stride_ci, stride_tap, ci_off, int(accumulate), _stream()))
    return dw


def gn_bwd_stats(x, c, dy, dy_c_off, fwd_coef, s, s_c_off, silu):
    b = x.shape[0]
    hw = x.numel() // (b * x.shape[-1])
    check(_lib.lib().csd_gn_bwd_stats_bf16(_ptr(x), c, x.shape[-1], _ptr(dy), dy.shape[-1], dy_c_off, _ptr(fwd_coef),
                                           _ptr(s), s.shape[1], s_c_off, b, hw, int(silu), _stream()))


def gn_bwd_coeffs(sums0, c0, sums1, c1, gamma, s, bwd_coef, dgamma, dbeta, hw, groups, eps=1e-6):
    b = sums0.shape[0]
    check(_lib.lib().csd_gn_bwd_coeffs_f32(_ptr(sums0), c0, _ptr(sums1), c1, _ptr(gamma), _ptr(s), _ptr(bwd_coef),
                                           _ptr(dgamma), _ptr(dbeta), b, hw, groups, float(eps), _stream()))


def gn_bwd_apply(x, c, dy, dy_c_off, fwd_coef, bwd_coef, b_c_off, dx, silu, accumulate):
    b = x.shape[0]
    hw = x.numel() // (b * x.shape[-1])
    check(_lib.lib().csd_gn_bwd_apply_bf16(_ptr(x), c, x.shape[-1], _ptr(dy), dy.shape[-1], dy_c_off, _ptr(fwd_coef),
                                           _ptr(bwd_coef), bwd_coef.shape[1], b_c_off, _ptr(dx), dx.shape[-1], b, hw,
                                           int(silu), int(accumulate), _stream()))


def fir_resample_bwd(g, din, mode, taps, accumulate=False):
    """g: gradient of the forward output, din [B, h, w, pitch]: gradient of the forward input."""
    b, h, w, pitch = din.shape
    arr = (ctypes.c_float * 4)(*[float(t) for t in taps])
    check(_lib.lib().csd_fir_resample_bwd_nhwc_bf16(_ptr(g), _ptr(din), b, h, w, pitch,
                                                    {"up": 1, "down": 2, "prefilter": 3}[mode], arr, int(accumulate),
                                                    _stream()))
    return din


def softmax_bwd(probs, dp, ds, cols, scale):
    rows = probs.numel() // probs.shape[-1]
    check(_lib.lib().csd_softmax_bwd_bf16(_ptr(probs), probs.shape[-1], _ptr(dp), dp.shape[-1], _ptr(ds), ds.shape[-1],
                                          rows, cols, float(scale), _stream()))
    return ds


def transpose(src, out, rows, cols):
    """src [z, rows, pitch_in] -> out [z, cols, pitch_out] (bf16)."""
    z = src.shape[0]
    check(_lib.lib().csd_transpose_bf16(_ptr(src), src.shape[-1], src.shape[-2] * src.shape[-1], _ptr(out), out.shape[-1],
                                        out.shape[-2] * out.shape[-1], rows, cols, z, _stream()))
    return out


def axpy(src, dst, alpha=1.0, accumulate=True):
    assert src.numel() == dst.numel()
    check(_lib.lib().csd_axpy_bf16(_ptr(src), _ptr(dst), src.numel(), float(alpha), int(accumulate), _stream()))
    return dst


def dropout(x, out, p, seed_dev, salt):
    check(_lib.lib().csd_dropout_bf16(_ptr(x), _ptr(out), x.numel(), float(p), _ptr(seed_dev), int(salt), _stream()))
    return out


def zero_stuff(src, dst, stride, offset):
    b, h, w, pitch = src.shape
    check(_lib.lib().csd_zero_stuff_nhwc_bf16(_ptr(src), _ptr(dst), b, h, w, dst.shape[1], dst.shape[2], pitch, stride,
                                              offset, _stream()))
    return dst


def nchw_grad_to_nhwc(g0, c0, rs0, g1, c1, rs1, out):
    b, h, w, cpad = out.shape
    check(_lib.lib().csd_nchw_grad_to_nhwc_bf16(_ptr(g0), c0, _ptr(rs0), _ptr(g1), c1, _ptr(rs1), _ptr(out), cpad, b, h, w,
                                                _stream()))
    return out


def bias_temb_grad(chan_sums, c, scale, dbias0=None, dbias1=None, dtproj=None, tproj_pitch=0):
    check(_lib.lib().csd_bias_temb_grad_f32(_ptr(chan_sums), chan_sums.shape[1], chan_sums.shape[0], c, float(scale), _ptr(dbias0),
                                            _ptr(dbias1), _ptr(dtproj), tproj_pitch, _stream()))


def sgemm_small(ta, tb, m, n, k, a, lda, b, ldb, c, ldc, alpha=1.0, beta=0.0, bias=None):
    check(_lib.lib().csd_sgemm_small_f32(int(ta), int(tb), m, n, k, float(alpha), _ptr(a), lda, _ptr(b), ldb, float(beta),
                                         _ptr(c), ldc, _ptr(bias), _stream()))
    return c


def silu_f32(x, y, dy=None):
    check(_lib.lib().csd_silu_f32(_ptr(x), _ptr(dy), _ptr(y), x.numel(), int(dy is not None), _stream()))
    return y


def time_features(labels, nf, embedding_type, fourier_w, emb):
    check(_lib.lib().csd_time_features_f32(_ptr(labels), labels.shape[0], nf, 1 if embedding_type == "fourier" else 0,
                                           _ptr(fourier_w), _ptr(emb), _stream()))
    return emb


def dsm_loss_bwd(score, z, a, c, w, grad_losses, dscore):
    b, ps = _per_sample(score)
    check(_lib.lib().csd_dsm_loss_bwd_f32(_ptr(score), _ptr(z), _ptr(a), _ptr(c), _ptr(w), _ptr(grad_losses), _ptr(dscore),
                                          b, ps, _stream()))
    return dscore
