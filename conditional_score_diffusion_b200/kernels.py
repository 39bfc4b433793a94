"""Torch-tensor front ends of the C ABI (pointer extraction, shape checks, stream).

torch is used here for device memory and streams only. Every function launches CUDA kernels from
libcsd_b200.so on torch's current stream and raises on any error; none has a PyTorch fallback.
"""
import ctypes
import math

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
    return out.reshape(n_pad, kh * kw * cpad).to(dtype).contiguous()


_TILE_CACHE = {}


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


def conv_gemm(segments, wt, n, out, *, batch, h, w, out_pitch=None, n_store=None, n_tile=None,
              tile=None, bias=None, bias_per_row=False, temb=None, temb_pitch=0, res=None,
              res_pitch=0, scale=1.0, out_f32=None, z_batches=1, a_batch_step=0, wt_batch_stride=0,
              out_z_stride=0, res_z_stride=0, wt_pitch=0, wt_k_off=0, k_valid=0, wt_rows=None):
    """Launch csd_conv_gemm. segments: list of (tensor, pitch, c_off, c_cnt, taps)."""
    _require_cuda(wt, out, bias, temb, res, *[s[0] for s in segments])
    d = ConvGemmDesc()
    d.batch, d.h, d.w = batch, h, w
    tw, th, tb = tile if tile is not None else pick_tile(h, w, batch)
    d.tile_w, d.tile_h, d.tile_b = tw, th, tb
    d.nseg = len(segments)
    k_total = 0
    for i, (a, pitch, c_off, c_cnt, taps) in enumerate(segments):
        assert a.dtype == _BF16
        d.seg[i].a = a.data_ptr()
        d.seg[i].pitch = pitch
        d.seg[i].c_off = c_off
        d.seg[i].c_cnt = c_cnt
        d.seg[i].taps = taps
        k_total += taps * ceil_to(c_cnt, 32)
    d.n = n
    d.n_store = n_store if n_store is not None else n
    if n_tile is None:
        n16 = ceil_to(d.n_store, 16)
        n_tile = n16 if n16 <= 256 else (ceil_to(n16 // 2, 16) if n16 <= 512 else 256)
    d.n_tile = n_tile
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
    check(_lib.lib().csd_conv_gemm(ctypes.byref(d), _stream()))
    return out
