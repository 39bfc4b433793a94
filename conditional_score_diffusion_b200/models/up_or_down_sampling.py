"""FIR resampling helpers (reference: models/up_or_down_sampling.py:23-56, 181-257) on `op.upfirdn2d`.

The filter taps are cached per (taps, factor, device) instead of being re-uploaded with
`torch.tensor(k, device=...)` on every call (up_or_down_sampling.py:140,176,223,256).
"""
import numpy as np
import torch
import torch.nn as nn

from ..op import upfirdn2d

_KERNEL_CACHE = {}


def _setup_kernel(k):
    """up_or_down_sampling.py:181-188."""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    assert k.ndim == 2 and k.shape[0] == k.shape[1]
    return k


def _device_kernel(k, scale, device):
    key = (tuple(np.asarray(k, dtype=np.float32).reshape(-1).tolist()), float(scale), str(device))
    t = _KERNEL_CACHE.get(key)
    if t is None:
        t = torch.tensor(_setup_kernel(k) * scale, device=device, dtype=torch.float32)
        _KERNEL_CACHE[key] = t
    return t


def upsample_2d(x, k=None, factor=2, gain=1):
    """up_or_down_sampling.py:195-224."""
    assert isinstance(factor, int) and factor >= 1
    if k is None:
        k = [1] * factor
    kk = _device_kernel(k, gain * (factor ** 2), x.device)
    p = kk.shape[0] - factor
    return upfirdn2d(x, kk, up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=None, factor=2, gain=1):
    """up_or_down_sampling.py:227-257."""
    assert isinstance(factor, int) and factor >= 1
    if k is None:
        k = [1] * factor
    kk = _device_kernel(k, gain, x.device)
    p = kk.shape[0] - factor
    return upfirdn2d(x, kk, down=factor, pad=((p + 1) // 2, p // 2))


class Conv2d(nn.Module):
    """StyleGAN2 conv with fused FIR up/down sampling (up_or_down_sampling.py:23-56): parameters only;
    evaluated by the engine."""

    def __init__(self, in_ch, out_ch, kernel, up=False, down=False, resample_kernel=(1, 3, 3, 1), use_bias=True,
                 kernel_init=None):
        super().__init__()
        assert not (up and down)
        assert kernel >= 1 and kernel % 2 == 1
        self.weight = nn.Parameter(torch.zeros(out_ch, in_ch, kernel, kernel))
        if kernel_init is not None:
            self.weight.data = kernel_init(self.weight.data.shape)
        if use_bias:
            self.bias = nn.Parameter(torch.zeros(out_ch))
        self.up, self.down = up, down
        self.resample_kernel = resample_kernel
        self.kernel = kernel
        self.use_bias = use_bias


def naive_upsample_2d(x, factor=2):
    """up_or_down_sampling.py:59-63: nearest-neighbour repetition = the default box filter [1] * factor of upsample_2d,
    so it runs on the same CUDA op (differentiable, no reshape/repeat copies)."""
    return upsample_2d(x, None, factor=factor)


def naive_downsample_2d(x, factor=2):
    """up_or_down_sampling.py:66-69: factor x factor mean = the default box filter of downsample_2d."""
    return downsample_2d(x, None, factor=factor)


def conv_downsample_2d(x, w, k=None, factor=2, gain=1):
    """up_or_down_sampling.py:144-178: FIR pre-filter with pad ((p+1)//2, p//2), p = (len(k) - factor) + (convW - 1),
    then a stride-`factor` VALID convolution. The filter runs on `op.upfirdn2d`; the convolution on the tcgen05
    implicit-GEMM kernel (bf16 operands, fp32 accumulation - the engine's precision contract). Forward only: inside the
    networks this op is planned by the engine (`_fir_conv_down`), which owns its backward."""
    from .. import kernels as K
    assert isinstance(factor, int) and factor >= 1
    out_c, in_c, conv_h, conv_w = w.shape
    assert conv_h == conv_w
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        raise NotImplementedError("conv_downsample_2d: forward-only helper (differentiate the network, not the helper)")
    if factor not in (1, 2) or conv_h not in (1, 3):
        raise NotImplementedError("conv_downsample_2d: 1x1 / 3x3 filters with stride 1 or 2 are supported")
    if k is None:
        k = [1] * factor
    kk = _device_kernel(k, gain, x.device)
    p = (kk.shape[0] - factor) + (conv_w - 1)
    y = upfirdn2d(x, kk, pad=((p + 1) // 2, p // 2))
    b, _, h, wd = y.shape
    oh, ow = (h - conv_h) // factor + 1, (wd - conv_w) // factor + 1
    cpad = K.ceil_to(in_c, 8)
    a = torch.zeros(b, h, wd, cpad, device=x.device, dtype=torch.bfloat16)
    a[..., :in_c] = y.permute(0, 2, 3, 1)
    wt = K.pack_conv_weight(w.to(x.device))
    n_store = K.ceil_to(out_c, 8)
    out = torch.empty(b, oh, ow, n_store, device=x.device, dtype=torch.float32)
    K.conv_gemm([(a, cpad, 0, in_c, conv_h * conv_w)], wt, out_c, out, batch=b, h=oh, w=ow, n_store=n_store, stride=factor,
                pad=0, in_h=h, in_w=wd, transposed=False)
    return out[..., :out_c].permute(0, 3, 1, 2).contiguous().to(x.dtype)


def upsample_conv_2d(x, w, k=None, factor=2, gain=1):
    """up_or_down_sampling.py:72-141 cannot run in the reference: `w[..., ::-1, ::-1]` (:123) is a negative-step slice,
    which torch rejects (ValueError). No config reaches it (it sits behind Upsample(fir=True, with_conv=True) and
    progressive='residual'); kept as a named entry point that says so."""
    raise NotImplementedError("upsample_conv_2d is dead code in the reference (negative-step slicing, "
                              "models/up_or_down_sampling.py:123); nothing to be compatible with")
