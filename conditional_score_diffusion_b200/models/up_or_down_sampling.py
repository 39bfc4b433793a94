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
