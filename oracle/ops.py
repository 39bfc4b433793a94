"""Oracle restatement of the reference's two native ops and the FIR resampling helpers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import torch


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """Upsample (zero insertion), pad, 2-D FIR with the flipped kernel, decimate.

    Follows op/upfirdn2d.py:159-200 (`upfirdn2d_native`) and the dispatch at :145-156
    (same `up`, `down`, `pad` for both axes): out = ((in*up + pad0 + pad1 - k) // down) + 1.
    x: [N, C, H, W] tensor, kernel: [kh, kw].
    """
    x = x.to(torch.float32)
    k = torch.as_tensor(kernel, dtype=torch.float32)
    n, c, h, w = x.shape
    kh, kw = k.shape
    pad0, pad1 = pad
    z = torch.zeros(n, c, h * up, w * up, dtype=torch.float32)
    z[:, :, ::up, ::up] = x  # zero insertion: sample i sits at i*up (op/upfirdn2d.py:168-170)

    def pad_or_crop(t, p0, p1, dim):
        if p0 > 0 or p1 > 0:
            shape = list(t.shape)
            parts = []
            if p0 > 0:
                shape[dim] = p0
                parts.append(torch.zeros(shape))
            parts.append(t)
            if p1 > 0:
                shape[dim] = p1
                parts.append(torch.zeros(shape))
            t = torch.cat(parts, dim=dim)
        if p0 < 0:
            t = t.narrow(dim, -p0, t.shape[dim] + p0)
        if p1 < 0:
            t = t.narrow(dim, 0, t.shape[dim] + p1)
        return t

    z = pad_or_crop(z, pad0, pad1, 2)
    z = pad_or_crop(z, pad0, pad1, 3)
    oh_full = z.shape[2] - kh + 1
    ow_full = z.shape[3] - kw + 1
    kf = torch.flip(k, [0, 1])  # true convolution = correlation with the flipped kernel (:187)
    acc = torch.zeros(n, c, oh_full, ow_full, dtype=torch.float32)
    for i in range(kh):
        for j in range(kw):
            acc += kf[i, j] * z[:, :, i:i + oh_full, j:j + ow_full]
    return acc[:, :, ::down, ::down].contiguous()


def setup_kernel(k):
    """models/up_or_down_sampling.py:181-188: outer product of a 1-D filter, normalised to sum 1."""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k = k / np.sum(k)
    return k


def upsample_2d(x, k=(1, 3, 3, 1), factor=2, gain=1):
    """models/up_or_down_sampling.py:195-224."""
    kk = setup_kernel(k) * (gain * factor ** 2)
    p = kk.shape[0] - factor
    return upfirdn2d(x, kk, up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=(1, 3, 3, 1), factor=2, gain=1):
    """models/up_or_down_sampling.py:227-257."""
    kk = setup_kernel(k) * gain
    p = kk.shape[0] - factor
    return upfirdn2d(x, kk, down=factor, pad=((p + 1) // 2, p // 2))


def conv_downsample_2d(x, w, k=(1, 3, 3, 1), factor=2, gain=1):
    """models/up_or_down_sampling.py:144-178: FIR pad/filter, then stride-`factor` VALID conv."""
    conv = w.shape[-1]
    kk = setup_kernel(k) * gain
    p = (kk.shape[0] - factor) + (conv - 1)
    x = upfirdn2d(x, kk, pad=((p + 1) // 2, p // 2))
    return torch.nn.functional.conv2d(x, w, stride=factor, padding=0)


def upsample_conv_2d(x, w, k=(1, 3, 3, 1), factor=2, gain=1):
    """models/up_or_down_sampling.py:72-141: transposed conv (stride `factor`) then FIR."""
    conv = w.shape[-1]
    out_c, in_c = w.shape[0], w.shape[1]
    kk = setup_kernel(k) * (gain * factor ** 2)
    p = (kk.shape[0] - factor) - (conv - 1)
    # weight [out,in,kh,kw] -> flipped, [in,out,kh,kw] for conv_transpose2d (:122-124; the reference
    # indexes with a negative step, which torch rejects, so this branch is dead there - we follow
    # the documented intent).
    wt = torch.flip(w, [2, 3]).permute(1, 0, 2, 3)
    x = torch.nn.functional.conv_transpose2d(x, wt, stride=factor, padding=0)
    return upfirdn2d(x, kk, pad=((p + 1) // 2 + factor - 1, p // 2 + 1))


def fused_bias_act(x, bias=None, refer=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """op/fused_bias_act_kernel.cu:18-49: y = act(x + b) * scale and its derivatives.

    act 1 = linear, 3 = leaky relu; grad 0 forward, 1 first derivative (sign taken from `refer`),
    2 second derivative (zero). bias broadcasts over dim 1 (op/fused_act.py:86-97).
    """
    x = x.to(torch.float32)
    if bias is not None and bias.numel() > 0:
        shape = [1, -1] + [1] * (x.ndim - 2)
        x = x + bias.to(torch.float32).reshape(shape)
    ref = refer.to(torch.float32) if (refer is not None and refer.numel() > 0) else torch.zeros_like(x)
    if act == 1:
        y = {0: x, 1: x, 2: torch.zeros_like(x)}[grad]
    elif act == 3:
        if grad == 0:
            y = torch.where(x > 0, x, x * alpha)
        elif grad == 1:
            y = torch.where(ref > 0, x, x * alpha)
        else:
            y = torch.zeros_like(x)
    else:
        raise ValueError(f"act {act} not in the reference kernel")
    return y * scale


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:86-97 (CUDA branch semantics: the slope argument is honoured)."""
    return fused_bias_act(x, bias, None, act=3, grad=0, alpha=negative_slope, scale=scale)
