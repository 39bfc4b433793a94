"""`op.upfirdn2d` (reference: op/upfirdn2d.py:19-156) on the libcsd_b200 kernel.

Same signature and autograd structure as the reference: the forward and every derivative are the
same native op with up/down swapped, the flipped kernel and the g_pad padding
(op/upfirdn2d.py:25-44,110-115), so double backward works. CUDA tensors only: the reference's CPU
branch (`upfirdn2d_native`, :159-200) has no counterpart here and a CPU tensor raises.
"""
import torch
from torch.autograd import Function

from .. import kernels as K


def _native(x4, kernel, up, down, pad):
    n, c, h, w = x4.shape
    out = K.upfirdn2d_planes(x4.reshape(n * c, h, w), kernel, up[0], up[1], down[0], down[1], pad[0], pad[1],
                             pad[2], pad[3])
    return out.view(n, c, out.shape[-2], out.shape[-1])


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        grad_input = _native(grad_output.contiguous(), grad_kernel, down, up, g_pad)
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input.view(in_size)

    @staticmethod
    def backward(ctx, gradgrad_input):
        (kernel,) = ctx.saved_tensors
        gg = _native(gradgrad_input.contiguous().view(ctx.in_size), kernel, ctx.up, ctx.down, ctx.pad)
        return gg, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = input.shape
        out = _native(input.contiguous(), kernel, up, down, pad)
        out_h, out_w = out.shape[-2:]
        ctx.out_size = (out_h, out_w)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.up, ctx.down, ctx.pad = up, down, pad
        # padding of the transposed operator (op/upfirdn2d.py:110-115)
        g_pad_x0 = kernel_w - pad_x0 - 1
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad,
                                             ctx.g_pad, ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """input [N, C, H, W], kernel [kh, kw]; same `up`/`down`/`pad` on both axes (op/upfirdn2d.py:145-156)."""
    if input.device.type != "cuda":
        raise RuntimeError("conditional_score_diffusion_b200.op.upfirdn2d runs on CUDA tensors only "
                           "(no CPU fallback); got device %s" % input.device)
    kernel = kernel.to(device=input.device, dtype=torch.float32)
    x = input if input.dtype == torch.float32 else input.float()
    out = UpFirDn2d.apply(x, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
    return out if input.dtype == torch.float32 else out.to(input.dtype)
