"""Shared layer definitions (reference: models/layers.py:29-132, 524-564).

Modules are parameter containers with the reference's attribute names and initialisers, so that
reference checkpoints load unchanged; their arithmetic is executed by the CUDA engine
(conditional_score_diffusion_b200/engine.py), never by torch ops.
"""
import math

import numpy as np
import torch
import torch.nn as nn


def get_act(config):
    """models/layers.py:29-41. Only swish (SiLU) is implemented by the CUDA kernels."""
    name = config.model.nonlinearity.lower()
    if name == "swish":
        return nn.SiLU()
    raise NotImplementedError(f"nonlinearity '{name}': the B200 kernels fuse SiLU only")


def variance_scaling(scale, mode, distribution, in_axis=1, out_axis=0, dtype=torch.float32, device="cpu"):
    """JAX-style variance scaling (models/layers.py:54-84)."""

    def _compute_fans(shape):
        receptive = np.prod(shape) / shape[in_axis] / shape[out_axis]
        return shape[in_axis] * receptive, shape[out_axis] * receptive

    def init(shape, dtype=dtype, device=device):
        fan_in, fan_out = _compute_fans(shape)
        denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        variance = scale / denom
        if distribution == "normal":
            return torch.randn(*shape, dtype=dtype, device=device) * np.sqrt(variance)
        if distribution == "uniform":
            return (torch.rand(*shape, dtype=dtype, device=device) * 2.0 - 1.0) * np.sqrt(3 * variance)
        raise ValueError("invalid distribution for variance scaling initializer")

    return init


def default_init(scale=1.0):
    """DDPM initialisation (models/layers.py:87-91): scale 0 means 1e-10."""
    scale = 1e-10 if scale == 0 else scale
    return variance_scaling(scale, "fan_avg", "uniform")


def ddpm_conv1x1(in_planes, out_planes, stride=1, bias=True, init_scale=1.0, padding=0):
    """models/layers.py:100-105."""
    conv = nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, padding=padding, bias=bias)
    conv.weight.data = default_init(init_scale)(conv.weight.data.shape)
    nn.init.zeros_(conv.bias)
    return conv


def ddpm_conv3x3(in_planes, out_planes, stride=1, bias=True, dilation=1, init_scale=1.0, padding=1):
    """models/layers.py:119-132 (2-D case)."""
    conv = nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=padding, dilation=dilation,
                     bias=bias)
    conv.weight.data = default_init(init_scale)(conv.weight.data.shape)
    nn.init.zeros_(conv.bias)
    return conv


class NIN(nn.Module):
    """Per-pixel linear layer, W [in, out] (models/layers.py:555-564)."""

    def __init__(self, in_dim, num_units, init_scale=0.1):
        super().__init__()
        self.W = nn.Parameter(default_init(scale=init_scale)((in_dim, num_units)), requires_grad=True)
        self.b = nn.Parameter(torch.zeros(num_units), requires_grad=True)


class AttnBlock(nn.Module):
    """DDPM channel-wise self-attention block (models/layers.py:567-590): GroupNorm(32) -> q,k,v NIN ->
    softmax(q k^T / sqrt(C)) v -> NIN -> x + h."""

    def __init__(self, channels):
        super().__init__()
        self.GroupNorm_0 = nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6)
        self.NIN_0 = NIN(channels, channels)
        self.NIN_1 = NIN(channels, channels)
        self.NIN_2 = NIN(channels, channels)
        self.NIN_3 = NIN(channels, channels, init_scale=0.0)


class Upsample(nn.Module):
    """Nearest x2 (+ 3x3 conv) (models/layers.py:593-604)."""

    def __init__(self, channels, with_conv=False):
        super().__init__()
        if with_conv:
            self.Conv_0 = ddpm_conv3x3(channels, channels)
        self.with_conv = with_conv


class Downsample(nn.Module):
    """pad (0,1,0,1) + 3x3 stride-2 VALID conv, or 2x2 average pooling (models/layers.py:607-629)."""

    def __init__(self, channels, with_conv=False, dim=2):
        super().__init__()
        if dim != 2:
            raise NotImplementedError("3-D DDPM layers are outside the B200 hot path (SURVEY.md §2)")
        if with_conv:
            self.Conv_0 = ddpm_conv3x3(channels, channels, stride=2, padding=0)
        self.with_conv = with_conv


class ResnetBlockDDPM(nn.Module):
    """The ResNet block used in DDPM (models/layers.py:632-675): GroupNorm(32) everywhere, NIN shortcut when the
    channel count changes, plain `x + h`."""

    def __init__(self, act, in_ch, out_ch=None, temb_dim=None, conv_shortcut=False, dropout=0.1, dim=2):
        super().__init__()
        if dim != 2:
            raise NotImplementedError("3-D DDPM layers are outside the B200 hot path (SURVEY.md §2)")
        out_ch = out_ch if out_ch is not None else in_ch
        self.GroupNorm_0 = nn.GroupNorm(num_groups=32, num_channels=in_ch, eps=1e-6)
        self.act = act
        self.Conv_0 = ddpm_conv3x3(in_ch, out_ch)
        if temb_dim is not None:
            self.Dense_0 = nn.Linear(temb_dim, out_ch)
            self.Dense_0.weight.data = default_init()(self.Dense_0.weight.data.shape)
            nn.init.zeros_(self.Dense_0.bias)
        self.GroupNorm_1 = nn.GroupNorm(num_groups=32, num_channels=out_ch, eps=1e-6)
        self.Dropout_0 = nn.Dropout(dropout)
        self.Conv_1 = ddpm_conv3x3(out_ch, out_ch, init_scale=0.0)
        if in_ch != out_ch:
            if conv_shortcut:
                self.Conv_2 = ddpm_conv3x3(in_ch, out_ch)
            else:
                self.NIN_0 = NIN(in_ch, out_ch)
        self.out_ch, self.in_ch, self.conv_shortcut = out_ch, in_ch, conv_shortcut
        self.skip_rescale = False
