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
