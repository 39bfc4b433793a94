"""Oracle restatement of the NCSN++ score network forward pass (CPU, fp32, functional).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The network is described by a plain `spec` (list of module records in the order of the
reference's `all_modules` ModuleList, models/ncsnpp.py:73-236) and evaluated over a flat
parameter dict whose keys are the reference state-dict names (`all_modules.<i>.<Sub>.<param>`).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

SQRT2 = math.sqrt(2.0)


def model_options(config):
    """Collect the hyper-parameters NCSNpp.__init__ reads (models/ncsnpp.py:43-71)."""
    m, d = config.model, config.data
    return SimpleNamespace(
        nf=m.nf, ch_mult=tuple(m.ch_mult), num_res_blocks=m.num_res_blocks,
        attn_resolutions=tuple(m.attn_resolutions), resamp_with_conv=m.resamp_with_conv,
        conditional=m.conditional, fir=m.fir, fir_kernel=tuple(m.fir_kernel),
        skip_rescale=m.skip_rescale, resblock_type=m.resblock_type.lower(),
        progressive=m.progressive.lower(), progressive_input=m.progressive_input.lower(),
        embedding_type=m.embedding_type.lower(), init_scale=m.init_scale,
        combine_method=m.progressive_combine.lower(), channels=d.num_channels,
        image_size=d.effective_image_size, centered=d.centered,
        fourier_scale=getattr(m, "fourier_scale", 16.0))


def build_spec(o):
    """Module order of NCSNpp.__init__ (models/ncsnpp.py:73-236) as (kind, info) records."""
    spec = []
    nf = o.nf
    num_res = len(o.ch_mult)
    res_sizes = [o.image_size // (2 ** i) for i in range(num_res)]
    if o.embedding_type == "fourier":
        spec.append(("fourier", {}))
    if o.conditional:
        spec.append(("linear", {}))
        spec.append(("linear", {}))
    spec.append(("conv3x3", {}))
    hs_c = [nf]
    in_ch = nf
    ipc = o.channels
    for lvl in range(num_res):
        for _ in range(o.num_res_blocks):
            out_ch = nf * o.ch_mult[lvl]
            spec.append(("resblock", {"in": in_ch, "out": out_ch, "up": False, "down": False}))
            in_ch = out_ch
            if res_sizes[lvl] in o.attn_resolutions:
                spec.append(("attn", {"ch": in_ch}))
            hs_c.append(in_ch)
        if lvl != num_res - 1:
            if o.resblock_type == "ddpm":
                spec.append(("downsample", {"ch": in_ch}))
            else:
                spec.append(("resblock", {"in": in_ch, "out": in_ch, "up": False, "down": True}))
            if o.progressive_input == "input_skip":
                spec.append(("combine", {"dim1": ipc, "dim2": in_ch}))
                if o.combine_method == "cat":
                    in_ch *= 2
            elif o.progressive_input == "residual":
                spec.append(("pyramid_down", {"in": ipc, "out": in_ch}))
                ipc = in_ch
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    spec.append(("resblock", {"in": in_ch, "out": in_ch, "up": False, "down": False}))
    spec.append(("attn", {"ch": in_ch}))
    spec.append(("resblock", {"in": in_ch, "out": in_ch, "up": False, "down": False}))
    pyramid_ch = 0
    for lvl in reversed(range(num_res)):
        for _ in range(o.num_res_blocks + 1):
            out_ch = nf * o.ch_mult[lvl]
            spec.append(("resblock", {"in": in_ch + hs_c.pop(), "out": out_ch, "up": False, "down": False}))
            in_ch = out_ch
        if res_sizes[lvl] in o.attn_resolutions:
            spec.append(("attn", {"ch": in_ch}))
        if o.progressive != "none":
            if lvl == num_res - 1:
                spec.append(("gn", {"ch": in_ch}))
                spec.append(("conv3x3", {}))
                pyramid_ch = o.channels if o.progressive == "output_skip" else in_ch
            else:
                if o.progressive == "output_skip":
                    spec.append(("gn", {"ch": in_ch}))
                    spec.append(("conv3x3", {}))
                    pyramid_ch = o.channels
                else:
                    spec.append(("pyramid_up", {"in": pyramid_ch, "out": in_ch}))
                    pyramid_ch = in_ch
        if lvl != 0:
            if o.resblock_type == "ddpm":
                spec.append(("upsample", {"ch": in_ch}))
            else:
                spec.append(("resblock", {"in": in_ch, "out": in_ch, "up": True, "down": False}))
    assert not hs_c
    if o.progressive != "output_skip":
        spec.append(("gn", {"ch": in_ch}))
        spec.append(("conv3x3", {}))
    return spec


# ---- leaf ops ------------------------------------------------------------------------------
def _p(params, idx, name):
    return params[f"all_modules.{idx}.{name}"]


def _silu(x):
    return x * torch.sigmoid(x)


def _gn(params, prefix, x):
    """nn.GroupNorm(min(C//4, 32), C, eps=1e-6) (models/layerspp.py:67,219,231)."""
    c = x.shape[1]
    return F.group_norm(x, min(c // 4, 32), params[prefix + ".weight"], params[prefix + ".bias"], eps=1e-6)


def _conv(params, prefix, x, padding):
    return F.conv2d(x, params[prefix + ".weight"], params[prefix + ".bias"], padding=padding)


def _nin(params, prefix, x):
    """models/layers.py:555-564: per-pixel x @ W + b with W [in, out]."""
    w, b = params[prefix + ".W"], params[prefix + ".b"]
    return torch.einsum("bchw,cd->bdhw", x, w) + b.view(1, -1, 1, 1)


def timestep_embedding(t, dim, max_positions=10000):
    """models/layers.py:524-538."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    arg = t.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def resblock_biggan(params, idx, x, temb, info, o):
    """ResnetBlockBigGANpp.forward (models/layerspp.py:242-274)."""
    pre = f"all_modules.{idx}"
    h = _silu(_gn(params, pre + ".GroupNorm_0", x))
    if info["up"]:
        if o.fir:
            h, x = ops.upsample_2d(h, o.fir_kernel), ops.upsample_2d(x, o.fir_kernel)
        else:
            h, x = (t.repeat_interleave(2, 2).repeat_interleave(2, 3) for t in (h, x))
    elif info["down"]:
        if o.fir:
            h, x = ops.downsample_2d(h, o.fir_kernel), ops.downsample_2d(x, o.fir_kernel)
        else:
            h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    h = _conv(params, pre + ".Conv_0", h, 1)
    if temb is not None:
        h = h + F.linear(_silu(temb), params[pre + ".Dense_0.weight"], params[pre + ".Dense_0.bias"])[:, :, None, None]
    h = _silu(_gn(params, pre + ".GroupNorm_1", h))
    h = _conv(params, pre + ".Conv_1", h, 1)  # dropout is identity in eval mode
    if info["in"] != info["out"] or info["up"] or info["down"]:
        x = _conv(params, pre + ".Conv_2", x, 0)
    return (x + h) / SQRT2 if o.skip_rescale else x + h


def resblock_ddpmpp(params, idx, x, temb, info, o):
    """ResnetBlockDDPMpp.forward (models/layerspp.py:195-209)."""
    pre = f"all_modules.{idx}"
    h = _silu(_gn(params, pre + ".GroupNorm_0", x))
    h = _conv(params, pre + ".Conv_0", h, 1)
    if temb is not None:
        h = h + F.linear(_silu(temb), params[pre + ".Dense_0.weight"], params[pre + ".Dense_0.bias"])[:, :, None, None]
    h = _silu(_gn(params, pre + ".GroupNorm_1", h))
    h = _conv(params, pre + ".Conv_1", h, 1)
    if info["in"] != info["out"]:
        x = _nin(params, pre + ".NIN_0", x)
    return (x + h) / SQRT2 if o.skip_rescale else x + h


def attn_block(params, idx, x, o):
    """AttnBlockpp.forward (models/layerspp.py:75-91)."""
    pre = f"all_modules.{idx}"
    b, c, hh, ww = x.shape
    h = _gn(params, pre + ".GroupNorm_0", x)
    q = _nin(params, pre + ".NIN_0", h).reshape(b, c, hh * ww)
    k = _nin(params, pre + ".NIN_1", h).reshape(b, c, hh * ww)
    v = _nin(params, pre + ".NIN_2", h).reshape(b, c, hh * ww)
    logits = torch.einsum("bci,bcj->bij", q, k) * (int(c) ** (-0.5))
    probs = torch.softmax(logits, dim=-1)
    h = torch.einsum("bij,bcj->bci", probs, v).reshape(b, c, hh, ww)
    h = _nin(params, pre + ".NIN_3", h)
    return (x + h) / SQRT2 if o.skip_rescale else x + h


def forward(params, o, x, time_cond, spec=None):
    """NCSNpp.forward (models/ncsnpp.py:238-388). x [B,C,H,W] fp32, time_cond [B]."""
    spec = spec or build_spec(o)
    i = 0
    if o.embedding_type == "fourier":
        w = _p(params, i, "W")
        proj = time_cond[:, None] * w[None, :] * 2 * np.pi  # layerspp.py:39-41
        temb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
        i += 1
    else:
        temb = timestep_embedding(time_cond, o.nf)
    if o.conditional:
        temb = F.linear(temb, _p(params, i, "weight"), _p(params, i, "bias"))
        i += 1
        temb = F.linear(_silu(temb), _p(params, i, "weight"), _p(params, i, "bias"))
        i += 1
    else:
        temb = None
    if not o.centered:
        x = 2 * x - 1.0
    input_pyramid = x if o.progressive_input != "none" else None

    def block(idx, t, tb):
        kind, info = spec[idx]
        assert kind == "resblock", (idx, kind)
        fn = resblock_biggan if o.resblock_type == "biggan" else resblock_ddpmpp
        return fn(params, idx, t, tb, info, o)

    hs = [_conv(params, f"all_modules.{i}", x, 1)]
    i += 1
    num_res = len(o.ch_mult)
    for lvl in range(num_res):
        for _ in range(o.num_res_blocks):
            h = block(i, hs[-1], temb)
            i += 1
            if h.shape[-1] in o.attn_resolutions:
                h = attn_block(params, i, h, o)
                i += 1
            hs.append(h)
        if lvl != num_res - 1:
            if o.resblock_type == "ddpm":
                h = _downsample_module(params, i, hs[-1], o)
            else:
                h = block(i, hs[-1], temb)
            i += 1
            if o.progressive_input == "input_skip":
                input_pyramid = ops.downsample_2d(input_pyramid, o.fir_kernel) if o.fir else F.avg_pool2d(input_pyramid, 2)
                hc = _conv(params, f"all_modules.{i}.Conv_0", input_pyramid, 0)  # Combine (layerspp.py:52-59)
                h = torch.cat([hc, h], dim=1) if o.combine_method == "cat" else hc + h
                i += 1
            elif o.progressive_input == "residual":
                input_pyramid = _pyramid_resample(params, i, input_pyramid, o, down=True)
                i += 1
                input_pyramid = (input_pyramid + h) / SQRT2 if o.skip_rescale else input_pyramid + h
                h = input_pyramid
            hs.append(h)

    h = hs[-1]
    h = block(i, h, temb); i += 1
    h = attn_block(params, i, h, o); i += 1
    h = block(i, h, temb); i += 1

    pyramid = None
    for lvl in reversed(range(num_res)):
        for _ in range(o.num_res_blocks + 1):
            h = block(i, torch.cat([h, hs.pop()], dim=1), temb)
            i += 1
        if h.shape[-1] in o.attn_resolutions:
            h = attn_block(params, i, h, o)
            i += 1
        if o.progressive != "none":
            if lvl == num_res - 1:
                pyramid = _silu(_gn(params, f"all_modules.{i}", h)); i += 1
                pyramid = _conv(params, f"all_modules.{i}", pyramid, 1); i += 1
            elif o.progressive == "output_skip":
                pyramid = ops.upsample_2d(pyramid, o.fir_kernel) if o.fir else \
                    pyramid.repeat_interleave(2, 2).repeat_interleave(2, 3)
                ph = _silu(_gn(params, f"all_modules.{i}", h)); i += 1
                ph = _conv(params, f"all_modules.{i}", ph, 1); i += 1
                pyramid = pyramid + ph
            else:  # residual
                pyramid = _pyramid_resample(params, i, pyramid, o, down=False); i += 1
                pyramid = (pyramid + h) / SQRT2 if o.skip_rescale else pyramid + h
                h = pyramid
        if lvl != 0:
            if o.resblock_type == "ddpm":
                h = _upsample_module(params, i, h, o)
            else:
                h = block(i, h, temb)
            i += 1
    assert not hs
    if o.progressive == "output_skip":
        h = pyramid
    else:
        h = _silu(_gn(params, f"all_modules.{i}", h)); i += 1
        h = _conv(params, f"all_modules.{i}", h, 1); i += 1
    assert i == len(spec), (i, len(spec))
    return h


def _pyramid_resample(params, idx, x, o, down):
    """layerspp.Downsample/Upsample with_conv=True (models/layerspp.py:94-163)."""
    pre = f"all_modules.{idx}"
    if o.fir:
        w, b = params[pre + ".Conv2d_0.weight"], params[pre + ".Conv2d_0.bias"]
        y = ops.conv_downsample_2d(x, w, o.fir_kernel) if down else ops.upsample_conv_2d(x, w, o.fir_kernel)
        return y + b.view(1, -1, 1, 1)
    if down:
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), params[pre + ".Conv_0.weight"], params[pre + ".Conv_0.bias"], stride=2)
    y = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
    return _conv(params, pre + ".Conv_0", y, 1)


def _downsample_module(params, idx, x, o):
    """layerspp.Downsample(with_conv=resamp_with_conv) for resblock_type='ddpm'."""
    if o.resamp_with_conv:
        return _pyramid_resample(params, idx, x, o, down=True)
    return ops.downsample_2d(x, o.fir_kernel) if o.fir else F.avg_pool2d(x, 2)


def _upsample_module(params, idx, x, o):
    if o.resamp_with_conv:
        return _pyramid_resample(params, idx, x, o, down=False)
    return ops.upsample_2d(x, o.fir_kernel) if o.fir else x.repeat_interleave(2, 2).repeat_interleave(2, 3)


def forward_paired(params, o, x, y, labels, spec=None):
    """NCSNpp_paired.forward (models/ncsnpp.py:395-401): cat on channels, split the output."""
    xc = x.shape[1]
    out = forward(params, o, torch.cat([x, y], dim=1), labels, spec)
    return {"x": out[:, :xc], "y": out[:, xc:]}
